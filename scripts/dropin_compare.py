"""The two drop-in executables against the pure-CPU reference build of segment_transfer on one synthetic scan pair of a
given size, run on the same files in the same call (integration/_build/*, built where /root/reference exists):

    pose_proposal_rsgpu  first-scan database + rescan  ->  proposals          (GPU; the CPU build takes minutes at this size)
    segment_transfer_rsgpu / segment_transfer_ref on that output  ->  placements, poses, per-vertex labels, stage times

    python scripts/dropin_compare.py [--points 200000] [--objects 10] [--static 2] [--room 7 2.6 5] [--skip-cpu]
                                     [--pp-exe pose_proposal_rsgpu_levels] [--st-exe segment_transfer_rsgpu_all]"""
import argparse
import json
import os
import re
import shutil
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "integration"))
import make_dropin_case  # noqa: E402
from rescan_b200 import rsio  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--points", type=int, default=200_000)
ap.add_argument("--objects", type=int, default=10)
ap.add_argument("--static", type=int, default=2)
ap.add_argument("--room", type=float, nargs=3, default=[7.0, 2.6, 5.0])
ap.add_argument("--folder", default="/tmp/rsgpu_dropin_compare")
ap.add_argument("--skip-cpu", action="store_true")
ap.add_argument("--pp-exe", default="pose_proposal_rsgpu", help="or pose_proposal_rsgpu_levels (level building bound as well)")
ap.add_argument("--st-exe", default="segment_transfer_rsgpu", help="or segment_transfer_rsgpu_planes / segment_transfer_rsgpu_all")
args = ap.parse_args()
B = os.path.join(ROOT, "integration", "_build")
shutil.rmtree(args.folder, ignore_errors=True)
db, scan, out, scan1 = make_dropin_case.write_case(args.folder, n_objects=args.objects, n_static=args.static, room=tuple(args.room),
                                                   target_points=args.points)
row = dict(scan_points=int(scan1.scan.n(0)), objects=args.objects, static=args.static, room=args.room)
t0 = time.perf_counter()
r = subprocess.run([os.path.join(B, args.pp_exe), db, scan, out, "-v"], capture_output=True, text=True, timeout=900)
assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
row["pose_proposal_rsgpu_wall_s"] = time.perf_counter() - t0
row["proposals_per_object"] = [len(p) for p in rsio.read_proposals(os.path.join(args.folder, "scan1_pp", "scan1_pp.bin"))]


def stages(stdout):
    pat = dict(greedy_s=r"Greedy estimation finished in ([0-9.e+-]+)s", annealing_s=r"Optimization finished in ([0-9.e+-]+)s",
               refine_s=r"Refining optimized poses done in ([0-9.e+-]+)s", segmentation_s=r"Segmentation finished in ([0-9.e+-]+)s",
               labels_ms=r"LABEL_TRANSFER:   Done in ([0-9.e+-]+)ms", neighbourhood_ms=r"Neighborhood compatibility computation took ([0-9.e+-]+)ms",
               unary_ms=r"Data and smoothness terms setting took ([0-9.e+-]+)ms", augment_s=r"Database augmentation finished in ([0-9.e+-]+)s")
    return {k: float(m.group(1)) for k, p in pat.items() for m in [re.search(p, stdout)] if m}


results = {}
row["executables"] = [args.pp_exe, args.st_exe]
for arm, exe in (("gpu", args.st_exe), ("cpu", "segment_transfer_ref")):
    if arm == "cpu" and args.skip_cpu:
        continue
    sub = os.path.join(args.folder, arm)
    os.makedirs(sub, exist_ok=True)
    t0 = time.perf_counter()
    stdout, rows, ply = make_dropin_case.run_segment_transfer(os.path.join(B, exe), out, sub)
    row[f"segment_transfer_{arm}"] = dict(wall_s=time.perf_counter() - t0, **stages(stdout))
    results[arm] = (rows, ply)
if len(results) == 2:
    (ra, pa), (rb, pb) = results["gpu"], results["cpu"]
    same = [a[0] for a in ra] == [b[0] for b in rb] and [a[2] for a in ra] == [b[2] for b in rb]
    row["same_placements"] = bool(same)
    if same:
        row["max_translation_diff_m"] = float(max([np.linalg.norm(a[4][:3, 3] - b[4][:3, 3]) for a, b in zip(ra, rb)] or [0.0]))
        row["max_rotation_diff"] = float(max([np.abs(a[4][:3, :3] - b[4][:3, :3]).max() for a, b in zip(ra, rb)] or [0.0]))
    row["labelled_vertices"] = int(len(pa))
    row["label_mismatches"] = int(((np.asarray(pa["class_idx"]) != np.asarray(pb["class_idx"])) |
                                   (np.asarray(pa["instance_idx"]) != np.asarray(pb["instance_idx"]))).sum()) if len(pa) == len(pb) else -1
print(json.dumps(row))
