"""pose_proposal drop-in (reference main.cpp + rsgpu shim) at a given scan size, with and without the look-ahead that batches
main.cpp's refinement loop (integration/rsgpu_dropin.cpp), optionally next to the pure-CPU build on the same files:
    python scripts/dropin_pp_timing.py [--points 200000] [--objects 10] [--cpu]
prints the stage times the executables report themselves and checks that both GPU runs wrote the same proposal bytes"""
import argparse
import json
import os
import re
import shutil
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "integration"))
import make_dropin_case  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--points", type=int, default=200_000)
ap.add_argument("--objects", type=int, default=10)
ap.add_argument("--static", type=int, default=2)
ap.add_argument("--room", type=float, nargs=3, default=[7.0, 2.6, 5.0])
ap.add_argument("--cpu", action="store_true")
args = ap.parse_args()
B = os.path.join(ROOT, "integration", "_build")
row = {}
blobs = {}
arms = [("gpu_lookahead", "pose_proposal_rsgpu", {"RSGPU_DROPIN_LOOKAHEAD": "1"}), ("gpu_one_by_one", "pose_proposal_rsgpu", {"RSGPU_DROPIN_LOOKAHEAD": "0"})]
if args.cpu:
    arms.append(("cpu", "pose_proposal_ref", {}))
for arm, exe, env in arms:
    folder = f"/tmp/rsgpu_pp_timing_{arm}"
    shutil.rmtree(folder, ignore_errors=True)
    db, scan, out, scan1 = make_dropin_case.write_case(folder, n_objects=args.objects, n_static=args.static, room=tuple(args.room), target_points=args.points)
    for rep in range(2 if arm != "cpu" else 1):  # the second run of a GPU arm is the warm one (context, module load)
        t0 = time.perf_counter()
        r = subprocess.run([os.path.join(B, exe), db, scan, out, "-v"], capture_output=True, text=True, timeout=3600, env=dict(os.environ, **env))
        wall = time.perf_counter() - t0
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    m1 = re.search(r"POSE PROPOSAL: Done in ([0-9.e+-]+)s", r.stdout)
    m2 = re.search(r"POSE_PROPOSAL: Computed poses in ([0-9.e+-]+)s", r.stdout)
    row[arm] = dict(wall_s=round(wall, 3), propose_s=float(m1.group(1)) if m1 else None, computed_poses_s=float(m2.group(1)) if m2 else None)
    blobs[arm] = open(os.path.join(folder, "scan1_pp", "scan1_pp.bin"), "rb").read()
row["scan_points"] = int(scan1.scan.n(0))
row["lookahead_same_bytes"] = blobs["gpu_lookahead"] == blobs["gpu_one_by_one"]
print(json.dumps(row))
