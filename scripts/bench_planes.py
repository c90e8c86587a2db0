"""One RANSAC round of the scan's wall detector (reference lib/rs/rs_pointcloud_filters.cpp:137-203): 5 000 candidate planes
counted against the level-2 scan points by rsgpu_plane_inlier_counts (host buffers in and out), next to the reference's own
evaluate_plane_model timed on a sample of the candidates (oracle/_ref, one thread as shipped).

    python scripts/bench_planes.py [C2] [--planes 5000] [--cpu-sample 100]"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rescan_b200 import api, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("workload", nargs="?", default="C2")
ap.add_argument("--planes", type=int, default=5000)
ap.add_argument("--cpu-sample", type=int, default=100)
args = ap.parse_args()
api.set_device(0)
scene = synth.make_scene(**synth.CONFIGS[args.workload]["scene"])
p, n = scene.scan.pos(2), scene.scan.nor(2)
rng = np.random.default_rng(12346)
active = np.abs(n[:, 1]) < np.float32(0.2)
idx = rng.integers(0, len(p), (args.planes, 3))
a, b, c = p[idx[:, 0]], p[idx[:, 1]], p[idx[:, 2]]
nr = np.cross((b - a).astype(np.float32), (c - a).astype(np.float32)).astype(np.float32)
with np.errstate(invalid="ignore", divide="ignore"):
    nr = (nr / np.sqrt((nr * nr).sum(1, keepdims=True), dtype=np.float32)).astype(np.float32)
planes = np.concatenate([a, nr], 1).astype(np.float32)
api.plane_inlier_counts(p[:1000], active[:1000], planes[:16], 0.033)  # warm-up
best = 1e9
for _ in range(5):
    api.profile_reset(); api.profile_enable(True)
    t0 = time.perf_counter()
    counts = api.plane_inlier_counts(p, active, planes, 0.033)
    best = min(best, time.perf_counter() - t0)
    api.profile_enable(False)
kernel_ms = api.profile_get("planes")[0]
tests = float(active.sum()) * args.planes
row = dict(workload=args.workload, points_lvl2=int(len(p)), active=int(active.sum()), planes=args.planes, call_ms=best * 1e3, kernel_ms=kernel_ms,
           point_plane_tests_per_s=tests / (kernel_ms * 1e-3) if kernel_ms > 0 else None,
           algorithmic_gb_per_s=(13.0 * len(p) * ((args.planes + 15) // 16)) / (kernel_ms * 1e-3) / 1e9 if kernel_ms > 0 else None)
from oracle import orcbind as O, refbind as R  # checker / CPU timing only
pick = rng.choice(args.planes, min(args.cpu_sample, args.planes), replace=False)
row["identical_to_oracle_on_sample"] = bool((counts[pick] == O.plane_inlier_counts(p, active, planes[pick], 0.033)).all())
if R.available():
    t0 = time.perf_counter()
    ref = R.plane_inlier_counts(p, active.astype(np.float64), planes[pick], 0.033)
    dt = time.perf_counter() - t0
    row["reference_cpu_ms_per_round"] = dt / len(pick) * args.planes * 1e3
    row["identical_to_reference_on_sample"] = bool((counts[pick] == ref).all())
print(json.dumps(row))
