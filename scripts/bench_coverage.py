"""Coverage term: bit masks of candidate placements over the scan's lit 5 cm cells on the GPU against the
rasterise-and-count loop of the reference's arrangement optimiser restated in the oracle (one thread).

    python scripts/bench_coverage.py [--objects 40] [--per-object 64]"""
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
ap.add_argument("--objects", type=int, default=40)
ap.add_argument("--per-object", type=int, default=64)
ap.add_argument("--cpu-sample", type=int, default=200)
args = ap.parse_args()
api.set_device(0)
scene = synth.make_scene(n_objects=args.objects, n_static=6, room=(12.0, 2.6, 9.0), spacing=0.024, seed=synth.SEED + 40)
mn, mx = scene.scan.pos(0).min(0), scene.scan.pos(0).max(0)
res, origin = api.coverage_grid(mn, mx, 0.05)
rng = np.random.default_rng(3)
clouds, poses, owners = [], [], []
for oi, o in enumerate(scene.objects):
    if o.is_static:
        continue
    c2 = api.PointCloud(o.cloud.pos(2), o.cloud.nor(2))
    for j in range(args.per_object):
        d = synth.yaw_pose(rng.uniform(0, 6.28), rng.uniform(-0.3, 0.3), rng.uniform(-0.3, 0.3), 0.0)
        clouds.append(c2); owners.append(oi)
        poses.append(np.ascontiguousarray((o.pose.astype(np.float64) @ d.astype(np.float64)).astype(np.float32).T.reshape(16)))
poses = np.stack(poses)
api.rasterize_points(scene.scan.pos(2)[:100], None, res, origin)  # warm-up
t0 = time.perf_counter(); scn = api.rasterize_points(scene.scan.pos(2), None, res, origin); t_scn = time.perf_counter() - t0
api.coverage_masks(clouds[:2], poses[:2], res, origin, scn)
best = 1e9
for _ in range(3):
    t0 = time.perf_counter(); masks, n_lit = api.coverage_masks(clouds, poses, res, origin, scn); best = min(best, time.perf_counter() - t0)
# arrangement evaluations from the masks (what one simulated annealing move costs once the masks exist)
picks = [rng.choice(len(poses), 30, replace=False) for _ in range(2000)]
t0 = time.perf_counter()
for p in picks:
    api.coverage_score(masks[p], n_lit)
t_eval = (time.perf_counter() - t0) / len(picks)
row = dict(scan_lvl2_points=int(len(scene.scan.pos(2))), grid=[int(x) for x in res], lit_cells=int(n_lit), placements=len(poses),
           points_per_placement=float(np.mean([len(c) for c in clouds])), scan_rasterise_ms=t_scn * 1e3, masks_ms=best * 1e3,
           mask_words=int(masks.shape[1]), arrangement_eval_from_masks_us=t_eval * 1e6)
from oracle import orcbind as O  # checker / CPU timing only
ro, oo, n_cells = O.cov_grid(mn, mx, 0.05)
want = O.cov_rasterize(scene.scan.pos(2), None, ro, oo)
row["scan_grid_identical"] = bool((want == scn).all())
t0 = time.perf_counter()
ok = True
for k in range(min(args.cpu_sample, len(poses))):
    g = O.cov_rasterize(scene.objects[owners[k]].cloud.pos(2), poses[k], ro, oo)
    O.cov_score(want, g)
    bits = np.unpackbits(masks[k].view(np.uint8), bitorder="little")[:n_lit].astype(bool)
    ok &= bool((bits == (g[np.nonzero(want)[0]] > 0)).all())
row["cpu_rasterise_and_count_ms_per_placement"] = (time.perf_counter() - t0) * 1e3 / min(args.cpu_sample, len(poses))
row["masks_identical_on_sample"] = ok
print(json.dumps(row))
