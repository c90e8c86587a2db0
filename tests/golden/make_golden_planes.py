"""Golden vectors of the plane detector's inlier count: the reference's own evaluate_plane_model (compiled in place,
oracle/_ref) on a seeded scan sample -> tests/golden/planes_golden.npz.  Run in the build container only:

    python tests/golden/make_golden_planes.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import refbind as R  # noqa: E402
from tests import common  # noqa: E402

scene = common.tiny_scene()
p, n = scene.scan.pos(2)[::4].copy(), scene.scan.nor(2)[::4].copy()
rng = np.random.default_rng(20191027)
weights = (np.abs(n[:, 1]) < np.float32(0.2)).astype(np.float64)  # the wall detector's initial weights (:141-146)
weights[rng.random(len(p)) < 0.1] = 0.0  # some points already explained by an earlier plane
idx = rng.integers(0, len(p), (150, 3))
a, b, c = p[idx[:, 0]], p[idx[:, 1]], p[idx[:, 2]]
nr = np.cross((b - a).astype(np.float32), (c - a).astype(np.float32)).astype(np.float32)
nr = (nr / np.sqrt((nr * nr).sum(1, keepdims=True), dtype=np.float32)).astype(np.float32)
planes = np.concatenate([a, nr], 1).astype(np.float32)
planes[7, 3:] = np.nan  # a degenerate triple (zero cross product): the reference counts nothing
planes[11, 3:] = 0.0    # zero normal: every active point is at distance 0
out = {}
for thr in (0.033, 0.05):
    out[f"counts_{int(thr * 1000)}"] = R.plane_inlier_counts(p, weights, planes, thr).astype(np.int32)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "planes_golden.npz"), pts=p, weights=weights, planes=planes, **out)
print({k: (v[:8], int(v.sum())) for k, v in out.items()}, len(p))
