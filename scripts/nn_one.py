"""one configuration of the C5 sweep (for ncu captures of the search kernel in its HBM-bound regime):
   python scripts/nn_one.py [points=10000000] [radius=0.05] [k=16] [queries=2000000]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rescan_b200 import api  # noqa: E402
from scripts.bench_nn import surface_cloud  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
r = float(sys.argv[2]) if len(sys.argv) > 2 else 0.05
k = int(sys.argv[3]) if len(sys.argv) > 3 else 16
nq = int(sys.argv[4]) if len(sys.argv) > 4 else 2_000_000
api.set_device(0)
dev = torch.device("cuda", 0)
rng = np.random.default_rng(20191027)
cloud = surface_cloud(n, rng)
grid = api.HashGrid(cloud, np.float32(r))
q = np.ascontiguousarray(cloud[rng.integers(0, n, nq)] + rng.uniform(-r / 2, r / 2, (nq, 3)).astype(np.float32), np.float32)
dq = torch.from_numpy(q).to(dev)
d2 = torch.empty((nq, k), dtype=torch.float32, device=dev)
idx = torch.empty((nq, k), dtype=torch.int32, device=dev)
nn = torch.empty(nq, dtype=torch.int64, device=dev)
for _ in range(2):
    api.profile_reset(); api.profile_enable(True)
    total = grid.radius_search_dev(dq.data_ptr(), nq, np.float32(r), k, d2.data_ptr(), idx.data_ptr(), nn.data_ptr())
    api.profile_enable(False)
    print(f"points {n} r {r} k {k} queries {nq}: {api.profile_get('search')[0]:.3f} ms, {total} neighbours")
grid.close()
