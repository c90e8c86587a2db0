"""Level building: rsgpu_poisson_level for levels 1-4 of the C2 / C3 scans against the reference's
rs_pointcloud_compute_levels (oracle/_ref, one thread as shipped) and the oracle restatement.

    python scripts/bench_levels.py [C2,C3] [--cpu]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rescan_b200 import api, synth  # noqa: E402

names = (sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("-") else "C2").split(",")
cpu = "--cpu" in sys.argv
api.set_device(0)
for name in names:
    scene = synth.make_scene(**synth.CONFIGS[name]["scene"])
    p0, n0 = scene.scan.pos(0), scene.scan.nor(0)
    api.poisson_level(p0[:1000], 1)  # warm-up (module load, pool)
    row = dict(workload=name, points=len(p0), levels={})
    total = 0.0
    for lvl in range(1, 5):
        best = 1e9
        for _ in range(3):
            t0 = time.perf_counter()
            idx, rounds = api.poisson_level(p0, lvl, return_rounds=True)
            best = min(best, time.perf_counter() - t0)
        total += best
        row["levels"][lvl] = dict(samples=len(idx), rounds=rounds, gpu_ms=best * 1e3)
    row["gpu_ms_levels_1_to_4"] = total * 1e3
    if cpu:
        from oracle import orcbind as O, refbind as R
        t0 = time.perf_counter()
        want = [O.poisson_level(p0, lvl) for lvl in range(1, 5)]
        row["oracle_cpu_ms"] = (time.perf_counter() - t0) * 1e3
        row["identical_to_oracle"] = all((api.poisson_level(p0, lvl) == want[lvl - 1]).all() for lvl in range(1, 5))
        if R.available():
            t0 = time.perf_counter()
            rc = R.RefCloud.from_level0(p0, n0)
            row["reference_cpu_ms_levels_and_grids"] = (time.perf_counter() - t0) * 1e3
            row["reference_level_sizes"] = [rc.n(l) for l in range(5)]
    print(json.dumps(row), flush=True)
