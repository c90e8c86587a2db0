"""C5 — NN-search microbenchmark sweep (BASELINE.json configs[4]): scene points 100 K - 10 M, radius 1 - 10 cm,
k in {1, 16, 64}; queries/s and algorithmic GB/s against the measured HBM copy peak.

    python scripts/bench_nn.py [--points 100000,1000000,10000000] [--radii 0.01,0.02,0.05,0.10] [--ks 1,16,64]
                               [--queries 4000000] [--cpu-queries 20000] [--out gpurun_out/nn_sweep.json]

Cloud = points uniform on the six faces of a 20 x 2.8 x 15 m room (1 mm Gaussian noise); the grid is built with the
search radius (cell = 2 r, like msh_hash_grid_init_3d); queries = random cloud points + U(-r/2, r/2)^3 jitter
(SURVEY.md 8d).  Everything is resident in HBM when the timed launch starts; the launch is timed with CUDA events on
the library stream (rsgpu_profile_*), best of `--reps`.  Algorithmic bytes per query = 12 + 8 B + 16 C + 8 min(k, hits)
+ 8 with B, C counted exactly by rsgpu_grid_search_census_dev (no early-out credit).  The CPU column is the reference's
own msh_hash_grid_radius_search (oracle/_ref, its OpenMP query loop, all host cores) on a query sample.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def surface_cloud(n, rng, room=(20.0, 2.8, 15.0), noise=0.001):
    X, Y, Z = room
    areas = np.array([X * Z, X * Z, X * Y, X * Y, Y * Z, Y * Z])
    face = rng.choice(6, size=n, p=areas / areas.sum())
    u, v = rng.random(n), rng.random(n)
    p = np.zeros((n, 3))
    for f, (fix_axis, fix_val, a, b) in enumerate([(1, 0.0, 0, 2), (1, Y, 0, 2), (2, 0.0, 0, 1), (2, Z, 0, 1), (0, 0.0, 1, 2), (0, X, 1, 2)]):
        m = face == f
        p[m, fix_axis] = fix_val
        p[m, a] = u[m] * room[a]
        p[m, b] = v[m] * room[b]
    p += rng.normal(0.0, noise, p.shape)
    return np.ascontiguousarray(p, np.float32)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--points", default="100000,1000000,10000000")
    ap.add_argument("--radii", default="0.01,0.02,0.05,0.10")
    ap.add_argument("--ks", default="1,16,64")
    ap.add_argument("--queries", type=int, default=4_000_000)
    ap.add_argument("--cpu-queries", type=int, default=20_000)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--impl", default="", help="force the search kernel: lane | warp (default: the library's own choice)")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "nn_sweep.json"))
    args = ap.parse_args()
    import torch
    from rescan_b200 import api
    api.set_device(0)
    if args.impl:
        api.set_option("search_impl", args.impl)
    dev = torch.device("cuda", 0)
    peak = 6650.0
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    try:
        from oracle import refbind
        have_ref = refbind.available(openmp=True)
    except Exception:
        have_ref = False
    rng = np.random.default_rng(20191027)
    rows = []
    for n in [int(x) for x in args.points.split(",")]:
        cloud = surface_cloud(n, rng)
        for r in [float(x) for x in args.radii.split(",")]:
            r32 = float(np.float32(r))
            # the sweep builds the grid with the search radius itself (cell = 2 r), as msh_hash_grid_init_3d( pts, n, r ) does
            api.profile_reset(); api.profile_enable(True)
            grid = api.HashGrid(cloud, np.float32(r))
            api.profile_enable(False)
            build_ms = api.profile_get("grid_build")[0]
            info = grid.info()
            nq = args.queries
            q = cloud[rng.integers(0, n, nq)] + rng.uniform(-r / 2, r / 2, (nq, 3)).astype(np.float32)
            q = np.ascontiguousarray(q, np.float32)
            dq = torch.from_numpy(q).to(dev)
            nB, nC = grid.search_census_dev(dq.data_ptr(), nq, r32)
            for k in [int(x) for x in args.ks.split(",")]:
                d2 = torch.empty((nq, k), dtype=torch.float32, device=dev)
                idx = torch.empty((nq, k), dtype=torch.int32, device=dev)
                nn = torch.empty(nq, dtype=torch.int64, device=dev)
                times = []
                for _ in range(args.reps + 1):  # first = warm-up
                    api.profile_reset(); api.profile_enable(True)
                    total = grid.radius_search_dev(dq.data_ptr(), nq, r32, k, d2.data_ptr(), idx.data_ptr(), nn.data_ptr())
                    api.profile_enable(False)
                    times.append(api.profile_get("search")[0])
                best = min(times[1:])
                bytes_alg = 12 * nq + 8 * nB + 16 * nC + 8 * total + 8 * nq
                row = dict(points=n, radius=r, k=k, queries=nq, cells=[int(x) for x in info["dims"]], n_bins=info["n_bins"],
                           max_pts_in_bin=info["max_n_pts_in_bin"], build_ms=build_ms, search_ms=best, queries_per_s=nq / (best * 1e-3),
                           cells_per_query=nB / nq, points_per_query=nC / nq, hits_per_query=total / nq,
                           algorithmic_GBps=bytes_alg / (best * 1e-3) / 1e9, hbm_peak_GBps=peak, frac=bytes_alg / (best * 1e-3) / 1e9 / peak)
                if have_ref and args.cpu_queries > 0 and n <= 1_000_000:
                    rg = refbind.RefGrid(cloud, np.float32(r), openmp=True)
                    qs = q[: args.cpu_queries]
                    t0 = time.perf_counter()
                    rg.radius_search(qs, r32, k)
                    row["cpu_queries_per_s"] = len(qs) / (time.perf_counter() - t0)
                    row["cpu_cores"] = os.cpu_count()
                    rg.close()
                rows.append(row)
                print(json.dumps(row), flush=True)
                del d2, idx, nn
            grid.close()
            del dq
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(rows, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
