#!/usr/bin/env python
"""bench.py — pose evaluations/s of the GPU pose_proposal hot path (BASELINE.json metric) on synthetic scans.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload C2|C3|C5] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workloads (BASELINE.json configs):
  C2 (default at N = 1)  ~200 K-point scan, 10 objects, 36 rotations x 2 048 translation seeds, one GPU.
  C3 (default at N > 1)  ~2 M-point scan, 60 objects, 72 rotations x 20 000 seeds: ONE fixed problem whose translation
                         seeds are sharded over the N ranks (reference loop apps/pose_proposal/pose_proposal.cpp:213-243)
                         => "scaling": "strong".  The line also carries the same problem run by rank 0 alone
                         (`single_gpu_same_workload`), so the scaling of THIS workload can be read from one line.
  C5 (--workload C5)     NN-search microbenchmark, the HBM-bound regime of the same gather: 10 M scene points, r = 0.10 m,
                         k = 64, 4 M queries (msh_hash_grid_radius_search, lib/msh/msh_hash_grid.h:1090-1259); metric =
                         NN queries/s, queries sharded over the ranks.

One pose step = one pass of the hot path over one scan (rescan_b200.pipeline.run_step): grid builds, dense pose search
(level 4) + verification (levels 3, 2) for every dynamic object, NMS, ICP refinement of the survivors and rescoring at
level 1, NMS, sort.  `value` times it with the scan resident in HBM; `e2e` times the same step through the host-buffer
C ABI (scan uploaded from pinned host memory and proposals read back every step).  Rank 0 prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")  # before torch creates the CUDA context (rescan_b200/api.py)

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "pose_evaluations_per_sec"
UNIT = "pose evaluations/s"
NN_METRIC = "nn_queries_per_sec"
NN_UNIT = "NN queries/s"
L2_FLUSH_BYTES = 256 << 20
C5 = dict(points=10_000_000, radius=0.10, k=64, queries=4_000_000, room=(20.0, 2.8, 15.0))


def resolve_workload(args, world):
    """N = 1: the configuration the metric is quoted on (C2); N > 1: the multi-GPU configuration BASELINE.json names (C3)"""
    if args.workload:
        return args.workload
    return "C2" if max(world, args.gpus) <= 1 else "C3"


def workload_config(name, world, nms=True, exchange="nvlink", scaling="strong"):
    from rescan_b200 import synth
    cfg = synth.CONFIGS[name]
    sc = cfg["scene"]
    total = cfg["n_seeds"] * (world if scaling == "weak" else 1)
    return {"workload": f"{name}: pose_proposal on one synthetic scene pair", "scan_points_target": sc.get("target_points"),
            "objects": sc["n_objects"], "static_objects": sc["n_static"], "rotations": cfg["n_rot"],
            "translation_seeds_total": total, "translation_seeds_per_gpu": -(-total // world),
            "top_k": 64, "parallelism": f"pose-sharded x{world} (blocks of 256 translations of the Z-order curve dealt round the ranks)", "exchange": exchange if world > 1 else None,
            "stages": "grid build, dense search lvl 4, verification lvl 3/2, top-k" + (", NMS" if nms else "") + ", ICP, rescoring lvl 1"
                      + (", NMS" if nms else "") + ", sort",
            "l2": "flushed between steps (256 MiB write inside the timed region); working set is L2-resident within a step"}


def synth_objects(name):
    from rescan_b200 import synth
    return int(synth.CONFIGS[name]["scene"]["n_objects"])


def build_workload(name, world, scaling="strong"):
    from rescan_b200 import synth, posegrid
    cfg = synth.CONFIGS[name]
    scene = synth.make_scene(**cfg["scene"])
    rotations = posegrid.rotation_xforms(cfg["n_rot"])
    translations = synth.translation_seeds(scene.scan, cfg["n_seeds"] * (world if scaling == "weak" else 1))
    return scene, rotations, translations


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower() == "active":
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def kernel_counters(key):
    """per-launch hardware counters of one kernel on one workload from the committed ncu --set full capture
    (profiles/kernel_counters.json, written by scripts/ncu_counters.py from the .ncu-rep named in its `source`; bench.py
    itself never runs under a profiler)"""
    p = os.path.join(ROOT, "profiles", "kernel_counters.json")
    try:
        return json.load(open(p)).get(key)
    except Exception:
        return None


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------------------------------------ reference arm
def cpu_reference_rate(scene, rotations, translations, target_seconds, threads=None, steps=1, warmup=0):
    """the reference's own mgs_compute_object_alignment_score (oracle/_ref, compiled from /root/reference) — or the
    oracle port when that library is absent — over a bounded, strided sample of the dense level-4 pose grid, with
    every host core (OpenMP over poses in the harness).  Returns (evals/s, info dict, per-step seconds)."""
    from oracle import refbind, orcbind
    from rescan_b200 import posegrid
    threads = threads or os.cpu_count() or 1
    dyn = [o for o in scene.objects if not o.is_static]
    n_rot, n_tr = len(rotations), len(translations)
    n_grid = n_rot * n_tr
    use_ref = refbind.available(openmp=True)
    if use_ref:
        scan = refbind.RefCloud.from_levels({l: (scene.scan.pos(l), scene.scan.nor(l)) for l in range(5)}, openmp=True)
        objs = [refbind.RefCloud.from_levels({l: (o.cloud.pos(l), o.cloud.nor(l)) for l in range(5)}, openmp=True) for o in dyn]

        def run(oi, x):
            return refbind.score_batch(objs[oi], scan, x, query_lvl=4, search_lvl=1, k=64, n_threads=threads)[1]
    else:
        og = orcbind.OrcGrid(scene.scan.pos(1), 0.05)

        def run(oi, x):
            return orcbind.score_poses(dyn[oi].cloud.pos(4), dyn[oi].cloud.nor(4), og, scene.scan.nor(1), x, 64, 0.10, threads)[1]

    def poses(ids):  # pose id = t * n_rot + r, formed like posegrid.pose_grid without materialising the whole grid
        x = np.ascontiguousarray(rotations[ids % n_rot].reshape(-1, 16).copy())
        x[:, 12:15] = translations[ids // n_rot]
        return x
    # pilot to size the sample
    pilot = poses(np.arange(0, n_grid, max(1, n_grid // 64))[:64])
    t = sum(run(oi, pilot) for oi in range(len(dyn)))
    rate = len(pilot) * len(dyn) / max(t, 1e-6)
    per_obj = int(min(n_grid, max(16, rate * target_seconds / len(dyn))))
    stride = max(1, n_grid // per_obj)
    sample = poses(np.arange(0, n_grid, stride)[:per_obj])
    times = []
    for s in range(warmup + steps):
        t = sum(run(oi, sample) for oi in range(len(dyn)))
        if s >= warmup:
            times.append(t)
    n_eval = len(sample) * len(dyn)
    info = {"cores": threads, "kind": "reference" if use_ref else "port",
            "sample": f"every {stride}th pose of the dense level-4 grid ({len(sample)} poses) x {len(dyn)} dynamic objects per step, "
                      f"mgs_compute_object_alignment_score with OpenMP over poses"}
    return n_eval / (sum(times) / len(times)), info, times


def surface_cloud(n, rng, room, noise=0.001):
    """C5 scene: points uniform on the six faces of a room with 1 mm Gaussian noise"""
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


def c5_inputs(n_points=None, n_queries=None):
    rng = np.random.default_rng(20191027)
    n, nq, r = n_points or C5["points"], n_queries or C5["queries"], C5["radius"]
    cloud = surface_cloud(n, rng, C5["room"])
    q = cloud[rng.integers(0, n, nq)] + rng.uniform(-r / 2, r / 2, (nq, 3)).astype(np.float32)
    return cloud, np.ascontiguousarray(q, np.float32)


def c5_config(world, n_points, n_queries):
    return {"workload": "C5: NN-search microbenchmark (msh_hash_grid_radius_search)", "scene_points": n_points, "radius_m": C5["radius"],
            "k": C5["k"], "queries_total": n_queries, "queries_per_gpu": -(-n_queries // world), "sort": 1,
            "parallelism": f"query-sharded x{world}", "exchange": None,
            "l2": "inputs larger than L2: 160 MB of records + 48 MB of queries + 2 GB of result rows per launch"}


def cpu_nn_rate(cloud, queries, target_seconds, steps=1, warmup=0):
    """the reference's own msh_hash_grid_radius_search (its OpenMP query loop, all host cores) on a bounded query sample"""
    from oracle import refbind, orcbind
    r32 = np.float32(C5["radius"])
    threads = os.cpu_count() or 1
    use_ref = refbind.available(openmp=True)
    if use_ref:
        g = refbind.RefGrid(cloud, r32, openmp=True)
        run = lambda qs: g.radius_search(qs, float(r32), C5["k"])
    else:
        g = orcbind.OrcGrid(cloud, float(r32))
        run = lambda qs: g.radius_search(qs, float(r32), C5["k"])
    t0 = time.perf_counter()
    run(queries[:2000])
    rate = 2000 / max(time.perf_counter() - t0, 1e-6)
    n = int(min(len(queries), max(2000, rate * target_seconds)))
    qs = np.ascontiguousarray(queries[:n])
    times = []
    for s in range(warmup + steps):
        t0 = time.perf_counter()
        run(qs)
        if s >= warmup:
            times.append(time.perf_counter() - t0)
    info = {"cores": threads if use_ref else 1, "kind": "reference" if use_ref else "port",
            "sample": f"the first {n} of the {len(queries)} queries per step, msh_hash_grid_radius_search k = {C5['k']}, r = {C5['radius']}"
                      + (", its own OpenMP query loop" if use_ref else "")}
    return n / (sum(times) / len(times)), info, times


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    name = resolve_workload(args, world)
    if name == "C5":
        cloud, q = c5_inputs(args.c5_points, args.c5_queries)
        rate, info, times = cpu_nn_rate(cloud, q, target_seconds=6.0, steps=args.steps, warmup=args.warmup)
        metric, unit, cfg = NN_METRIC, NN_UNIT, c5_config(max(world, 1), len(cloud), len(q))
    else:
        scene, rotations, translations = build_workload(name, 1)
        rate, info, times = cpu_reference_rate(scene, rotations, translations, target_seconds=8.0, steps=args.steps, warmup=args.warmup)
        metric, unit, cfg = METRIC, UNIT, workload_config(name, max(world, 1))
    line = {"impl": "reference", "metric": metric, "value": rate, "unit": unit, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
            "cpu_baseline": dict(info, value=rate, unit=unit),
            "e2e": {"value": rate, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ C5: NN search
def run_nn_workload(args, rank, world, local_rank, device, dist, emit=True, with_clocks=True):
    """query-sharded radius search, no collective on the data path: every rank holds the whole grid.  emit=False returns the
    line instead of printing it (the default C2 run embeds it as `nn_search_c5`, so that the HBM-regime kernel is measured
    in the same driver-run record)"""
    import torch
    from rescan_b200 import api, pipeline
    cloud, q_all = c5_inputs(args.c5_points, args.c5_queries)
    lo, hi = pipeline.shard_range(len(q_all), rank, world)
    q = np.ascontiguousarray(q_all[lo:hi])
    nq, k, r32 = len(q), C5["k"], float(np.float32(C5["radius"]))
    grid = api.HashGrid(cloud, np.float32(r32))
    dq = torch.from_numpy(q).to(device)
    d2 = torch.empty((nq, k), dtype=torch.float32, device=device)
    idx = torch.empty((nq, k), dtype=torch.int32, device=device)
    nn = torch.empty(nq, dtype=torch.int64, device=device)
    hq = torch.from_numpy(q).pin_memory()
    h_d2 = torch.empty((nq, k), dtype=torch.float32).pin_memory()
    h_idx = torch.empty((nq, k), dtype=torch.int32).pin_memory()
    h_nn = torch.empty(nq, dtype=torch.int64).pin_memory()
    nB, nC = grid.search_census_dev(dq.data_ptr(), nq, r32)  # exact algorithmic cell / candidate counts, untimed
    totals = []

    def step(resident):
        if resident:
            totals.append(grid.radius_search_dev(dq.data_ptr(), nq, r32, k, d2.data_ptr(), idx.data_ptr(), nn.data_ptr()))
        else:  # host buffers through the C ABI: queries uploaded, result rows copied back
            totals.append(grid.radius_search_host(hq.numpy(), r32, k, h_d2.numpy(), h_idx.numpy(), h_nn.numpy()))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(resident, n):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            step(resident)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for _ in range(args.warmup):
        step(True)
    clocks = ClockSampler(local_rank) if with_clocks else None
    if clocks:
        clocks.start()
    l0 = api.launch_count()
    api.profile_reset()
    api.profile_enable(True)
    ms_value = timed(True, args.steps)
    api.profile_enable(False)
    launches = api.launch_count() - l0
    kern_ms, kern_launches = api.profile_get("search")
    hits = totals[-1]
    step(False)
    ms_e2e = timed(False, args.steps)
    clk = clocks.stop() if clocks else None

    def total(x):
        t = torch.tensor([x], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())
    nq_all, launches_all = total(nq), total(launches)
    if rank != 0:
        return
    peak, peak_src = measured_peaks()
    bytes_alg = 12 * nq + 8 * nB + 16 * nC + 8 * hits + 8 * nq  # SURVEY.md 8d, no early-out credit, this rank's shard
    avg_ms = kern_ms / max(kern_launches, 1)
    achieved = bytes_alg / (avg_ms * 1e-3) / 1e9 if avg_ms > 0 else 0.0
    ctr = kernel_counters("C5:radius_search") or {}
    traffic = ctr.get("dram_bytes")
    line = {"metric": NN_METRIC, "value": nq_all * args.steps / (ms_value * 1e-3), "unit": NN_UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_value / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": c5_config(world, len(cloud), len(q_all)),
            "e2e": {"value": nq_all * args.steps / (ms_e2e * 1e-3), "unit": NN_UNIT, "h2d_bytes_per_step": 12 * nq_all,
                    "d2h_bytes_per_step": (8 * k + 8) * nq_all, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches_all),
            "roofline": {"bound": "hbm", "kernel": "radius_search_sub_kernel (warp per query, sub-cells ranked by gap, k-list in registers)", "achieved": achieved,
                         "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "traffic_frac_of_peak": (traffic / (avg_ms * 1e-3) / 1e9 / peak) if traffic and avg_ms > 0 else None,
                         "traffic_source": ctr.get("source"), "peak_source": peak_src, "algorithmic_bytes_per_launch": bytes_alg,
                         "launches_timed": kern_launches, "avg_launch_ms": avg_ms,
                         "cells_per_query": nB / nq, "candidates_per_query": nC / nq, "hits_per_query": hits / nq,
                         "note": "algorithmic bytes = 12 (query) + 8/cell + 16/candidate point + 8/returned neighbour + 8 per query "
                                 "(SURVEY.md 8d, no early-out credit), rank 0 shard; traffic = dram__bytes_read + dram__bytes_write of one launch "
                                 "from the committed ncu capture of the same launch"},
            "clocks": clk}
    if world == 1 and not args.no_cpu_baseline and emit:
        rate, info, _ = cpu_nn_rate(cloud, q_all, target_seconds=10.0)
        line["cpu_baseline"] = dict(info, value=rate, unit=NN_UNIT)
    grid.close()
    if not emit:
        return line
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ our arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, help="C2 | C3 | C5 (default: C2 at N = 1, C3 at N > 1)")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="N > 1: strong = the named problem split over the ranks (default); weak = every rank gets the named number of seeds")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-nn", action="store_true", help="N = 1, C2: skip the embedded C5 NN-search measurement (`nn_search_c5`)")
    ap.add_argument("--no-single", action="store_true", help="N > 1: skip the single-GPU run of the same problem on rank 0")
    ap.add_argument("--lanes", type=int, default=None, help="objects in flight at once (default RSGPU_LANES or 8); 1 = serial object loop")
    ap.add_argument("--exchange", default="nvlink", choices=["nvlink", "host", "nccl"],
                    help="N > 1: transport of the two per-step list exchanges (nvlink = peer-mapped slots written over NVLink by copy "
                         "engines, no SM; host = gloo all-gather of the host-resident lists; nccl = staged through HBM, NCCL all-gather)")
    ap.add_argument("--nms", type=int, default=1, help="1: the two NMS passes of main.cpp:161/205 run on the GPU inside the step; 0: top-k only")
    ap.add_argument("--c5-points", type=int, default=None)
    ap.add_argument("--c5-queries", type=int, default=None)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from rescan_b200 import api, pipeline

    if not torch.cuda.is_available() or api.device_count() == 0:
        raise SystemExit("bench.py: no CUDA device - the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    api.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    name = resolve_workload(args, world)
    host_group, peer = None, None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)
    if name == "C5":
        run_nn_workload(args, rank, world, local_rank, device, dist)
        if world > 1:
            dist.destroy_process_group()
        return
    exchange_name = None
    if world > 1:
        if args.exchange == "nvlink":
            try:
                from rescan_b200 import peerx
                peer = peerx.PeerExchange(dist, rank, world, local_rank, n_slots=max(64, synth_objects(name)))
                exchange_name = "nvlink (one peer-mapped slot per object chain, copy engines, no collective order; rescan_b200/peerx.py)"
            except Exception as e:
                print(f"bench.py: peer exchange unavailable ({e}); falling back to the host exchange", file=sys.stderr)
                args.exchange = "host"
        if args.exchange == "host":
            # the per-object top-k lists are host-resident and a few KB: exchanged over a gloo group next to the NCCL one
            try:
                os.environ.setdefault("GLOO_SOCKET_IFNAME", "lo")
                host_group = dist.new_group(backend="gloo")
                exchange_name = "host (gloo)"
            except Exception as e:  # no usable host transport: NCCL carries the lists (staged through HBM)
                print(f"bench.py: gloo group unavailable ({e}); exchanging over NCCL", file=sys.stderr)
                host_group = None
        if exchange_name is None:
            exchange_name = "nccl"

    scene, rotations, translations = build_workload(name, world, args.scaling)
    models = pipeline.upload_objects(scene.objects)
    p1, n1 = scene.scan.pos(1), scene.scan.nor(1)
    p2, n2 = scene.scan.pos(2), scene.scan.nor(2)
    # HBM-resident copy for `value`, pinned host copy for `e2e`
    dev = {k: torch.from_numpy(v).to(device) for k, v in dict(p1=p1, n1=n1, p2=p2, n2=n2).items()}
    scan_dev = {k: v.data_ptr() for k, v in dev.items()}
    pinned = {k: torch.from_numpy(v).pin_memory() for k, v in dict(p1=p1, n1=n1, p2=p2, n2=n2).items()}
    hp = {k: v.numpy() for k, v in pinned.items()}
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=device)

    def step(resident, lanes=None, alone=False):
        flush.zero_()
        solo = alone or world == 1
        return pipeline.run_step((hp["p1"], hp["n1"]), (hp["p2"], hp["n2"]), models, rotations, translations, top_k=64,
                                 rank=0 if solo else rank, world=1 if solo else world, dist=None if solo else dist, device=device,
                                 scan_dev=scan_dev if resident else None, nms_dist=0.2 if args.nms else None,
                                 lanes=args.lanes if lanes is None else lanes, host_group=None if solo else host_group,
                                 peer=None if solo else peer)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(resident, k, lanes=None):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        res = [step(resident, lanes) for _ in range(k)]
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), res

    # algorithmic bytes of the dense level-4 scoring (what the reference's search reads for the same poses), exact census on
    # this rank's shard, untimed; big pose grids are counted on every `cstride`-th translation and scaled
    my_ids = pipeline.shard_translations(translations, rank, world)
    cstride = max(1, len(my_ids) // 2048)
    g1 = api.HashGrid(p1, 0.05, normals=n1)
    census = [api.score_pose_grid_count(m.levels[4], g1, rotations, np.ascontiguousarray(translations[my_ids[::cstride]])) for m in models if not m.is_static]
    g1.close()
    cscale = len(my_ids) / max(len(my_ids[::cstride]), 1)
    dense_bytes = sum(c["bytes"] for c in census) * cscale
    dense_queries = sum(c["queries"] for c in census) * cscale

    for _ in range(args.warmup):
        step(True)
    clocks = ClockSampler(local_rank)
    clocks.start()
    l0 = api.launch_count()
    ms_value, res = timed(True, args.steps)
    launches = api.launch_count() - l0
    for _ in range(1):
        step(False)
    ms_e2e, res_e2e = timed(False, args.steps)
    # per-kernel device times: the same steps with ONE object in flight, so that the CUDA-event interval around a launch
    # (taken on its launch stream) is that kernel's own duration and not a share of the device
    api.profile_reset()
    api.profile_enable(True)
    ms_serial, _ = timed(True, max(1, min(args.steps, 2 if name == "C3" else args.steps)), lanes=1)
    api.profile_enable(False)
    n_prof_steps = max(1, min(args.steps, 2 if name == "C3" else args.steps))
    prof_names = ("grid_build", "score_dense", "dense_prefilter", "dense_bin", "dense_search", "dense_reduce", "score", "icp", "overlap", "nms")
    prof = {n: api.profile_get(n) for n in prof_names}  # ms over those steps, launches
    clk = clocks.stop()

    def total(x):
        t = torch.tensor([x], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    evals = total(sum(r.n_evaluations for r in res))
    queries = total(sum(r.n_queries for r in res))
    evals_e2e = total(sum(r.n_evaluations for r in res_e2e))
    h2d = total(sum(r.h2d_bytes for r in res_e2e)) / args.steps
    d2h = total(sum(r.d2h_bytes for r in res_e2e)) / args.steps
    launches_all = total(launches)

    # N > 1: the same fixed problem on ONE GPU (rank 0 alone, the other ranks wait), so that the scaling of this workload
    # can be read from this line whatever workload the N = 1 run of the driver uses
    single = None
    if world > 1 and not args.no_single:
        if rank == 0:
            for _ in range(2):
                step(True, alone=True)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            rs = [step(True, alone=True) for _ in range(2)]
            e1.record()
            torch.cuda.synchronize()
            ms1 = e0.elapsed_time(e1)
            single = {"n_gpus": 1, "steps": 2, "warmup": 2, "ms_per_step": ms1 / 2,
                      "value": sum(r.n_evaluations for r in rs) / (ms1 * 1e-3), "unit": UNIT}
        barrier()

    if rank == 0:
        peak, peak_src = measured_peaks()
        dense_ms, dense_launches = prof["score_dense"]
        per_step_alg = dense_bytes
        alg_tput = per_step_alg * n_prof_steps / (dense_ms * 1e-3) / 1e9 if dense_ms > 0 else 0.0
        # counters exist for the workload they were captured on (C2: profiles/r02_dense_c2_metrics.csv); a line of another
        # workload carries no issue-rate fraction rather than one scaled from C2
        ctr = kernel_counters(f"{name}:dense_search") or {}
        sm_mhz = (clk.get("sm_mhz") or 1965.0)
        issue_peak = 148 * 4 * sm_mhz * 1e6 / 1e9  # G warp-instructions/s: 4 schedulers per SM, one instruction per cycle each
        avg_launch_ms = dense_ms / max(dense_launches, 1)
        inst = ctr.get("inst_executed")  # warp instructions of the dense launches of ONE step, from the committed ncu capture
        inst_rate = (inst * n_prof_steps / (dense_ms * 1e-3) / 1e9) if inst and dense_ms > 0 else None
        traffic = ctr.get("dram_bytes")
        line = {"metric": METRIC, "value": evals / (ms_value * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_value / args.steps, "higher_is_better": True,
                "scaling": args.scaling if world > 1 else "strong",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": workload_config(name, world, bool(args.nms), exchange_name, args.scaling),
                "nn_queries_per_sec": queries / (ms_value * 1e-3),
                "nn_queries_note": "object points of evaluated poses (the reference searches every one); most are answered by the "
                                   "block-cone prefilter without a search",
                "e2e": {"value": evals_e2e / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": ms_e2e / args.steps},
                "gpu_launches": int(launches_all),
                "roofline": {"bound": "issue", "kernel": "dense level-4 pose scoring (prefilter + bin + cell-staged search + reduce)",
                             "achieved": inst_rate, "peak": issue_peak, "unit": "G warp-inst/s",
                             "frac": (inst_rate / issue_peak) if inst_rate else None,
                             "traffic": traffic, "counters_source": ctr.get("source"),
                             "hbm": {"peak": peak, "peak_source": peak_src, "unit": "GB/s",
                                     "achieved_from_measured_traffic": (traffic * n_prof_steps / (dense_ms * 1e-3) / 1e9) if traffic and dense_ms > 0 else None,
                                     "frac": (traffic * n_prof_steps / (dense_ms * 1e-3) / 1e9 / peak) if traffic and dense_ms > 0 else None},
                             "algorithmic_throughput": alg_tput, "algorithmic_throughput_unit": "GB/s",
                             "algorithmic_bytes_per_step": per_step_alg, "census_stride": cstride,
                             "launches_timed": dense_launches, "timed_in": "a separate pass of the same steps with one object in flight (lanes=1)",
                             "avg_launch_ms": avg_launch_ms, "dense_ms_per_step": dense_ms / n_prof_steps,
                             "dense_nn_queries_per_sec": dense_queries * n_prof_steps / (dense_ms * 1e-3) if dense_ms > 0 else 0.0,
                             "note": "The dense search is not HBM-bound on this workload: its working set (grid + cone tables) is "
                                     "L2-resident, so the bound it is held against is the SM issue rate (warp instructions of one step "
                                     "from the committed ncu capture / measured time / (148 SMs x 4 schedulers x SM clock)); hbm.frac is the "
                                     "measured DRAM traffic against the copy peak.  algorithmic_throughput = what the reference's search "
                                     "reads for the same poses (8 B/cell + 16 B/point + 12 B/normal per query + 68 B/pose, SURVEY.md 8d, no "
                                     "early-out credit) per second: a work rate, not a bandwidth (the kernel prunes).  The HBM-bound regime "
                                     "of the same gather is `bench.py --workload C5`."},
                "kernel_ms_per_step": dict({k: v[0] / n_prof_steps for k, v in prof.items() if v[1]}, step_one_object_in_flight=ms_serial / n_prof_steps),
                "lanes": args.lanes if args.lanes is not None else pipeline.default_lanes(),
                "clocks": clk}
        if single is not None:
            line["single_gpu_same_workload"] = single
        if world == 1 and not args.no_cpu_baseline:
            rate, info, _ = cpu_reference_rate(scene, rotations, translations, target_seconds=15.0)
            line["cpu_baseline"] = dict(info, value=rate, unit=UNIT)
        if world == 1 and name == "C2" and not args.no_nn:
            # the HBM-regime kernel of the same path (C5: 10 M points, r = 0.10 m, k = 64, 4 M queries) measured in the same run;
            # `bench.py --workload C5` prints it as a line of its own (with its cpu_baseline)
            try:
                del dev, pinned, flush
                torch.cuda.empty_cache()
                nn_args = argparse.Namespace(**dict(vars(args), steps=3, warmup=3))
                nn = run_nn_workload(nn_args, rank, world, local_rank, device, dist, emit=False, with_clocks=False)
                line["nn_search_c5"] = {k: nn[k] for k in ("metric", "value", "unit", "ms_per_step", "steps", "warmup", "e2e", "gpu_launches", "roofline", "config")}
            except Exception as e:  # the pose line stands on its own
                line["nn_search_c5"] = {"unavailable": f"{type(e).__name__}: {e}"}
        print(json.dumps(line), flush=True)
    if peer is not None:
        peer.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
