"""C4 — temporal rescan sequence (BASELINE.json configs[3]): 5 synthetic scans of one room, 40 objects; every scan runs
level building (the reference's Poisson-disk levels 1-4 of the new scan, on the GPU), pose_proposal (dense search +
verification + NMS + previous placements + ICP + rescoring + NMS, apps/pose_proposal/main.cpp:159-206) and then
segment_transfer's unary path (label transfer, data_cost, 8-NN edge weights) on the GPU.  The arrangement handed to the label transfer is the best refined proposal
of every dynamic object plus the static objects at their known poses (the reference's arrangement optimisation is
host code outside the path).  Prints one JSON line per scan and a summary.

    python scripts/run_sequence.py [--scans 5] [--objects 40] [--seeds 2048] [--rot 36] [--check]
--check compares labels / data_cost of every scan with the CPU oracle (slow: for tests and spot checks)."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rescan_b200 import api, pipeline, posegrid, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scans", type=int, default=5)
    ap.add_argument("--objects", type=int, default=40)
    ap.add_argument("--static", type=int, default=6)
    ap.add_argument("--seeds", type=int, default=2048)
    ap.add_argument("--rot", type=int, default=36)
    ap.add_argument("--room", default="12.0,2.6,9.0")
    ap.add_argument("--spacing", type=float, default=0.024)
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--host-levels", action="store_true", help="keep the generator's voxel-thinned stand-in levels instead of building them on the GPU")
    args = ap.parse_args()
    api.set_device(0)
    room = tuple(float(x) for x in args.room.split(","))
    scene = synth.make_scene(n_objects=args.objects, n_static=args.static, room=room, spacing=args.spacing, seed=synth.SEED + 40)
    rotations = posegrid.rotation_xforms(args.rot)
    models = pipeline.upload_objects(scene.objects)
    obj_grids = pipeline.upload_object_grids(scene.objects)
    is_static = [o.is_static for o in scene.objects]
    dyn_idx = [i for i, s in enumerate(is_static) if not s]
    totals = dict(evals=0, queries=0, ms_levels=0.0, ms_propose=0.0, ms_unary=0.0, vertices=0)
    previous = None  # best refined placement of every dynamic object in the previous scan (main.cpp:163-173 appends those)
    for s in range(args.scans):
        if s > 0:  # the next rescan: same objects, fresh poses (30 % unmoved)
            scene = synth.make_scene(n_objects=args.objects, n_static=args.static, room=room, spacing=args.spacing,
                                     seed=synth.SEED + 40 + s, objects=scene.objects)
        api.synchronize()
        tl = time.perf_counter()
        if not args.host_levels:  # rs_pointcloud_compute_levels of the new scan (rs_pointcloud.h:1305), sampling on the GPU
            lv = api.compute_levels(scene.scan.pos(0), scene.scan.nor(0))
            scene.scan = synth.make_cloud(scene.scan.pos(0), scene.scan.nor(0), levels=lv)
        translations = synth.translation_seeds(scene.scan, args.seeds, seed=synth.SEED + 100 + s)
        p1, n1, p2, n2 = scene.scan.pos(1), scene.scan.nor(1), scene.scan.pos(2), scene.scan.nor(2)
        api.synchronize()
        t0 = time.perf_counter()
        res = pipeline.run_step((p1, n1), (p2, n2), models, rotations, translations, top_k=64, nms_dist=0.2, previous=previous)
        api.synchronize()
        t1 = time.perf_counter()
        placements, recovered = [], 0
        previous = []
        for k, i in enumerate(dyn_idx):
            props = res.proposals[k]
            previous.append(props[:1, :16].copy() if len(props) and props[0, 16] > 0 else np.zeros((0, 16), np.float32))
            if len(props) and props[0, 16] > 0:
                placements.append((i, props[0, :16]))
                best = props[0, :16].reshape(4, 4).T
                recovered += int(np.linalg.norm(best[:3, 3] - scene.objects[i].pose[:3, 3]) < 0.05)
        for i, st in enumerate(is_static):
            if st:
                placements.append((i, np.ascontiguousarray(scene.objects[i].pose.T.reshape(16), np.float32)))
        g1 = api.HashGrid(p1, 0.05, normals=n1)
        t2 = time.perf_counter()
        un = pipeline.run_unary((p1, n1), g1, placements, obj_grids, is_static)
        api.synchronize()
        t3 = time.perf_counter()
        line = dict(scan=s, scan_points_lvl1=len(p1), objects=args.objects, dynamic=len(dyn_idx), evaluations=res.n_evaluations,
                    levels_ms=(t0 - tl) * 1e3, propose_ms=(t1 - t0) * 1e3, unary_ms=(t3 - t2) * 1e3, recovered_within_5cm=recovered,
                    labelled_vertices=int((un.labels > 0).sum()), labels=int(un.data_cost.shape[1]),
                    data_cost_bytes=int(un.data_cost.nbytes), edges=int((un.neighbors >= 0).sum()))
        if args.check:
            from oracle import orcbind as O
            order = un.placement_order
            poses = np.stack([np.asarray(placements[i][1], np.float32).reshape(16) for i in order])
            og = [O.OrcGrid(scene.objects[placements[i][0]].cloud.pos(1), 0.05) for i in order]
            on = [scene.objects[placements[i][0]].cloud.nor(1) for i in order]
            n_dyn = sum(not is_static[placements[i][0]] for i in order)
            lab, mind = np.zeros(len(p1), np.int8), np.full(len(p1), 1e9, np.float32)
            for first, last, r in ((0, n_dyn, 0.05), (n_dyn, len(order), 0.075)):
                O.assign_labels(p1, n1, poses, og, on, first, last, r, lab, mind)
            L = len(order) + 5
            st = np.zeros(L, np.uint8)
            for j, i in enumerate(order):
                st[j + 1] = 1 if is_static[placements[i][0]] else 0
            line["labels_match_oracle"] = bool((lab == un.labels).all())
            line["data_cost_match_oracle"] = bool((O.unary_costs(lab.astype(np.int32), st, L) == un.data_cost).all())
        g1.close()
        print(json.dumps(line), flush=True)
        totals["evals"] += res.n_evaluations
        totals["queries"] += res.n_queries
        totals["ms_levels"] += (t0 - tl) * 1e3
        totals["ms_propose"] += (t1 - t0) * 1e3
        totals["ms_unary"] += (t3 - t2) * 1e3
        totals["vertices"] += len(p1)
    print(json.dumps(dict(summary="C4", scans=args.scans, pose_evaluations_per_s=totals["evals"] / (totals["ms_propose"] * 1e-3),
                          unary_vertices_per_s=totals["vertices"] / (totals["ms_unary"] * 1e-3), **totals)), flush=True)


if __name__ == "__main__":
    main()
