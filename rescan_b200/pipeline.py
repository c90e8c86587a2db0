"""The pose_proposal hot path end to end on the GPU, as one "step" (used by bench.py, __graft_entry__.smoke()
and the pipeline tests).

One step = what reference apps/pose_proposal/main.cpp:118-206 does between loading and saving, minus the host
stages that stay reference code (NMS, sorting, .rsdb / .bin I/O):

  1. build the scan's level-1 grid (scoring) and level-2 grid (ICP)           rs_pointcloud.h:849-863
  2. per dynamic object: dense pose search at level 4 + verification at 3, 2   pose_proposal.cpp:325-369
  2b. non-maxima suppression of every object's proposals (``nms_dist``)        main.cpp:161, pose_proposal.cpp:371-452
  2c. the placements of previous arrangements join the list with score 10.0   main.cpp:163-173
  3. ICP refinement of the per-object survivors at level 2                    main.cpp:175-197
  4. rescoring of the refined poses at object level 1 with k = 32             main.cpp:199-201
  5. non-maxima suppression again, then descending-score order                main.cpp:205-206

Multi-GPU: translations are sharded over ranks in contiguous blocks (the per-translation arg-max over rotations
stays rank-local); the only exchange is one all-gather of the per-object top-k proposals, after which every
rank refines an interleaved share of the merged list and a second small all-gather returns the refined poses.
"""
from __future__ import annotations

import dataclasses
import numpy as np

from . import api, posegrid, synth


@dataclasses.dataclass
class ObjectModel:
    uidx: int
    class_idx: int
    is_static: bool
    levels: dict  # lvl -> api.PointCloud
    centroid: np.ndarray = None  # rs_pointcloud_centroid( shape, 0 ) (rs_pointcloud.h:1319-1339), what NMS measures distances between


@dataclasses.dataclass
class StepResult:
    proposals: list          # per dynamic object: float32 [n, 17] (xform + rescored score), descending score
    pose_ids: list           # per dynamic object: int64 [n] dense pose ids (t * n_rot + r)
    n_evaluations: int       # mgs_compute_object_alignment_score call-equivalents done by THIS rank
    n_queries: int           # object points searched by THIS rank
    h2d_bytes: int
    d2h_bytes: int


def upload_objects(objects, levels=(4, 3, 2, 1)):
    out = []
    for o in objects:
        out.append(ObjectModel(o.uidx, o.class_idx, o.is_static,
                               {l: api.PointCloud(o.cloud.pos(l), o.cloud.nor(l)) for l in levels},
                               posegrid.cloud_centroid(o.cloud.pos(0))))
    return out


def shard_range(n, rank, world):
    """contiguous block [lo, hi) of n items for `rank` of `world` (sizes differ by at most one)"""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def merge_topk(per_rank_props, per_rank_ids, top_k):
    """deterministic merge of per-rank proposal lists: descending score, ties by pose id (identical on every rank)"""
    props = np.concatenate(per_rank_props) if per_rank_props else np.zeros((0, api.POSE_FLOATS), np.float32)
    ids = np.concatenate(per_rank_ids) if per_rank_ids else np.zeros(0, np.int64)
    order = np.lexsort((ids, -props[:, 16].astype(np.float64)))
    if top_k > 0:
        order = order[:top_k]
    return props[order], ids[order]


def _allgather_bytes(buf, dist, device):
    """ONE all-gather of equally sized byte buffers (NCCL over NVLink on GPU, gloo on CPU) -> uint8 [world, nbytes]"""
    import torch
    world = dist.get_world_size()
    send = torch.from_numpy(np.ascontiguousarray(buf).view(np.uint8).reshape(-1)).to(device, non_blocking=True)
    recv = torch.empty((world, send.numel()), dtype=torch.uint8, device=device)
    dist.all_gather_into_tensor(recv.view(-1), send)
    return recv.cpu().numpy()


def exchange_topk(props_list, ids_list, top_k, dist, device):
    """The per-object top-k exchange of the pose-sharded search as ONE collective for all objects: every rank packs
    its lists into {counts int64 [O], ids int64 [O, cap], rows float32 [O, cap, 17]}, one all-gather moves them, and the
    same deterministic merge (descending score, ties by pose id) runs on every rank.  cap = top_k, or the largest list
    on any rank when top_k == 0 (one extra 8-byte all-gather)."""
    n_obj = len(props_list)
    cap = int(top_k)
    if cap <= 0:
        mine = np.array([max([len(p) for p in props_list], default=0)], np.int64)
        cap = max(int(_allgather_bytes(mine, dist, device).view(np.int64).max()), 1)
    counts = np.array([len(p) for p in props_list], np.int64)
    ids = np.zeros((n_obj, cap), np.int64)
    rows = np.zeros((n_obj, cap, api.POSE_FLOATS), np.float32)
    for k, (p, i) in enumerate(zip(props_list, ids_list)):
        ids[k, : len(i)] = i
        rows[k, : len(p)] = p
    buf = np.concatenate([counts.view(np.uint8), ids.reshape(-1).view(np.uint8), rows.reshape(-1).view(np.uint8)])
    got = _allgather_bytes(buf, dist, device)
    o1, o2 = 8 * n_obj, 8 * n_obj + 8 * n_obj * cap
    out_p, out_i = [], []
    for k in range(n_obj):
        gp, gi = [], []
        for r in range(got.shape[0]):
            c = int(got[r, :o1].view(np.int64)[k])
            gi.append(got[r, o1:o2].view(np.int64).reshape(n_obj, cap)[k, :c])
            gp.append(got[r, o2:].view(np.float32).reshape(n_obj, cap, api.POSE_FLOATS)[k, :c])
        mp_, mi = merge_topk(gp, gi, top_k)
        out_p.append(mp_.copy())
        out_i.append(mi.copy())
    return out_p, out_i


def exchange_rows(upd_list, n_list, world, dist, device):
    """Second exchange: object k's list has n_list[k] entries on every rank and rank r refined entries r, r + world, ...;
    ONE all-gather of the padded [sum_k ceil(n_k / world), 17] rows returns, per object, the rows in list order."""
    caps = [-(-n // world) for n in n_list]
    off = np.concatenate([[0], np.cumsum(caps)]).astype(np.int64)
    rows = np.zeros((max(int(off[-1]), 1), api.POSE_FLOATS), np.float32)
    for k, u in enumerate(upd_list):
        rows[off[k]: off[k] + len(u)] = u
    got = _allgather_bytes(rows, dist, device).view(np.float32).reshape(world, -1, api.POSE_FLOATS)
    out = []
    for k, n in enumerate(n_list):
        full = np.zeros((n, api.POSE_FLOATS), np.float32)
        for r in range(world):
            c = len(range(r, n, world))
            full[r::world] = got[r, off[k]: off[k] + c]
        out.append(full)
    return out


def run_step(scan_lvl1, scan_lvl2, models, rotations, translations, top_k=64, icp_max_dist=0.10,
             icp_max_angle=np.float32(np.deg2rad(60.0)), rank=0, world=1, dist=None, device=None,
             scan_dev=None, do_icp=True, nms_dist=None, previous=None):
    """scan_lvl1 / scan_lvl2: (pos, nor) host arrays of the scan levels; scan_dev: optional dict of device pointers
    {"p1","n1","p2","n2"} (+ sizes from the host arrays) to build the grids from HBM-resident data instead.
    nms_dist: centroid-distance threshold of the two NMS passes (the reference passes 0.2, main.cpp:161/205); None
    skips both.  previous: per dynamic object, float32 [n,16] placements of earlier arrangements, appended with
    score 10.0 and pose id -1 before the ICP (main.cpp:163-173)."""
    h2d = d2h = 0
    p1, n1 = scan_lvl1
    p2, n2 = scan_lvl2
    if scan_dev is None:
        g1 = api.HashGrid(p1, 0.05, normals=n1)
        g2 = api.HashGrid(p2, 0.05, normals=n2) if do_icp else None
        h2d += p1.nbytes + n1.nbytes + (p2.nbytes + n2.nbytes if do_icp else 0)
    else:
        g1 = api.HashGrid(device_ptr=scan_dev["p1"], n_pts=len(p1), radius=0.05)
        api._check(api.lib().rsgpu_grid_set_normals_dev(g1.h, scan_dev["n1"]))
        g2 = None
        if do_icp:
            g2 = api.HashGrid(device_ptr=scan_dev["p2"], n_pts=len(p2), radius=0.05)
            api._check(api.lib().rsgpu_grid_set_normals_dev(g2.h, scan_dev["n2"]))
    n_rot = len(rotations)
    lo, hi = shard_range(len(translations), rank, world)
    my_trans = np.ascontiguousarray(translations[lo:hi])
    h2d += rotations.nbytes + my_trans.nbytes
    n_eval = n_query = 0
    dyn = [m for m in models if not m.is_static]  # pose_proposal.cpp:198
    out_props, out_ids = [], []
    # ---- dense search + verification per object on this rank's block of translations
    for m in dyn:
        props, ids = api.propose_poses(m.levels[4], m.levels[3], m.levels[2], g1, rotations, my_trans, top_k=top_k)
        d2h += props.nbytes + ids.nbytes
        n_eval += n_rot * len(my_trans)
        n_query += n_rot * len(my_trans) * len(m.levels[4])
        out_props.append(props)
        out_ids.append(ids + lo * n_rot)
    # ---- the only exchange of the search: one all-gather of every object's top-k, identical merge on every rank
    if world > 1:
        out_props, out_ids = exchange_topk(out_props, out_ids, top_k, dist, device)
    # ---- NMS (every rank suppresses the same merged lists: no exchange), then the previous placements join
    if nms_dist is not None:
        for k, m in enumerate(dyn):
            if len(out_props[k]):
                keep = api.non_maxima_suppression(m.levels[3], m.levels[1], m.centroid, out_props[k], nms_dist)
                h2d += out_props[k].nbytes
                d2h += keep.nbytes
                out_props[k], out_ids[k] = out_props[k][keep], out_ids[k][keep]
    if previous is not None:
        for k, prev in enumerate(previous):
            prev = np.asarray(prev, np.float32).reshape(-1, 16)
            if len(prev):
                add = np.concatenate([prev, np.full((len(prev), 1), 10.0, np.float32)], axis=1)
                out_props[k] = np.concatenate([out_props[k], add])
                out_ids[k] = np.concatenate([out_ids[k], np.full(len(prev), -1, np.int64)])
    # ---- ICP refinement of every surviving proposal of every object in ONE launch, then rescoring
    if do_icp:
        # without NMS the -1 verification failures (pose_proposal.cpp:292) are not worth refining; with it the list is
        # what the reference's main refines: every survivor (main.cpp:175-204)
        cands = [np.arange(len(p)) if nms_dist is not None else np.nonzero(p[:, 16] > 0)[0] for p in out_props]
        shares = [c[rank::world] for c in cands]
        jobs = [(m, p[s, :16]) for m, p, s in zip(dyn, out_props, shares) if len(s)]
        refined = api.icp_align_multi([m.levels[2] for m, _ in jobs], g2, [t for _, t in jobs], icp_max_dist, icp_max_angle) if jobs else []
        ri = 0
        updates = []
        for m, mine in zip(dyn, shares):
            if len(mine):
                T, err, it = refined[ri]
                ri += 1
                sc = api.compute_object_alignment_scores(m.levels[1], g1, T, 32, 0.10)  # main.cpp:199
                h2d += 2 * T.nbytes
                d2h += T.nbytes + err.nbytes + it.nbytes + sc.nbytes
                n_eval += len(mine)
                n_query += len(mine) * len(m.levels[1])
                updates.append(np.concatenate([T, sc[:, None]], axis=1).astype(np.float32))
            else:
                updates.append(np.zeros((0, api.POSE_FLOATS), np.float32))
        if world > 1:  # second (and last) exchange: the refined rows of every object in one all-gather
            updates = exchange_rows(updates, [len(c) for c in cands], world, dist, device)
        for k, (m, c, upd) in enumerate(zip(dyn, cands, updates)):
            props, ids = out_props[k], out_ids[k]
            if len(c):
                props[c] = upd
            if nms_dist is not None and len(props):
                keep = api.non_maxima_suppression(m.levels[3], m.levels[1], m.centroid, props, nms_dist)
                props, ids = props[keep], ids[keep]
            order = np.lexsort((ids, -props[:, 16].astype(np.float64)))  # mgs_sort_poses: descending score
            out_props[k], out_ids[k] = props[order], ids[order]
    g1.close()
    if g2 is not None:
        g2.close()
    return StepResult(out_props, out_ids, n_eval, n_query, h2d, d2h)


def make_workload(name):
    """synthetic scene + pose grid of a named config (rescan_b200.synth.CONFIGS)"""
    cfg = synth.CONFIGS[name]
    scene = synth.make_scene(**cfg["scene"])
    from . import posegrid
    rotations = posegrid.rotation_xforms(cfg["n_rot"])
    translations = synth.translation_seeds(scene.scan, cfg["n_seeds"])
    return scene, rotations, translations


# ------------------------------------------------------------------------------------------------ label transfer / unary terms
@dataclasses.dataclass
class UnaryResult:
    labels: np.ndarray      # int8 [V]: 1 + index into the sorted placement list, 0 = unlabelled (rs_pointcloud_filters.cpp:738-778)
    placement_order: list   # indices into the caller's placement list, dynamic first / static last (:823-835)
    data_cost: np.ndarray   # int32 [V, L] (:926-939)
    neighbors: np.ndarray   # int32 [V, 8] candidate edges (-1 = none) and their weights (:674-722)
    weights: np.ndarray


def upload_object_grids(objects):
    """level-1 hash grids of the object models (what rspf__assign_temporary_labels searches), built once per database"""
    return [api.HashGrid(o.cloud.pos(1), 0.05, normals=o.cloud.nor(1)) for o in objects]


def run_unary(scan_lvl1, scan_grid, placements, object_grids, is_static, with_edges=True):
    """segment_transfer's unary path on the GPU for one scan: rspf_arrangement_to_labels (dynamic placements with
    r = 0.05, then static ones with r = 0.075, reference rs_pointcloud_filters.cpp:823-848), the data_cost block of
    rspf_smooth_labels (:926-939) and the 8-NN edge weights of rspf_compute_neighborhood (:674-722).
    placements: list of (object index, 4x4 column-major float32[16]); is_static: per object."""
    p1, n1 = scan_lvl1
    V = len(p1)
    order = sorted(range(len(placements)), key=lambda i: bool(is_static[placements[i][0]]))  # stable: dynamic first
    poses = np.stack([np.asarray(placements[i][1], np.float32).reshape(16) for i in order]) if order else np.zeros((0, 16), np.float32)
    grids = [object_grids[placements[i][0]] for i in order]
    n_dyn = sum(not is_static[placements[i][0]] for i in order)
    labels, min_d = np.zeros(V, np.int8), np.full(V, 1e9, np.float32)
    if n_dyn == len(order):
        # no static placement: first_static_obj_idx stays 0, pass 1 is empty and everything is matched with r = 0.075 (:830-835)
        passes = ((0, len(order), 0.075),)
    else:
        passes = ((0, n_dyn, 0.05), (n_dyn, len(order), 0.075))
    for first, last, r in passes:
        if last > first:
            api.assign_labels(p1, n1, poses, grids, first, last, r, labels, min_d)
    lab32 = labels.astype(np.int32)
    L = len(order) + 5
    static_flag = np.zeros(L, np.uint8)
    for j, i in enumerate(order):
        static_flag[j + 1] = 1 if is_static[placements[i][0]] else 0
    cost = api.unary_costs(lab32, static_flag, L)
    nbr = wgt = None
    if with_edges:
        nbr, wgt = api.neighborhood(scan_grid, p1, n1)
    return UnaryResult(labels, order, cost, nbr, wgt)
