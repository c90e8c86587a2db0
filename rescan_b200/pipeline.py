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
import os
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
    trace: list = None       # with trace=True: (object index, stage, t_begin, t_end) host times in seconds since the step began


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


SHARD_BLOCK = 256  # translations per block of the Z-order curve dealt to one rank


def shard_translations(translations, rank, world, spatial=True):
    """indices (caller's numbering) of the translations `rank` of `world` searches, in the order it hands them to the search:
    the Z-order curve over (x, z) cut into blocks of SHARD_BLOCK, block b to rank b % world (spatial=False: the caller's order
    cut the same way).  Every translation belongs to exactly one rank; world = 1 gives the whole curve."""
    n = len(translations)
    order = posegrid.spatial_order(translations) if spatial else np.arange(n, dtype=np.int64)
    if world <= 1:
        return np.ascontiguousarray(order, np.int64)
    block = np.arange(n, dtype=np.int64) // SHARD_BLOCK
    return np.ascontiguousarray(order[block % world == rank], np.int64)


def merge_topk(per_rank_props, per_rank_ids, top_k):
    """deterministic merge of per-rank proposal lists (identical on every rank), in the order a single rank's
    rsgpu_propose_poses returns the same list: top_k > 0 descending score with ties by pose id; top_k <= 0 (the reference's
    behaviour: every survivor) ascending pose id = emission order, so the NMS that follows breaks ties between equal scores
    the same way for any number of ranks (pose_proposal.cpp:348-359, 404-411)"""
    props = np.concatenate(per_rank_props) if per_rank_props else np.zeros((0, api.POSE_FLOATS), np.float32)
    ids = np.concatenate(per_rank_ids) if per_rank_ids else np.zeros(0, np.int64)
    if top_k > 0:
        order = np.lexsort((ids, -props[:, 16].astype(np.float64)))[:top_k]
    else:
        order = np.argsort(ids, kind="stable")
    return props[order], ids[order]


def _allgather_bytes(buf, dist, device, group=None):
    """ONE all-gather of equally sized byte buffers -> uint8 [world, nbytes].  group = a peerx.PeerExchange: peer-mapped
    slots written over NVLink by copy engines (the default of bench.py); device = a CUDA device: staged through HBM and moved
    by NCCL; device = cpu (with a gloo `group`): the host-resident bytes never touch the GPU"""
    if hasattr(group, "allgather"):
        return group.allgather(buf)
    import torch
    world = dist.get_world_size()
    send = torch.from_numpy(np.ascontiguousarray(buf).view(np.uint8).reshape(-1)).to(device, non_blocking=True)
    recv = torch.empty((world, send.numel()), dtype=torch.uint8, device=device)
    dist.all_gather_into_tensor(recv.view(-1), send, group=group)
    return recv.cpu().numpy()


def exchange_topk(props_list, ids_list, top_k, dist, device, group=None):
    """The per-object top-k exchange of the pose-sharded search as ONE collective for all objects: every rank packs
    its lists into {counts int64 [O], ids int64 [O, cap], rows float32 [O, cap, 17]}, one all-gather moves them, and the
    same deterministic merge (descending score, ties by pose id) runs on every rank.  cap = top_k, or the largest list
    on any rank when top_k == 0 (one extra 8-byte all-gather)."""
    n_obj = len(props_list)
    cap = int(top_k)
    if cap <= 0:
        mine = np.array([max([len(p) for p in props_list], default=0)], np.int64)
        cap = max(int(_allgather_bytes(mine, dist, device, group).view(np.int64).max()), 1)
    counts = np.array([len(p) for p in props_list], np.int64)
    ids = np.zeros((n_obj, cap), np.int64)
    rows = np.zeros((n_obj, cap, api.POSE_FLOATS), np.float32)
    for k, (p, i) in enumerate(zip(props_list, ids_list)):
        ids[k, : len(i)] = i
        rows[k, : len(p)] = p
    buf = np.concatenate([counts.view(np.uint8), ids.reshape(-1).view(np.uint8), rows.reshape(-1).view(np.uint8)])
    got = _allgather_bytes(buf, dist, device, group)
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


def exchange_rows(upd_list, n_list, world, dist, device, group=None):
    """Second exchange: object k's list has n_list[k] entries on every rank and rank r refined entries r, r + world, ...;
    ONE all-gather of the padded [sum_k ceil(n_k / world), 17] rows returns, per object, the rows in list order."""
    caps = [-(-n // world) for n in n_list]
    off = np.concatenate([[0], np.cumsum(caps)]).astype(np.int64)
    rows = np.zeros((max(int(off[-1]), 1), api.POSE_FLOATS), np.float32)
    for k, u in enumerate(upd_list):
        rows[off[k]: off[k] + len(u)] = u
    got = _allgather_bytes(rows, dist, device, group).view(np.float32).reshape(world, -1, api.POSE_FLOATS)
    out = []
    for k, n in enumerate(n_list):
        full = np.zeros((n, api.POSE_FLOATS), np.float32)
        for r in range(world):
            c = len(range(r, n, world))
            full[r::world] = got[r, off[k]: off[k] + c]
        out.append(full)
    return out


_POOLS = {}


def lane_pool(n_lanes):
    """persistent host threads, each bound to its own rsgpu lane (stream): per-object call chains submitted to the
    pool overlap on the device (ctypes releases the GIL for the duration of every library call)"""
    import concurrent.futures
    import itertools
    import threading
    n_lanes = max(1, min(int(n_lanes), api.lane_count()))
    if n_lanes not in _POOLS:
        counter, lock = itertools.count(), threading.Lock()

        def attach():
            with lock:
                lane = next(counter)
            api.thread_attach(lane)
        _POOLS[n_lanes] = concurrent.futures.ThreadPoolExecutor(max_workers=n_lanes, initializer=attach)
    return _POOLS[n_lanes]


def _run(pool, fn, items, order=None):
    """fn(*item) for every item, results in item order; with a pool the items are SUBMITTED in `order` (the exposed tail
    of a step is the chain that finishes last, so the cheapest chain should start last)"""
    if pool is None:
        return [fn(*it) for it in items]
    order = range(len(items)) if order is None else order
    futs = {i: pool.submit(fn, *items[i]) for i in order}
    return [futs[i].result() for i in range(len(items))]


def default_lanes():
    import os
    return int(os.environ.get("RSGPU_LANES", "8"))


def configure_host_waits(lanes, local_world=None):
    """every lane thread of every rank on this host can be waiting on its stream at the same time; when they outnumber the
    cores, all waits sleep on blocking events instead of spinning (csrc/runtime.cu stream_sync)"""
    import os
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", "1")) if local_world is None else local_world
    oversubscribed = (lanes + 2) * local_world > (os.cpu_count() or 1)
    api.set_option("sync", "block" if oversubscribed else None)
    return oversubscribed


def run_step(scan_lvl1, scan_lvl2, models, rotations, translations, top_k=64, icp_max_dist=0.10,
             icp_max_angle=np.float32(np.deg2rad(60.0)), rank=0, world=1, dist=None, device=None,
             scan_dev=None, do_icp=True, nms_dist=None, previous=None, lanes=None, trace=False, host_group=None, peer=None):
    """top_k: per-object cap on the proposals that leave the search (descending score, ties by pose id).  The reference keeps
    every survivor (pose_proposal.cpp:348-359) = top_k 0; the default 64 is BASELINE.json's C3 configuration ("top-k = 64 per
    object") and a deliberate deviation for the bench workloads - pass top_k=0 for the reference's behaviour.
    scan_lvl1 / scan_lvl2: (pos, nor) host arrays of the scan levels; scan_dev: optional dict of device pointers
    {"p1","n1","p2","n2"} (+ sizes from the host arrays) to build the grids from HBM-resident data instead.
    nms_dist: centroid-distance threshold of the two NMS passes (the reference passes 0.2, main.cpp:161/205); None
    skips both.  previous: per dynamic object, float32 [n,16] placements of earlier arrangements, appended with
    score 10.0 and pose id -1 before the ICP (main.cpp:163-173).
    host_group: optional gloo process group: the two exchanges then run over the host.  The lists are host-resident
    (they come back through the C ABI) and a few KB; an NCCL kernel issued while the dense search saturates the GPU is
    not dispatched before the pending dense blocks drain (measured 14 ms at N = 2), a host all-gather of
    the same bytes takes 0.3 ms.  Without it the exchanges use `dist` / `device` as given (NCCL staged through HBM).
    peer: optional peerx.PeerExchange: both exchanges then go over NVLink through peer-mapped slots (copy engines, no
    collective kernel); takes precedence over host_group.
    lanes: how many objects are in flight at once (default RSGPU_LANES or 8; 1 = the reference's serial object loop).
    The objects are independent (pose_proposal.cpp:190-250, main.cpp:175-204), so every object's chain runs on its own
    lane: the latency-bound stages of one object (verification, NMS rounds, ICP iterations) fill the device next to the
    dense search of another.  Results do not depend on the number of lanes."""
    import time
    stats = dict(h2d=0, d2h=0, n_eval=0, n_query=0)
    t_step, events = time.perf_counter(), []

    def staged(k, stage, fn, *a):
        if not trace:
            return fn(*a)
        t0 = time.perf_counter()
        out = fn(*a)
        events.append((k, stage, t0 - t_step, time.perf_counter() - t_step))
        return out
    p1, n1 = scan_lvl1
    p2, n2 = scan_lvl2
    if peer is not None:
        host_group = peer
    t_g = time.perf_counter()
    import threading
    if scan_dev is None:
        g1 = api.HashGrid(p1, 0.05, normals=n1)
        stats["h2d"] += p1.nbytes + n1.nbytes + (p2.nbytes + n2.nbytes if do_icp else 0)
    else:
        g1 = api.HashGrid(device_ptr=scan_dev["p1"], n_pts=len(p1), radius=0.05)
        api._check(api.lib().rsgpu_grid_set_normals_dev(g1.h, scan_dev["n1"]))
    g2_box, g2_lock = [None], threading.Lock()

    def scan_grid_lvl2():
        """the level-2 grid the ICP searches: built by the first chain that needs it, on its lane, next to the dense launches
        of the other objects (it is not needed before the first refinement)"""
        with g2_lock:
            if g2_box[0] is None:
                if scan_dev is None:
                    g2_box[0] = api.HashGrid(p2, 0.05, normals=n2)
                else:
                    g = api.HashGrid(device_ptr=scan_dev["p2"], n_pts=len(p2), radius=0.05)
                    api._check(api.lib().rsgpu_grid_set_normals_dev(g.h, scan_dev["n2"]))
                    g2_box[0] = g
            return g2_box[0]
    if trace:
        events.append((-1, "grids", t_g - t_step, time.perf_counter() - t_step))
    n_rot = len(rotations)
    # The translations are handed to the dense search along a Z-order curve (neighbouring poses search neighbouring scan cells;
    # the cell-binned search lives on many queries per staged block of cells), and a rank's share is every world-th BLOCK of
    # that curve: compact pieces of the scene, so the query density per cell is what it is on one GPU (contiguous shares of
    # the caller's - unordered - list spread a rank's poses over the whole scan: 1.8 x the search time per pose at N = 8),
    # dealt round the ranks, so the shares cost the same.  translation_ids carries the caller's numbering: ids, order and
    # every tie are those of the caller's own order.
    my_ids = shard_translations(translations, rank, world, spatial=os.environ.get("RSGPU_SPATIAL_ORDER", "1") != "0")
    my_trans = walk_trans = np.ascontiguousarray(translations[my_ids])
    walk = my_ids
    stats["h2d"] += rotations.nbytes + my_trans.nbytes
    dyn = [m for m in models if not m.is_static]  # pose_proposal.cpp:198
    lanes = default_lanes() if lanes is None else lanes
    pool = lane_pool(lanes) if lanes > 1 and len(dyn) > 1 else None
    configure_host_waits(lanes if pool is not None else 1)
    nms = nms_dist is not None
    big_first = sorted(range(len(dyn)), key=lambda i: -len(dyn[i].levels[2]))  # ICP cost grows with the level-2 size

    stats_lock = threading.Lock()

    def add(**kw):  # called from the lane threads; the totals are order-free
        with stats_lock:
            for k, v in kw.items():
                stats[k] += int(v)

    def search(m):
        """dense search + verification on this rank's block of translations"""
        props, ids = api.propose_poses(m.levels[4], m.levels[3], m.levels[2], g1, rotations, walk_trans, top_k=top_k, translation_ids=walk)
        add(d2h=props.nbytes + ids.nbytes, n_eval=n_rot * len(my_trans), n_query=n_rot * len(my_trans) * len(m.levels[4]))
        return props, ids

    def suppress(m, props, ids):
        if nms and len(props):
            keep = api.non_maxima_suppression(m.levels[3], m.levels[1], m.centroid, props, nms_dist)
            add(h2d=props.nbytes, d2h=keep.nbytes)
            return props[keep], ids[keep]
        return props, ids

    def candidates(k, m, props, ids):
        """first NMS, the previous placements, and which entries get refined"""
        props, ids = suppress(m, props, ids)
        if previous is not None:
            prev = np.asarray(previous[k], np.float32).reshape(-1, 16)
            if len(prev):
                props = np.concatenate([props, np.concatenate([prev, np.full((len(prev), 1), 10.0, np.float32)], axis=1)])
                ids = np.concatenate([ids, np.full(len(prev), -1, np.int64)])
        # without NMS the -1 verification failures (pose_proposal.cpp:292) are not worth refining; with it the list is
        # what the reference's main refines: every survivor (main.cpp:175-204)
        cand = np.arange(len(props)) if nms else np.nonzero(props[:, 16] > 0)[0]
        return props, ids, cand

    def refine(m, props, mine):
        """ICP at level 2 + rescoring at level 1 with k = 32 of the entries `mine` (main.cpp:195-201) -> rows [len(mine), 17]"""
        if not len(mine):
            return np.zeros((0, api.POSE_FLOATS), np.float32)
        T, err, it = api.icp_align(m.levels[2], scan_grid_lvl2(), props[mine, :16], icp_max_dist, icp_max_angle)
        sc = api.compute_object_alignment_scores(m.levels[1], g1, T, 32, 0.10)
        add(h2d=2 * T.nbytes, d2h=T.nbytes + err.nbytes + it.nbytes + sc.nbytes, n_eval=len(mine), n_query=len(mine) * len(m.levels[1]))
        return np.concatenate([T, sc[:, None]], axis=1).astype(np.float32)

    def finish(m, props, ids, cand, upd):
        if len(cand):
            props[cand] = upd
        props, ids = suppress(m, props, ids)
        order = np.lexsort((ids, -props[:, 16].astype(np.float64)))  # mgs_sort_poses: descending score
        return props[order], ids[order]

    if world == 1:
        def chain(k, m):
            props, ids = staged(k, "search", search, m)
            if not do_icp:
                return suppress(m, props, ids) if nms else (props, ids)
            props, ids, cand = staged(k, "nms1", candidates, k, m, props, ids)
            upd = staged(k, "refine", refine, m, props, cand)
            return staged(k, "nms2", finish, m, props, ids, cand, upd)
        res = _run(pool, chain, list(enumerate(dyn)), big_first)
        out_props, out_ids = [r[0] for r in res], [r[1] for r in res]
    elif hasattr(host_group, "slot"):
        # Pose-sharded over peer-mapped slots (peerx.PeerExchange).  Every object has an OWNER rank: all ranks search their
        # block of translations for it and put their top-k list into the owner's area (a gather: nobody but the owner
        # waits); the owner merges, suppresses, refines ALL survivors, rescoring and second NMS included, and puts the final
        # list into every rank's area, where it is picked up at the end of the step.  The chain behind the search is
        # latency-bound (a few dozen dependent ICP iterations, greedy NMS rounds), so refining 1/N of every object's
        # candidates on every rank - the first version - left each rank with ALL the chains; with owners a rank runs 1/N of
        # them, and the NMS is no longer replicated.  Exchanges use the object's own slot, so nothing orders the objects
        # globally.  (No deadlock: every rank starts the chains in the same order, only owners ever wait, and the earliest
        # object some rank has not searched yet is never behind a blocked lane of that rank - everything ahead of it has
        # been searched by all ranks, so the gathers its lanes wait for complete.)
        peer = host_group
        if len(dyn) > peer.n_slots:
            raise ValueError(f"peer exchange has {peer.n_slots} slots for {len(dyn)} dynamic objects")
        owner = [0] * len(dyn)
        for pos, k in enumerate(big_first):
            owner[k] = pos % world  # the big objects (long ICP chains) dealt round the ranks
        everyone = list(range(world))
        uses = {k: (peer.begin_use(k), peer.begin_use(k)) for k in range(len(dyn))}  # (gather, broadcast) of this step

        def pack(props, ids):
            return np.concatenate([np.asarray(ids, np.int64).view(np.uint8), np.ascontiguousarray(props, np.float32).reshape(-1).view(np.uint8)])

        def unpack(buf):
            n = len(buf) // (8 + 4 * api.POSE_FLOATS)
            return buf[8 * n:].view(np.float32).reshape(n, api.POSE_FLOATS).copy(), buf[: 8 * n].view(np.int64).copy()

        def chain(k, m):
            use_gather, use_bcast = uses[k]
            props, ids = staged(k, "search", search, m)
            staged(k, "put_topk", peer.put, k, use_gather, [owner[k]], pack(props, ids))
            if owner[k] != rank:
                return None
            got = staged(k, "gather", peer.get, k, use_gather, everyone)
            lists = [unpack(got[r]) for r in everyone]
            props, ids = merge_topk([l[0] for l in lists], [l[1] for l in lists], top_k)
            if not do_icp:
                props, ids = suppress(m, props, ids)
            else:
                props, ids, cand = staged(k, "nms1", candidates, k, m, props, ids)
                upd = staged(k, "refine", refine, m, props, cand)
                props, ids = staged(k, "nms2", finish, m, props, ids, cand, upd)
            staged(k, "bcast", peer.put, k, use_bcast, [r for r in everyone if r != rank], pack(props, ids))
            return props, ids
        res = _run(pool, chain, list(enumerate(dyn)), big_first)
        out_props, out_ids = [None] * len(dyn), [None] * len(dyn)
        for k, r in enumerate(res):
            if r is None:  # another rank's object: its final list has been (or is being) put into this rank's area
                r = unpack(staged(k, "collect", peer.get, k, uses[k][1], [owner[k]])[owner[k]])
            out_props[k], out_ids[k] = r
    else:
        # Pose-sharded with a COLLECTIVE transport (gloo / NCCL): the per-translation arg-max and the verification are
        # rank-local, the per-object top-k lists are merged across ranks.  The objects go through in groups so that the refinement of one group overlaps the dense
        # search of the next: every collective is issued by THIS thread in a fixed order (top-k of group 0, 1, ..., then
        # the refined rows of group 0, 1, ...), identical on every rank whatever the timing of the lanes.
        n_groups = min(int(os.environ.get("RSGPU_GROUPS", "4")), max(1, len(dyn)))
        groups = [big_first[j::n_groups] for j in range(n_groups)]  # big objects spread over the groups, in submission order
        groups = [g for g in groups if g]
        order = [k for g in groups for k in g]
        out_props, out_ids = [None] * len(dyn), [None] * len(dyn)
        if pool is not None:
            sf = {k: pool.submit(staged, k, "search", search, dyn[k]) for k in order}
            get_search = lambda k: sf[k].result()
        else:
            get_search = lambda k: staged(k, "search", search, dyn[k])

        def middle_(k, props, ids):
            return staged(k, "middle", middle, k, props, ids)

        def finish_(k, *a):
            return staged(k, "finish", finish, *a)

        def middle(k, props, ids):  # every rank suppresses the same merged list (no exchange) and refines its interleaved share
            props, ids, cand = candidates(k, dyn[k], props, ids)
            return props, ids, cand, refine(dyn[k], props, cand[rank::world])
        mids = []
        for g in groups:
            res = [get_search(k) for k in g]
            # one all-gather for the group's top-k lists, identical merge on every rank
            gp, gi = staged(-2, "xchg_topk", exchange_topk, [r[0] for r in res], [r[1] for r in res], top_k, dist,
                            device if host_group is None else "cpu", host_group)
            if not do_icp:
                mids.append([(pool.submit(suppress, dyn[k], p, i) if pool is not None else suppress(dyn[k], p, i)) for k, p, i in zip(g, gp, gi)])
            elif pool is not None:
                mids.append([pool.submit(middle_, k, p, i) for k, p, i in zip(g, gp, gi)])
            else:
                mids.append([middle_(k, p, i) for k, p, i in zip(g, gp, gi)])
        fins = []
        for g, mf in zip(groups, mids):
            mid = [f.result() if pool is not None else f for f in mf]
            if not do_icp:
                fins.append(mid)
                continue
            # the refined rows of the group's objects in one all-gather
            updates = staged(-2, "xchg_rows", exchange_rows, [r[3] for r in mid], [len(r[2]) for r in mid], world, dist,
                             device if host_group is None else "cpu", host_group)
            args = [(k, dyn[k], r[0], r[1], r[2], u) for k, r, u in zip(g, mid, updates)]
            fins.append([pool.submit(finish_, *a) for a in args] if pool is not None else [finish_(*a) for a in args])
        for g, ff in zip(groups, fins):
            for k, f in zip(g, ff):
                r = f.result() if (pool is not None and do_icp) else f
                out_props[k], out_ids[k] = r[0], r[1]
    g1.close()
    if g2_box[0] is not None:
        g2_box[0].close()
    if trace:
        events.append((-1, "step", 0.0, time.perf_counter() - t_step))
    return StepResult(out_props, out_ids, stats["n_eval"], stats["n_query"], stats["h2d"], stats["d2h"], sorted(events, key=lambda e: e[2]) if trace else None)


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
