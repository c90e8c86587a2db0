"""CPU, world_size 2 over gloo: the multi-GPU exchange of pipeline.run_step (one all-gather of every object's
top-k list + deterministic merge, one all-gather of the refined rows) gives every rank the same list, equal to the single-rank result."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rescan_b200 import pipeline


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_proposals(n_trans, n_rot):
    rng = np.random.default_rng(5)
    scores = rng.uniform(0, 1, n_trans).astype(np.float32)
    scores[::7] = 0.5  # ties
    props = np.zeros((n_trans, 17), np.float32)
    props[:, 12] = np.arange(n_trans)
    props[:, 16] = scores
    ids = (np.arange(n_trans) * n_rot + rng.integers(0, n_rot, n_trans)).astype(np.int64)
    return props, ids


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    dev = torch.device("cpu")
    props, ids = _fake_proposals(101, 36)
    lo, hi = pipeline.shard_range(len(props), rank, world)
    keep = props[lo:hi, 16] > 0.3  # each rank's survivors of its own translation block
    # three "objects": this rank's survivors, an empty list on rank 0 only, an empty list everywhere
    mine_p = [props[lo:hi][keep], props[:0] if rank == 0 else props[:3], props[:0]]
    mine_i = [ids[lo:hi][keep], ids[:0] if rank == 0 else ids[:3], ids[:0]]
    for top_k in (16, 0):
        # rank-local top-k first (what rsgpu_propose_poses returns), then ONE all-gather for all objects
        loc = [pipeline.merge_topk([p], [i], top_k) for p, i in zip(mine_p, mine_i)]
        mp_, mi = pipeline.exchange_topk([l[0] for l in loc], [l[1] for l in loc], top_k, dist, dev)
        assert len(mp_[1]) == 3 and len(mp_[2]) == 0
        np.save(os.path.join(out, f"p{rank}_{top_k}.npy"), mp_[0])
        np.save(os.path.join(out, f"i{rank}_{top_k}.npy"), mi[0])
    # second exchange: rank r refined entries r::world of lists every rank holds
    n_list = [7, 0, 4, 1]
    full = [np.arange(n * 17, dtype=np.float32).reshape(n, 17) + 1000 * k for k, n in enumerate(n_list)]
    got = pipeline.exchange_rows([f[rank::world] for f in full], n_list, world, dist, dev)
    for g, f in zip(got, full):
        assert (g == f).all()
    dist.destroy_process_group()


def test_two_rank_allgather_merge(tmp_path):
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    props, ids = _fake_proposals(101, 36)
    keep = props[:, 16] > 0.3
    for top_k in (16, 0):
        p0, p1 = np.load(tmp_path / f"p0_{top_k}.npy"), np.load(tmp_path / f"p1_{top_k}.npy")
        i0, i1 = np.load(tmp_path / f"i0_{top_k}.npy"), np.load(tmp_path / f"i1_{top_k}.npy")
        assert (p0 == p1).all() and (i0 == i1).all()
        sp, si = pipeline.merge_topk([props[keep]], [ids[keep]], top_k)
        assert (sp == p0).all() and (si == i0).all()


def _step_worker(rank, world, port, out, lanes):
    """the whole pose-sharded run_step on CPU: rescan_b200.api replaced by tests/fake_api.py (input-determined fake kernels)"""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    res = _fake_step(rank, world, dist, lanes)
    np.savez(os.path.join(out, f"s{rank}_{lanes}.npz"), **{f"p{k}": p for k, p in enumerate(res.proposals)},
             **{f"i{k}": i for k, i in enumerate(res.pose_ids)}, n_eval=res.n_evaluations)
    dist.destroy_process_group()


def _fake_step(rank, world, dist_mod, lanes, peer_dir=None):
    from tests import fake_api
    pipeline.api = fake_api
    pipeline._POOLS.clear()
    rng = np.random.default_rng(11)
    models = [pipeline.ObjectModel(100 + k, 5, k == 2, {l: fake_api.PointCloud(np.zeros((10 * (k + 1) + l, 3)), None) for l in (1, 2, 3, 4)},
                                   np.zeros(3, np.float32)) for k in range(7)]
    trans = rng.uniform(0, 6, (501, 3)).astype(np.float32)
    rots = np.tile(np.eye(4, dtype=np.float32).reshape(1, 16), (9, 1))
    scan = (np.zeros((50, 3), np.float32), np.zeros((50, 3), np.float32))
    prev = [rng.uniform(0, 6, (1, 16)).astype(np.float32) for _ in range(6)]
    return pipeline.run_step(scan, scan, models, rots, trans, top_k=16, nms_dist=0.2, previous=prev, rank=rank, world=world,
                             dist=dist_mod, device=torch.device("cpu"), lanes=lanes,
                             peer=fake_api.FilePeer(peer_dir, rank, world) if peer_dir else None)


def _owner_worker(rank, world, out, lanes):
    """the owner-per-object schedule (the product's multi-GPU path: pipeline.run_step with a peer exchange) on CPU; no
    torch.distributed on the data path at all"""
    res = _fake_step(rank, world, None, lanes, peer_dir=os.path.join(out, f"peer_w{world}_l{lanes}"))
    np.savez(os.path.join(out, f"o{rank}_{world}_{lanes}.npz"), **{f"p{k}": p for k, p in enumerate(res.proposals)},
             **{f"i{k}": i for k, i in enumerate(res.pose_ids)}, n_eval=res.n_evaluations)


def test_owner_per_object_step_on_cpu(tmp_path):
    """pipeline.run_step over a peer exchange (gather of the top-k lists to the object's owner rank, owner refines, broadcast),
    world_size 2 and 3, serial and with lanes: every rank ends with the single-rank lists, and the refinement work is split"""
    try:
        ref = _fake_step(0, 1, None, 1)
        for world, lanes in ((2, 1), (2, 4), (3, 4)):
            mp.spawn(_owner_worker, args=(world, str(tmp_path), lanes), nprocs=world, join=True)
            evals = []
            for rank in range(world):
                z = np.load(tmp_path / f"o{rank}_{world}_{lanes}.npz")
                evals.append(int(z["n_eval"]))
                for k, (p, i) in enumerate(zip(ref.proposals, ref.pose_ids)):
                    assert z[f"p{k}"].shape == p.shape and (z[f"p{k}"] == p).all() and (z[f"i{k}"] == i).all(), (world, lanes, rank, k)
            assert sum(evals) == ref.n_evaluations and min(evals) > 0
    finally:
        import rescan_b200.api as real_api
        pipeline.api = real_api
        pipeline._POOLS.clear()


def test_pose_sharded_step_host_logic_on_cpu(tmp_path):
    """lanes, object groups, the fixed order of the collectives and both exchanges of pipeline.run_step, world_size 2 over gloo on
    the CPU: every rank must end with the single-rank lists"""
    import importlib
    try:
        ref = _fake_step(0, 1, None, 1)
        ref4 = _fake_step(0, 1, None, 4)
        for a, b, ia, ib in zip(ref.proposals, ref4.proposals, ref.pose_ids, ref4.pose_ids):
            assert (a == b).all() and (ia == ib).all()
        assert any(len(p) for p in ref.proposals)
        for lanes in (1, 4):
            world, port = 2, _free_port()
            mp.spawn(_step_worker, args=(world, port, str(tmp_path), lanes), nprocs=world, join=True)
            total = 0
            for rank in range(world):
                z = np.load(tmp_path / f"s{rank}_{lanes}.npz")
                total += int(z["n_eval"])
                for k, (p, i) in enumerate(zip(ref.proposals, ref.pose_ids)):
                    assert z[f"p{k}"].shape == p.shape and (z[f"p{k}"] == p).all() and (z[f"i{k}"] == i).all(), (lanes, rank, k)
            assert total == ref.n_evaluations
    finally:
        import rescan_b200.api as real_api
        pipeline.api = real_api
        pipeline._POOLS.clear()
