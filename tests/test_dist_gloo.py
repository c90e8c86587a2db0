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
