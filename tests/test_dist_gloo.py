"""CPU, world_size 2 over gloo: the multi-GPU exchange of pipeline.run_step (variable-length all-gather of per-object
top-k lists + deterministic merge) gives every rank the same list, equal to the single-rank result."""
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
    props, ids = _fake_proposals(101, 36)
    lo, hi = pipeline.shard_range(len(props), rank, world)
    keep = props[lo:hi, 16] > 0.3  # each rank's survivors of its own translation block
    gp = pipeline._allgather_var(props[lo:hi][keep], dist, torch.device("cpu"))
    gi = pipeline._allgather_var(ids[lo:hi][keep], dist, torch.device("cpu"))
    mp_, mi = pipeline.merge_topk(gp, gi, 16)
    # empty contribution from one rank must also work
    ge = pipeline._allgather_var(props[:0] if rank == 0 else props[:3], dist, torch.device("cpu"))
    assert [len(x) for x in ge] == [0, 3]
    np.save(os.path.join(out, f"p{rank}.npy"), mp_)
    np.save(os.path.join(out, f"i{rank}.npy"), mi)
    dist.destroy_process_group()


def test_two_rank_allgather_merge(tmp_path):
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    p0, p1 = np.load(tmp_path / "p0.npy"), np.load(tmp_path / "p1.npy")
    i0, i1 = np.load(tmp_path / "i0.npy"), np.load(tmp_path / "i1.npy")
    assert (p0 == p1).all() and (i0 == i1).all()
    props, ids = _fake_proposals(101, 36)
    keep = props[:, 16] > 0.3
    sp, si = pipeline.merge_topk([props[keep]], [ids[keep]], 16)
    assert (sp == p0).all() and (si == i0).all()
