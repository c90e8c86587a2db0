"""The whole pose_proposal step of reference apps/pose_proposal/main.cpp:159-206 on the GPU (propose -> NMS -> previous
placements appended with score 10 -> ICP -> rescore at level 1 with k = 32 -> NMS -> descending score) against the same
sequence composed from the CPU oracle's routines."""
import numpy as np
import pytest

from oracle import orcbind as O
from rescan_b200 import api, pipeline, posegrid, synth
from tests import common

pytestmark = pytest.mark.gpu


def _oracle_step(scene, rots_ang, trans, previous):
    ang = rots_ang
    og1 = O.OrcGrid(scene.scan.pos(1), 0.05)
    n1 = scene.scan.nor(1)
    xf = O.make_pose_grid(ang, trans)
    out = []
    dyn = [o for o in scene.objects if not o.is_static]
    for k, o in enumerate(dyn):
        ref, _ = O.score_poses(o.cloud.pos(4), o.cloud.nor(4), og1, n1, xf, 64, 0.10, n_threads=8)
        ot, orr, osc = O.select_proposals(ref.reshape(len(trans), len(ang)), 0.25)
        props = []
        for t, r, s in zip(ot, orr, osc):
            x = xf[t, r]
            for lvl, thr in ((3, 0.35), (2, 0.40)):
                if s > 0:
                    v, _ = O.score_poses(o.cloud.pos(lvl), o.cloud.nor(lvl), og1, n1, x[None], 64, 0.10)
                    s = v[0] if v[0] > thr else -1.0
            props.append(np.concatenate([x.reshape(16), [s]]))
        props = np.asarray(props, np.float32).reshape(-1, 17)
        cen = O.centroid(o.cloud.pos(0))
        if len(props):
            props = props[O.nms(o.cloud.pos(3), o.cloud.pos(1), cen, props, 0.2)]
        prev = np.asarray(previous[k], np.float32).reshape(-1, 16)
        props = np.concatenate([props, np.concatenate([prev, np.full((len(prev), 1), 10.0, np.float32)], axis=1)])
        for p in props:
            T, err, it = O.icp_align(o.cloud.pos(2), o.cloud.nor(2), scene.scan.pos(2), scene.scan.nor(2), p[:16], 0.10,
                                     np.float32(np.deg2rad(60.0)))
            v, _ = O.score_poses(o.cloud.pos(1), o.cloud.nor(1), og1, n1, np.asarray(T, np.float32).reshape(1, 16), 32, 0.10)
            p[:16] = np.asarray(T, np.float32).reshape(16)
            p[16] = v[0]
        if len(props):
            props = props[O.nms(o.cloud.pos(3), o.cloud.pos(1), cen, props, 0.2)]
        out.append(props[np.argsort(-props[:, 16].astype(np.float64), kind="stable")])
    return out


def test_full_step_with_nms_matches_oracle_sequence():
    scene = common.tiny_scene()
    rots, ang = common.rotation_xforms(8)
    trans = synth.translation_seeds(scene.scan, 64, seed=9)
    dyn = [o for o in scene.objects if not o.is_static]
    for i, o in enumerate(scene.objects):
        trans[i] = [o.pose[0, 3], 0.0, o.pose[2, 3]]
    rng = np.random.default_rng(4)
    # one "previous arrangement" placement per dynamic object: the true pose, slightly off
    previous = []
    for o in dyn:
        d = synth.yaw_pose(rng.uniform(-0.05, 0.05), rng.uniform(-0.03, 0.03), rng.uniform(-0.03, 0.03), 0.0)
        previous.append(common.colmajor((d.astype(np.float64) @ o.pose.astype(np.float64)).astype(np.float32))[None])
    models = pipeline.upload_objects(scene.objects)
    for m, o in zip(models, scene.objects):
        assert (m.centroid == O.centroid(o.cloud.pos(0))).all()
    res = pipeline.run_step((scene.scan.pos(1), scene.scan.nor(1)), (scene.scan.pos(2), scene.scan.nor(2)), models, rots, trans,
                            top_k=0, nms_dist=0.2, previous=previous)
    want = _oracle_step(scene, ang, trans, previous)
    assert len(res.proposals) == len(want)
    n_checked = 0
    for got, w in zip(res.proposals, want):
        assert len(got) == len(w)
        # scores within the stated tolerance (1e-4 relative), poses within 1e-5 m / 1e-5 rad
        assert np.allclose(got[:, 16], w[:, 16], rtol=1e-4, atol=1e-7)
        for a, b in zip(got, w):
            A, B = a[:16].reshape(4, 4).T.astype(np.float64), b[:16].reshape(4, 4).T.astype(np.float64)
            assert np.linalg.norm(A[:3, 3] - B[:3, 3]) < 1e-5
            R = A[:3, :3] @ B[:3, :3].T
            assert np.linalg.norm(R - R.T) / (2 * np.sqrt(2)) < 1e-5
            n_checked += 1
    assert n_checked >= len(dyn)  # at least the previous placements came through


def test_step_without_nms_is_a_superset():
    """NMS only removes proposals: every pose id that survives the NMS step is in the plain top-k step"""
    scene = common.tiny_scene()
    rots, _ = common.rotation_xforms(8)
    trans = synth.translation_seeds(scene.scan, 64, seed=9)
    models = pipeline.upload_objects(scene.objects)
    scan1, scan2 = (scene.scan.pos(1), scene.scan.nor(1)), (scene.scan.pos(2), scene.scan.nor(2))
    a = pipeline.run_step(scan1, scan2, models, rots, trans, top_k=16)
    b = pipeline.run_step(scan1, scan2, models, rots, trans, top_k=16, nms_dist=0.2)
    for ia, ib in zip(a.pose_ids, b.pose_ids):
        assert set(ib.tolist()) <= set(ia.tolist())


def test_lanes_do_not_change_results():
    """per-object chains on concurrent lanes (streams) give bit-identical lists to the serial object loop"""
    scene = common.small_scene()
    rots, _ = common.rotation_xforms(12)
    trans = synth.translation_seeds(scene.scan, 192, seed=3)
    for i, o in enumerate(scene.objects):
        trans[i] = [o.pose[0, 3], 0.0, o.pose[2, 3]]
    models = pipeline.upload_objects(scene.objects)
    scan1, scan2 = (scene.scan.pos(1), scene.scan.nor(1)), (scene.scan.pos(2), scene.scan.nor(2))
    ref = pipeline.run_step(scan1, scan2, models, rots, trans, top_k=16, nms_dist=0.2, lanes=1)
    for lanes in (2, 4, 8):
        for _ in range(2):
            got = pipeline.run_step(scan1, scan2, models, rots, trans, top_k=16, nms_dist=0.2, lanes=lanes)
            assert got.n_evaluations == ref.n_evaluations and got.n_queries == ref.n_queries
            for a, b, ia, ib in zip(got.proposals, ref.proposals, got.pose_ids, ref.pose_ids):
                assert a.shape == b.shape and (a == b).all() and (ia == ib).all()


def _rank_worker(rank, world, port, out_dir, lanes, exchange="gloo", top_k=16):
    import os
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)  # both ranks share cuda:0
    api.set_device(0)
    peer = None
    if exchange == "peer":
        # the product's transport (bench.py default): peer-mapped slots + copy engines (csrc/peer.cu); with both ranks on one
        # GPU the peer copies are local, the protocol (IPC mapping, ring, flags, wait kernel) is the same
        from rescan_b200 import peerx
        peer = peerx.PeerExchange(dist, rank, world, slot_bytes=1 << 20, timeout_s=20.0)
        got = peer.allgather(np.full(1000, rank + 1, np.uint8))
        assert got.shape == (world, 1000) and all((got[r] == r + 1).all() for r in range(world))
        for n in (0, 1, 7, 4096, 1 << 20):  # ring wrap-around, empty and full-slot payloads
            got = peer.allgather((np.arange(n) * (rank + 3) % 251).astype(np.uint8))
            assert all((got[r] == (np.arange(n) * (r + 3) % 251).astype(np.uint8)).all() for r in range(world))
    scene = common.small_scene()
    rots, _ = common.rotation_xforms(12)
    trans = synth.translation_seeds(scene.scan, 192, seed=3)
    for i, o in enumerate(scene.objects):
        trans[i] = [o.pose[0, 3], 0.0, o.pose[2, 3]]
    models = pipeline.upload_objects(scene.objects)
    res = pipeline.run_step((scene.scan.pos(1), scene.scan.nor(1)), (scene.scan.pos(2), scene.scan.nor(2)), models, rots, trans, top_k=top_k,
                            nms_dist=0.2, rank=rank, world=world, dist=dist, device=torch.device("cpu"), lanes=lanes, peer=peer)
    np.savez(os.path.join(out_dir, f"r{rank}_{lanes}.npz"), **{f"p{k}": p for k, p in enumerate(res.proposals)},
             **{f"i{k}": i for k, i in enumerate(res.pose_ids)}, n_eval=res.n_evaluations)
    if peer is not None:
        peer.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("lanes,exchange,top_k", [(1, "gloo", 16), (4, "gloo", 16), (4, "peer", 16), (1, "peer", 0)])
def test_pose_sharded_step_equals_single_rank(tmp_path, lanes, exchange, top_k):
    """two ranks (sharing the one GPU; lists exchanged over gloo or through the peer-mapped slots of csrc/peer.cu) through the
    pose-sharded step: every rank ends with the single-rank lists, bit for bit - also with top_k = 0 (every survivor, the
    reference's behaviour), where the merged list must reach the NMS in the single-rank order"""
    import socket
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_rank_worker, args=(2, port, str(tmp_path), lanes, exchange, top_k), nprocs=2, join=True)
    scene = common.small_scene()
    rots, _ = common.rotation_xforms(12)
    trans = synth.translation_seeds(scene.scan, 192, seed=3)
    for i, o in enumerate(scene.objects):
        trans[i] = [o.pose[0, 3], 0.0, o.pose[2, 3]]
    models = pipeline.upload_objects(scene.objects)
    ref = pipeline.run_step((scene.scan.pos(1), scene.scan.nor(1)), (scene.scan.pos(2), scene.scan.nor(2)), models, rots, trans, top_k=top_k,
                            nms_dist=0.2, lanes=1)
    total_eval = 0
    for rank in range(2):
        z = np.load(tmp_path / f"r{rank}_{lanes}.npz")
        total_eval += int(z["n_eval"])
        for k, (p, i) in enumerate(zip(ref.proposals, ref.pose_ids)):
            assert z[f"p{k}"].shape == p.shape and (z[f"p{k}"] == p).all() and (z[f"i{k}"] == i).all(), (rank, k)
    assert total_eval == ref.n_evaluations
