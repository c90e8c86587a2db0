"""NMS on the GPU: rsgpu_overlap_factors / rsgpu_nms through the C ABI against the reference's golden
vectors (tests/golden/nms_golden.npz, written from the compiled reference) and against the oracle on a denser scene.
Overlap factors are integer ratios formed in float exactly like the reference: bit equality is required."""
import os

import numpy as np
import pytest

from oracle import orcbind as O
from rescan_b200 import api, synth
from tests import common

pytestmark = pytest.mark.gpu


def test_overlap_and_nms_match_golden():
    z, scan, objs = common.golden()
    g = dict(np.load(os.path.join(os.path.dirname(common.GOLDEN), "nms_golden.npz")))
    for i, o in enumerate(objs):
        c3, c1 = api.PointCloud(o.pos(3), o.nor(3)), api.PointCloud(o.pos(1), o.nor(1))
        props = g[f"nms{i}_props"]
        got = api.overlap_factors(c3, c1, props[0, :16], props[:, :16], 0.1, 1, 0)
        assert (got == g[f"nms{i}_overlap"]).all()
        gotb = api.overlap_factors(c3, c1, props[0, :16], props[:16, :16], 0.1, 0, 1)
        assert (gotb == g[f"nms{i}_overlap_boundary"]).all()
        keep = api.non_maxima_suppression(c3, c1, g[f"nms{i}_centroid"], props, 0.2)
        assert keep.sum() == len(g[f"nms{i}_kept"]) and (props[keep] == g[f"nms{i}_kept"]).all()


def test_nms_matches_oracle_dense_scene():
    scene = common.small_scene()
    rng = np.random.default_rng(91)
    for o in scene.objects:
        n = 60
        props = np.zeros((n, 17), np.float32)
        for j in range(n):
            d = synth.yaw_pose(rng.uniform(-0.8, 0.8), rng.uniform(-0.6, 0.6), rng.uniform(-0.6, 0.6), rng.uniform(-0.02, 0.02))
            props[j, :16] = common.colmajor((d.astype(np.float64) @ o.pose.astype(np.float64)).astype(np.float32))
            props[j, 16] = np.float32(rng.uniform(0.2, 1.0)) if j % 9 else np.float32(-1.0)
        c3, c1 = api.PointCloud(o.cloud.pos(3), o.cloud.nor(3)), api.PointCloud(o.cloud.pos(1), o.cloud.nor(1))
        want = np.array([O.overlap_factor(o.cloud.pos(3), o.cloud.pos(1), props[5, :16], props[j, :16]) for j in range(20)], np.float32)
        got = api.overlap_factors(c3, c1, props[5, :16], props[:20, :16])
        assert (got == want).all()
        cen = O.centroid(o.cloud.pos(0))
        keep_o = O.nms(o.cloud.pos(3), o.cloud.pos(1), cen, props, 0.2)
        keep_g = api.non_maxima_suppression(c3, c1, cen, props, 0.2)
        assert (keep_g == keep_o).all() and 0 < keep_g.sum() < n


def test_nms_edge_cases():
    scene = common.tiny_scene()
    o = scene.objects[0]
    c3, c1 = api.PointCloud(o.cloud.pos(3), o.cloud.nor(3)), api.PointCloud(o.cloud.pos(1), o.cloud.nor(1))
    cen = O.centroid(o.cloud.pos(0))
    assert len(api.non_maxima_suppression(c3, c1, cen, np.zeros((0, 17), np.float32))) == 0
    one = np.concatenate([common.colmajor(o.pose), [np.float32(0.7)]]).astype(np.float32)[None]
    assert api.non_maxima_suppression(c3, c1, cen, one).tolist() == [True]
    # identical poses: overlap 1 -> only the first (highest, first on ties) survives
    same = np.repeat(one, 4, axis=0)
    assert api.non_maxima_suppression(c3, c1, cen, same).tolist() == [True, False, False, False]
    assert (api.overlap_factors(c3, c1, one[0, :16], same[:, :16]) == 1.0).all()
    # all failed scores (-1): the best is kept, the rest goes by score < 0.01 (pose_proposal.cpp:422)
    failed = same.copy()
    failed[:, 16] = -1.0
    far = synth.yaw_pose(0.3, 2.0, 1.5)
    failed[2, :16] = common.colmajor(far)
    assert api.non_maxima_suppression(c3, c1, cen, failed).tolist() == O.nms(o.cloud.pos(3), o.cloud.pos(1), cen, failed).tolist()


def test_nms_one_launch_equals_round_by_round():
    """default: every pair the greedy loop can ask for in ONE overlap launch; "nms_impl" = "rounds": one launch per kept pose
    (round 1).  Same keep vector, also for a list too long for the one-launch path (400 poses -> round by round either way)."""
    scene = common.small_scene()
    rng = np.random.default_rng(17)
    n_diff = 0
    for o in scene.objects:
        for n in (3, 64, 150, 400):
            props = np.zeros((n, 17), np.float32)
            for j in range(n):
                d = synth.yaw_pose(rng.uniform(-3.1, 3.1), rng.uniform(-0.7, 0.7), rng.uniform(-0.7, 0.7), rng.uniform(-0.02, 0.02))
                props[j, :16] = common.colmajor((d.astype(np.float64) @ o.pose.astype(np.float64)).astype(np.float32))
                props[j, 16] = np.float32(rng.uniform(0.0, 1.0)) if j % 7 else np.float32(-1.0)
            props[n // 2, 16] = props[n // 3, 16]  # a tie: first index wins
            c3, c1 = api.PointCloud(o.cloud.pos(3), o.cloud.nor(3)), api.PointCloud(o.cloud.pos(1), o.cloud.nor(1))
            cen = O.centroid(o.cloud.pos(0))
            a = api.non_maxima_suppression(c3, c1, cen, props, 0.2)
            api.set_option("nms_impl", "rounds")
            b = api.non_maxima_suppression(c3, c1, cen, props, 0.2)
            api.set_option("nms_impl", None)
            assert (a == b).all(), (o.uidx, n)
            n_diff += int(a.sum())
            if n == 64:
                assert (a == O.nms(o.cloud.pos(3), o.cloud.pos(1), cen, props, 0.2)).all()
    assert n_diff > 0
