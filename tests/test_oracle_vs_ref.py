"""CPU: the oracle against the UNMODIFIED reference compiled in place (oracle/_ref/librescan_ref.so) on denser seeded
data than the golden fixture.  Skipped where that library does not exist (it is built only where /root/reference is)."""
import numpy as np
import pytest

from oracle import orcbind as O, refbind as R
from rescan_b200 import synth
from tests import common

pytestmark = pytest.mark.skipif(not R.available(), reason="oracle/_ref/librescan_ref.so not built (no /root/reference here)")


@pytest.fixture(scope="module")
def S():
    scene = common.tiny_scene()
    scan = R.RefCloud.from_levels({l: (scene.scan.pos(l), scene.scan.nor(l)) for l in range(5)})
    objs = [R.RefCloud.from_levels({l: (o.cloud.pos(l), o.cloud.nor(l)) for l in range(5)}) for o in scene.objects]
    return scene, scan, objs


def test_sizeof_and_levels(S):
    scene, scan, _ = S
    assert R.load().ref_sizeof_hash_grid() == 120
    assert scan.n(1) == scene.scan.n(1)


@pytest.mark.parametrize("radius,k", [(0.10, 64), (0.05, 16), (0.05, 1), (0.075, 8), (0.2, 32)])
def test_radius_search(S, radius, k):
    scene, scan, _ = S
    p = scene.scan.pos(1)
    rng = np.random.default_rng(7)
    q = np.ascontiguousarray(p[rng.choice(len(p), 1500)] + rng.uniform(-0.03, 0.03, (1500, 3)).astype(np.float32), np.float32)
    og, rg = O.OrcGrid(p, 0.05), scan.grid(1)
    a, b = og.radius_search(q, radius, k), rg.radius_search(q, radius, k)
    assert a[3] == b[3]
    common.check_rows(a[0], a[1], a[2], b[0], b[1], b[2], k)


def test_radius_search_cell_cap(S):
    """more than 512 cells in range: the reference stops enumerating (msh_hash_grid.h:1213)"""
    scene, _, _ = S
    p = scene.scan.pos(3)
    og, rg = O.OrcGrid(p, 0.02), R.RefGrid(p, 0.02)
    rng = np.random.default_rng(8)
    q = np.ascontiguousarray(p[rng.choice(len(p), 200)], np.float32)
    a, b = og.radius_search(q, 0.25, 8), rg.radius_search(q, 0.25, 8)
    common.check_rows(a[0], a[1], a[2], b[0], b[1], b[2], 8)


def test_knn_search(S):
    scene, scan, _ = S
    p = scene.scan.pos(1)
    rng = np.random.default_rng(9)
    q = np.ascontiguousarray(p[rng.choice(len(p), 1000)] + rng.uniform(-0.005, 0.005, (1000, 3)).astype(np.float32), np.float32)
    a, b = O.OrcGrid(p, 0.05).knn_search(q, 8), scan.grid(1).knn_search(q, 8)
    common.check_rows(a[0], a[1], a[2], b[0], b[1], b[2], 8)


@pytest.mark.parametrize("lvl", [4, 3])
def test_scores(S, lvl):
    scene, scan, objs = S
    og = O.OrcGrid(scene.scan.pos(1), 0.05)
    rng = np.random.default_rng(10 + lvl)
    for o, ro in zip(scene.objects, objs):
        xs = np.stack([common.colmajor(m) for _, m in common.perturbed_poses(rng, type("X", (), {"objects": [o]})(), 12)])
        got, _ = O.score_poses(o.cloud.pos(lvl), o.cloud.nor(lvl), og, scene.scan.nor(1), xs, 64, 0.10)
        want, _ = R.score_batch(ro, scan, xs, query_lvl=lvl)
        assert (got == want).all()


def test_icp_align(S):
    scene, _, _ = S
    o = scene.objects[-1]
    rng = np.random.default_rng(12)
    s = common.colmajor(common.perturbed_poses(rng, type("X", (), {"objects": [o]})(), 1, 0.03, 0.08)[0][1])
    ang = np.float32(np.deg2rad(60.0))
    To, eo, _ = O.icp_align(o.cloud.pos(2), o.cloud.nor(2), scene.scan.pos(2), scene.scan.nor(2), s, 0.10, ang)
    Tr, er = R.icp_align(o.cloud.pos(2), o.cloud.nor(2), scene.scan.pos(2), scene.scan.nor(2), s, 0.10, ang)
    assert (To == Tr).all() and np.float32(eo) == np.float32(er)


@pytest.mark.parametrize("lvl,max_dist,max_angle_deg,with_T2", [(2, 0.075, 50.0, False), (0, 0.05, 10.0, False), (2, 0.10, 60.0, True)])
def test_icp_align_other_call_sites(S, lvl, max_dist, max_angle_deg, with_T2):
    """icp_align as rsdb_refine_alignment_of_objects_to_scene (rs_database.h:227) and rsdu_augment_database
    (database_update.cpp:65) call it, and with a non-identity T2: oracle bit-equal to the compiled reference"""
    scene, _, _ = S
    ang = np.float32(np.deg2rad(max_angle_deg))
    rng = np.random.default_rng(40 + lvl)
    T2m = synth.yaw_pose(0.3, 0.4, -0.4, 0.0).astype(np.float64)
    T2 = common.colmajor(T2m.astype(np.float32)) if with_T2 else None
    for o in scene.objects[:2]:
        d = synth.yaw_pose(rng.uniform(-0.04, 0.04), rng.uniform(-0.015, 0.015), rng.uniform(-0.015, 0.015), 0.0)
        m = o.pose.astype(np.float64) @ d.astype(np.float64)
        if with_T2:
            m = T2m @ m
        s = common.colmajor(m.astype(np.float32))
        To, eo, it = O.icp_align(o.cloud.pos(lvl), o.cloud.nor(lvl), scene.scan.pos(lvl), scene.scan.nor(lvl), s, max_dist, ang, T2=T2)
        Tr, er = R.icp_align(o.cloud.pos(lvl), o.cloud.nor(lvl), scene.scan.pos(lvl), scene.scan.nor(lvl), s, max_dist, ang, T2_colmajor=T2)
        assert (To == Tr).all() and np.float32(eo) == np.float32(er) and it > 0


def _nms_case(scene, oi, rng, n=40):
    """proposals of one object: jittered copies of its true pose, far-away poses and a few with failed scores"""
    o = scene.objects[oi]
    props = np.zeros((n, 17), np.float32)
    for j in range(n):
        if j % 5 == 4:
            m = synth.yaw_pose(rng.uniform(0, 6.28), rng.uniform(0.5, 2.5), rng.uniform(0.5, 2.0))
        else:
            d = synth.yaw_pose(rng.uniform(-0.6, 0.6), rng.uniform(-0.5, 0.5), rng.uniform(-0.5, 0.5), rng.uniform(-0.02, 0.02))
            m = (d.astype(np.float64) @ o.pose.astype(np.float64)).astype(np.float32)
        props[j, :16] = common.colmajor(m)
        props[j, 16] = np.float32(rng.uniform(0.3, 0.99)) if j % 7 else np.float32(-1.0)
    props[3, 16] = props[2, 16]  # a score tie: the first maximum wins
    return props


def test_overlap_factor_and_nms(S):
    """intersect.h:309-368 and pose_proposal.cpp:371-452: bit-identical overlap factors, identical survivor lists"""
    scene, _, objs = S
    rng = np.random.default_rng(77)
    db = R.RefDB()
    for o, rc in zip(scene.objects, objs):
        db.add_object(rc, o.uidx, o.class_idx)
    n_pairs = n_pos = 0
    for oi, (o, rc) in enumerate(zip(scene.objects, objs)):
        props = _nms_case(scene, oi, rng)
        for j in range(1, 12):
            for inside, smaller in ((1, 0), (0, 0), (1, 1)):
                want = R.overlap_factor(rc, props[0, :16], props[j, :16], 0.1, inside, smaller)
                got = O.overlap_factor(o.cloud.pos(3), o.cloud.pos(1), props[0, :16], props[j, :16], 0.1, inside, smaller)
                assert got == want, (oi, j, inside, smaller, got, want)
                n_pairs += 1
                n_pos += want > 0
        c_ref, c_orc = R.cloud_centroid(rc), O.centroid(o.cloud.pos(0))
        assert (c_ref == c_orc).all()
        kept_ref = db.nms(oi, props, 0.2)
        keep = O.nms(o.cloud.pos(3), o.cloud.pos(1), c_orc, props, 0.2)
        assert len(kept_ref) == keep.sum() and (kept_ref == props[keep]).all()
        assert 0 < keep.sum() < len(props)
    assert n_pos > n_pairs // 3


def _match_rows(pos0, rows):
    key = {p.tobytes(): i for i, p in enumerate(pos0)}
    return np.array([key[p.tobytes()] for p in rows], np.int32)


def test_poisson_levels_against_reference():
    """level building: rs_pointcloud_compute_levels of the compiled reference against the oracle's restatement, on a scan,
    and on a patch dense enough that the reference's max_n_neigh cap binds at level 4 (a ball of 8 cm holds > 1024 points)"""
    scene = common.tiny_scene()
    rng = np.random.default_rng(8)
    u, v = np.meshgrid(np.arange(0, 0.32, 0.0035), np.arange(0, 0.32, 0.0035), indexing="ij")
    dense = np.stack([u.reshape(-1), np.zeros(u.size), v.reshape(-1)], axis=1) + rng.normal(0, 0.0004, (u.size, 3))
    for pos0 in (scene.scan.pos(0), dense.astype(np.float32)):
        pos0 = np.ascontiguousarray(pos0, np.float32)
        nor0 = np.tile(np.array([0, 1, 0], np.float32), (len(pos0), 1))
        rc = R.RefCloud.from_level0(pos0, nor0)
        for lvl in range(1, 5):
            want = _match_rows(pos0, rc.level(lvl)[0])
            assert (O.poisson_level(pos0, lvl) == want).all(), lvl


def test_coverage_grids_against_reference(S):
    """coverage term of the arrangement optimiser: grid set-up and rasterisation of the oracle against the
    reference's own primitives (isect_grid3d_init, msh_mat4_vec3_mul, isect_grid3d_cell_from_world_space)"""
    scene, _, _ = S
    mn, mx = scene.scan.pos(0).min(0), scene.scan.pos(0).max(0)
    for voxel in (0.05, 0.15):
        ro, oo, no = O.cov_grid(mn, mx, voxel)
        rr, orr, nr = R.cov_grid(mn, mx, voxel)
        assert (ro == rr).all() and (oo == orr).all() and no == nr
        go = O.cov_rasterize(scene.scan.pos(2), None, ro, oo, voxel)
        gr = R.cov_rasterize(mn, mx, scene.scan.pos(2), None, nr, voxel)
        assert (go == gr).all() and go.sum() > 100
        rng = np.random.default_rng(6)
        for o in scene.objects:
            d = synth.yaw_pose(rng.uniform(0, 6.28), rng.uniform(-1.5, 1.5), rng.uniform(-1.5, 1.5), rng.uniform(-0.2, 0.2))
            pose = common.colmajor((d.astype(np.float64) @ o.pose.astype(np.float64)).astype(np.float32))  # some of it leaves the grid
            ao = O.cov_rasterize(o.cloud.pos(2), pose, ro, oo, voxel)
            ar = R.cov_rasterize(mn, mx, o.cloud.pos(2), pose, nr, voxel)
            assert (ao == ar).all()


def test_plane_inlier_counts_against_reference(S):
    """oracle restatement of evaluate_plane_model against the compiled reference on random point triples of the scan"""
    scene = S[0]
    p, n = scene.scan.pos(2), scene.scan.nor(2)
    rng = np.random.default_rng(8)
    w = (np.abs(n[:, 1]) < 0.2).astype(np.float64)
    idx = rng.integers(0, len(p), (60, 3))
    a, b, c = p[idx[:, 0]], p[idx[:, 1]], p[idx[:, 2]]
    nr = np.cross(b - a, c - a).astype(np.float32)
    nr = (nr / np.linalg.norm(nr, axis=1, keepdims=True)).astype(np.float32)
    planes = np.concatenate([a, nr], 1).astype(np.float32)
    assert (O.plane_inlier_counts(p, w > 0.01, planes, 0.033) == R.plane_inlier_counts(p, w, planes, 0.033)).all()
