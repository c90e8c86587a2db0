"""GPU parity tests proper: every call goes through the C ABI (rescan_b200.api -> librsgpu.so) and is compared
with the CPU oracle (oracle/rescan_oracle.c, pinned against the compiled reference) on the same seeded inputs.

Tolerances (BASELINE.json north_star): NN indices bit-exact wherever distances are not tied, squared distances
bit-exact; scores within 1e-4 relative; ICP poses within 1e-5 m / 1e-5 rad.
"""
import numpy as np
import pytest

from oracle import orcbind as O
from rescan_b200 import api, synth
from tests import common

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def scene():
    return common.small_scene()


@pytest.fixture(scope="module")
def scan1(scene):
    p, n = scene.scan.pos(1), scene.scan.nor(1)
    return dict(pos=p, nor=n, orc=O.OrcGrid(p, 0.05), gpu=api.HashGrid(p, 0.05, normals=n))


def _queries(rng, pts, n, jitter):
    q = pts[rng.choice(len(pts), n)] + rng.uniform(-jitter, jitter, (n, 3)).astype(np.float32)
    return np.ascontiguousarray(q, np.float32)


def _compare_rows(gi, gd, gn, oi, od, on, k):
    assert (gn == on).all(), f"neighbour counts differ for {(gn != on).sum()} queries"
    m = np.arange(k)[None, :] < on[:, None]
    assert (gd[m] == od[m]).all(), "squared distances are not bit-identical"
    # indices must agree wherever the distance is unique within its row and not tied with the cut-off
    diff = m & (gi != oi)
    if diff.any():
        rows = np.unique(np.nonzero(diff)[0])
        for r in rows:
            c = np.nonzero(diff[r])[0]
            d = od[r, : on[r]]
            for j in c:
                tied = (d == d[j]).sum() > 1 or j == k - 1
                assert tied, f"index mismatch without a distance tie (row {r}, col {j})"


def test_grid_build_matches_reference_layout(scan1):
    gi, oi = scan1["gpu"].info(), scan1["orc"].info()
    assert (gi["dims"] == oi["dims"]).all()
    assert gi["cell_size"] == oi["cell_size"] and gi["inv_cell_size"] == oi["inv_cell_size"]
    assert (gi["min_pt"] == oi["min_pt"]).all() and (gi["max_pt"] == oi["max_pt"]).all()
    assert gi["n_pts"] == oi["n_pts"] and gi["n_bins"] == oi["n_bins"] and gi["max_n_pts_in_bin"] == oi["max_n_pts_in_bin"]
    gx, gidx = scan1["gpu"].data()
    ox, oidx = scan1["orc"].data()
    assert (gidx == oidx).all() and (gx == ox).all()


@pytest.mark.parametrize("radius,k", [(0.10, 64), (0.05, 16), (0.05, 1), (0.075, 8), (0.10, 100), (0.03, 300), (0.2, 32)])
def test_radius_search(scan1, radius, k):
    rng = np.random.default_rng(11)
    q = _queries(rng, scan1["pos"], 3000, 0.03)
    q[:50] += 5.0  # far outside the grid: no neighbours
    q[50:100] = scan1["pos"][:50]  # exact hits (distance 0)
    gi, gd, gn, gt = scan1["gpu"].radius_search(q, radius, k)
    oi, od, on, ot = scan1["orc"].radius_search(q, radius, k)
    assert gt == ot
    _compare_rows(gi, gd, gn, oi, od, on, k)
    assert (gn[:50] == 0).all()


def test_radius_search_many_cells_cap(scene):
    """radius far above the cell size: the reference examines at most 512 cells per query (msh_hash_grid.h:1213)"""
    p = scene.scan.pos(3)
    og, gg = O.OrcGrid(p, 0.02), api.HashGrid(p, 0.02)
    rng = np.random.default_rng(5)
    q = _queries(rng, p, 500, 0.02)
    for radius, k in [(0.15, 16), (0.25, 8)]:
        gi, gd, gn, gt = gg.radius_search(q, radius, k)
        oi, od, on, ot = og.radius_search(q, radius, k)
        assert gt == ot
        _compare_rows(gi, gd, gn, oi, od, on, k)


def test_radius_search_empty_and_auto_cell():
    g = api.HashGrid(np.zeros((0, 3), np.float32), 0.05)
    i, d, n, t = g.radius_search(np.zeros((4, 3), np.float32), 0.1, 4)
    assert t == 0 and (n == 0).all()
    rng = np.random.default_rng(3)
    p = rng.uniform(0, 1, (5000, 3)).astype(np.float32)
    og, gg = O.OrcGrid(p, 0.0), api.HashGrid(p, 0.0)  # radius <= 0: automatic cell size (msh_hash_grid.h:443-444)
    assert gg.info()["cell_size"] == og.info()["cell_size"] and (gg.info()["dims"] == og.info()["dims"]).all()
    q = rng.uniform(0, 1, (500, 3)).astype(np.float32)
    gi, gd, gn, gt = gg.radius_search(q, 0.08, 12)
    oi, od, on, ot = og.radius_search(q, 0.08, 12)
    assert gt == ot
    _compare_rows(gi, gd, gn, oi, od, on, 12)


@pytest.mark.parametrize("k", [1, 8, 40])
def test_knn_search(scan1, k):
    rng = np.random.default_rng(12)
    q = _queries(rng, scan1["pos"], 2000, 0.01)
    gi, gd, gn, gt = scan1["gpu"].knn_search(q, k)
    oi, od, on, ot = scan1["orc"].knn_search(q, k)
    assert gt == ot
    _compare_rows(gi, gd, gn, oi, od, on, k)


def test_k_above_limit_fails_loudly(scan1):
    with pytest.raises(api.RsgpuError):
        scan1["gpu"].radius_search(scan1["pos"][:4], 0.1, api.RSGPU_MAX_K + 1)


def _score_case(scene, scan1, lvl, poses, k=64):
    res = []
    for o, m in poses:
        x = common.colmajor(m)
        cloud = api.PointCloud(o.cloud.pos(lvl), o.cloud.nor(lvl))
        g = api.compute_object_alignment_scores(cloud, scan1["gpu"], x[None, :], k, 0.10)[0]
        r, _ = O.score_poses(o.cloud.pos(lvl), o.cloud.nor(lvl), scan1["orc"], scan1["nor"], x[None, :], k, 0.10)
        res.append((g, r[0]))
    return np.array(res)


@pytest.mark.parametrize("lvl,k", [(4, 64), (3, 64), (2, 64), (1, 32), (4, 4)])
def test_pose_scores_match_oracle(scene, scan1, lvl, k):
    rng = np.random.default_rng(100 + lvl)
    poses = common.perturbed_poses(rng, scene, 5)
    # plus poses in free space / half way into walls
    for o in scene.objects:
        poses.append((o, synth.yaw_pose(rng.uniform(0, 6.28), rng.uniform(0, 3), rng.uniform(0, 2.5))))
    r = _score_case(scene, scan1, lvl, poses, k)
    g, o = r[:, 0], r[:, 1]
    assert (o > 0.3).sum() >= 5, "test poses should include well-aligned ones"
    assert np.allclose(g, o, rtol=1e-4, atol=1e-7), f"max rel err {np.abs(g - o).max()}"
    assert (g == o).mean() > 0.9, "scores are expected to be bit-identical almost everywhere"


def test_pose_grid_and_propose_match_oracle(scene, scan1):
    rots, ang = common.rotation_xforms(8)
    rng = np.random.default_rng(21)
    trans = synth.translation_seeds(scene.scan, 96, seed=5)
    # make sure the true translations are among the seeds so that proposals survive all three levels
    for i, o in enumerate(scene.objects):
        trans[i] = [o.pose[0, 3], 0.0, o.pose[2, 3]]
    for o in scene.objects:
        if o.is_static:
            continue
        c4, c3, c2 = (api.PointCloud(o.cloud.pos(l), o.cloud.nor(l)) for l in (4, 3, 2))
        g = api.score_pose_grid(c4, scan1["gpu"], rots, trans)
        xf = O.make_pose_grid(ang, trans)
        ref, _ = O.score_poses(o.cloud.pos(4), o.cloud.nor(4), scan1["orc"], scan1["nor"], xf, 64, 0.10, n_threads=8)
        ref = ref.reshape(len(trans), len(ang))
        assert np.allclose(g, ref, rtol=1e-4, atol=1e-7)
        # oracle pipeline: select at level 4, verify at 3 and 2 (pose_proposal.cpp:213-243, 285-297)
        ot, orr, osc = O.select_proposals(ref, 0.25)
        want = []
        for t, r, s in zip(ot, orr, osc):
            x = xf[t, r]
            for lvl, thr in ((3, 0.35), (2, 0.40)):
                if s > 0:
                    v, _ = O.score_poses(o.cloud.pos(lvl), o.cloud.nor(lvl), scan1["orc"], scan1["nor"], x[None], 64, 0.10)
                    s = v[0] if v[0] > thr else -1.0
            want.append((t * len(ang) + r, s))
        props, ids = api.propose_poses(c4, c3, c2, scan1["gpu"], rots, trans)
        assert list(ids) == [w[0] for w in want]
        assert np.allclose(props[:, 16], [w[1] for w in want], rtol=1e-4, atol=1e-7)
        for p, pid in zip(props, ids):
            assert (p[:16] == xf[pid // len(ang), pid % len(ang)]).all()
        # top-k variant: descending scores, ties by pose id
        top, tid = api.propose_poses(c4, c3, c2, scan1["gpu"], rots, trans, top_k=3)
        order = sorted(range(len(want)), key=lambda i: (-props[i, 16], ids[i]))[:3]
        assert list(tid) == [ids[i] for i in order]


def _pose_delta(a16, b16):
    A = a16.reshape(4, 4).T.astype(np.float64)
    B = b16.reshape(4, 4).T.astype(np.float64)
    dt = np.linalg.norm(A[:3, 3] - B[:3, 3])
    R = A[:3, :3] @ B[:3, :3].T
    ang = np.linalg.norm(R - R.T) / (2 * np.sqrt(2))  # = |sin(theta)|, well conditioned for tiny angles (arccos is not)
    return dt, ang


def test_icp_align_matches_oracle(scene):
    """icp_align at the pose_proposal call site (main.cpp:195: level 2, 0.10 m, 60 degrees) for perturbed starts of every object:
    the default path accumulates in the reference's float order, so pose, error and iteration count are BIT-identical to the
    oracle's (itself pinned bit-equal to the compiled reference); 1e-5 m / 1e-5 rad is north_star's tolerance, kept as a second
    assertion so that a failure says how far off it is"""
    p2, n2 = scene.scan.pos(2), scene.scan.nor(2)
    grid = api.HashGrid(p2, 0.05, normals=n2)
    rng = np.random.default_rng(31)
    n_iterated = 0
    for o in scene.objects:
        starts = [common.colmajor(m) for _, m in common.perturbed_poses(rng, type("S", (), {"objects": [o]})(), 4, 0.03, 0.08)]
        cloud = api.PointCloud(o.cloud.pos(2), o.cloud.nor(2))
        T, err, it = api.icp_align(cloud, grid, np.stack(starts), 0.10, np.float32(np.deg2rad(60.0)))
        for b, s in enumerate(starts):
            To, eo, ito = O.icp_align(o.cloud.pos(2), o.cloud.nor(2), p2, n2, s, 0.10, np.float32(np.deg2rad(60.0)))
            dt, da = _pose_delta(T[b], To)
            assert dt < 1e-5 and da < 1e-5, (dt, da)
            assert (T[b] == To).all() and err[b] == np.float32(eo) and it[b] == ito
            n_iterated += int(ito > 0)
    assert n_iterated >= 2 * len(scene.objects)


@pytest.mark.parametrize("lvl,max_dist,max_angle_deg,T2_shift", [
    (2, 0.075, 50.0, None),   # rsdb_refine_alignment_of_objects_to_scene (lib/rs/rs_database.h:227-229)
    (0, 0.05, 10.0, None),    # rsdu_augment_database (apps/segment_transfer/database_update.cpp:65-67), level 0
    (2, 0.10, 60.0, 0.4),     # a non-identity T2 (icp.h:339-347 maps through T2^-1 * T1)
])
def test_icp_call_sites_bit_exact(scene, lvl, max_dist, max_angle_deg, T2_shift):
    """the other icp_align call sites of the reference, with their own levels / distances / angles: refined pose, error and
    iteration count identical to the oracle's (which is pinned bit-equal to the compiled reference)"""
    ps, ns = scene.scan.pos(lvl), scene.scan.nor(lvl)
    grid = api.HashGrid(ps, 0.05, normals=ns)
    rng = np.random.default_rng(77 + lvl)
    ang = np.float32(np.deg2rad(max_angle_deg))
    T2 = None
    if T2_shift is not None:
        T2 = common.colmajor(synth.yaw_pose(0.3, T2_shift, -T2_shift, 0.0))
    n = 0
    for o in scene.objects[:3]:
        cloud = api.PointCloud(o.cloud.pos(lvl), o.cloud.nor(lvl))
        starts = []
        for _ in range(3):
            d = synth.yaw_pose(rng.uniform(-0.04, 0.04), rng.uniform(-0.015, 0.015), rng.uniform(-0.015, 0.015), 0.0)
            m = o.pose.astype(np.float64) @ d.astype(np.float64)
            if T2_shift is not None:  # T1 is then expressed so that T2^-1 * T1 lands on the scan
                m = synth.yaw_pose(0.3, T2_shift, -T2_shift, 0.0).astype(np.float64) @ m
            starts.append(common.colmajor(m.astype(np.float32)))
        T, err, it = api.icp_align(cloud, grid, np.stack(starts), max_dist, ang, T2=T2)
        for b, s in enumerate(starts):
            To, eo, ito = O.icp_align(o.cloud.pos(lvl), o.cloud.nor(lvl), ps, ns, s, max_dist, ang, T2=T2)
            assert (T[b] == To).all() and err[b] == np.float32(eo) and it[b] == ito
            n += int(ito > 0)
    assert n >= 6  # the runs actually iterated


def test_icp_degenerate_inputs(scene):
    p2, n2 = scene.scan.pos(2), scene.scan.nor(2)
    grid = api.HashGrid(p2, 0.05, normals=n2)
    o = scene.objects[-1]
    cloud = api.PointCloud(o.cloud.pos(2), o.cloud.nor(2))
    far = common.colmajor(synth.yaw_pose(0.0, 50.0, 50.0))  # no correspondences at all
    T, err, it = api.icp_align(cloud, grid, far[None], 0.10, np.float32(np.deg2rad(60.0)))
    To, eo, ito = O.icp_align(o.cloud.pos(2), o.cloud.nor(2), p2, n2, far, 0.10, np.float32(np.deg2rad(60.0)))
    assert (T[0] == far).all() and (To == far).all() and it[0] == 0 and ito == 0 and err[0] == np.float32(eo)


def test_label_transfer_and_unary_terms(scene):
    p1, n1 = scene.scan.pos(1), scene.scan.nor(1)
    V = len(p1)
    objs = sorted(scene.objects, key=lambda o: o.is_static)  # dynamic first, static last (rs_pointcloud_filters.cpp:823-835)
    poses = np.stack([common.colmajor(o.pose) for o in objs])
    og = [O.OrcGrid(o.cloud.pos(1), 0.05) for o in objs]
    gg = [api.HashGrid(o.cloud.pos(1), 0.05, normals=o.cloud.nor(1)) for o in objs]
    n_dyn = sum(not o.is_static for o in objs)
    lab_o, min_o = np.zeros(V, np.int8), np.full(V, 1e9, np.float32)
    lab_g, min_g = np.zeros(V, np.int8), np.full(V, 1e9, np.float32)
    for first, last, r in ((0, n_dyn, 0.05), (n_dyn, len(objs), 0.075)):
        O.assign_labels(p1, n1, poses, og, [o.cloud.nor(1) for o in objs], first, last, r, lab_o, min_o)
        api.assign_labels(p1, n1, poses, gg, first, last, r, lab_g, min_g)
    assert (lab_o > 0).sum() > 1000
    assert (lab_g == lab_o).all() and (min_g == min_o).all()
    labels = lab_g.astype(np.int32)
    L = int(labels.max()) + 5
    is_static = np.zeros(L, np.uint8)
    is_static[n_dyn + 1: len(objs) + 1] = 1
    assert (api.unary_costs(labels, is_static, L) == O.unary_costs(labels, is_static, L)).all()


def test_neighborhood_edges(scan1):
    n = 20000
    p, nr = scan1["pos"], scan1["nor"]
    gn, gw = api.neighborhood(scan1["gpu"], p, nr)
    on, ow = O.neighborhood(scan1["orc"], p, nr)
    assert gn.shape == on.shape
    same = gn == on
    assert same.mean() > 0.999  # rows differ only by tie order
    assert np.allclose(gw[same], ow[same], rtol=2e-6, atol=1e-9)
    # a row that differs holds the same neighbours in another order (exact distance ties), or - when the tie straddles the
    # k-th place - differs in tied entries only: compare the distances of the entries that are not shared
    pf = p.astype(np.float32)
    for r in np.unique(np.nonzero(~same)[0])[:200]:
        a, b = set(gn[r].tolist()), set(on[r].tolist())
        if a == b:
            continue
        da = sorted(float(((pf[i] - pf[r]) ** 2).sum()) for i in a - b if i >= 0)
        db = sorted(float(((pf[i] - pf[r]) ** 2).sum()) for i in b - a if i >= 0)
        assert len(da) == len(db) and np.allclose(da, db, rtol=0, atol=1e-12), (r, da, db)
