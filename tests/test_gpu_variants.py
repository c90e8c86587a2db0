"""Every kernel variant behind rsgpu_set_option must give the same answers: the production defaults are tested against
the oracle in test_gpu_parity.py; here the alternatives (thread- vs warp-per-query search, warp-per-query vs 4-lane
group scoring, bound pruning on/off, resident-block vs iteration-synchronous ICP) are held to the defaults bit for bit,
and the thread-per-query search to the oracle directly."""
import numpy as np
import pytest

from oracle import orcbind as O
from rescan_b200 import api, synth
from tests import common
from tests.test_gpu_parity import _compare_rows, _queries

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def scene():
    return common.small_scene()


@pytest.fixture(autouse=True)
def _restore_options():
    yield
    for name in ("search_impl", "score_impl", "prune", "icp_impl", "search", "score_g", "score_minb", "score_warps", "dense_impl", "dense_cap", "dense_bps", "icp_ctas", "dense_scratch", "dense_sub", "dense_serial", "nms_impl", "search_sub"):
        api.set_option(name, None)


@pytest.mark.parametrize("impl", ["lane", "warp"])
@pytest.mark.parametrize("radius,k", [(0.05, 1), (0.05, 16), (0.075, 8), (0.10, 3), (0.02, 5)])
def test_radius_search_variants_match_oracle(scene, impl, radius, k):
    p = scene.scan.pos(1)
    og, gg = O.OrcGrid(p, 0.05), api.HashGrid(p, 0.05)
    rng = np.random.default_rng(17)
    q = _queries(rng, p, 4000, 0.03)
    q[:40] += 7.0           # outside the grid
    q[40:80] = p[:40]       # exact hits
    api.set_option("search_impl", impl)
    gi, gd, gn, gt = gg.radius_search(q, radius, k)
    oi, od, on, ot = og.radius_search(q, radius, k)
    assert gt == ot
    _compare_rows(gi, gd, gn, oi, od, on, k)


def test_radius_search_lane_many_cells(scene):
    """thread-per-query kernel on windows far larger than 27 cells (capped at 512 like the reference)"""
    p = scene.scan.pos(3)
    og, gg = O.OrcGrid(p, 0.02), api.HashGrid(p, 0.02)
    q = _queries(np.random.default_rng(5), p, 400, 0.02)
    api.set_option("search_impl", "lane")
    for radius, k in [(0.15, 16), (0.25, 8)]:
        gi, gd, gn, gt = gg.radius_search(q, radius, k)
        oi, od, on, ot = og.radius_search(q, radius, k)
        assert gt == ot
        _compare_rows(gi, gd, gn, oi, od, on, k)


def _propose_all(scene, grid, rots, trans, top_k=0):
    out = []
    for o in scene.objects:
        if o.is_static:
            continue
        c4, c3, c2 = (api.PointCloud(o.cloud.pos(l), o.cloud.nor(l)) for l in (4, 3, 2))
        props, ids = api.propose_poses(c4, c3, c2, grid, rots, trans, top_k=top_k)
        out.append((props.copy(), ids.copy()))
    return out


def test_scoring_variants_bit_identical(scene):
    p, n = scene.scan.pos(1), scene.scan.nor(1)
    grid = api.HashGrid(p, 0.05, normals=n)
    rots, _ = common.rotation_xforms(12)
    trans = synth.translation_seeds(scene.scan, 160, seed=9)
    for i, o in enumerate(scene.objects):
        trans[i] = [o.pose[0, 3], 0.0, o.pose[2, 3]]
    o = [x for x in scene.objects if not x.is_static][0]
    c4 = api.PointCloud(o.cloud.pos(4), o.cloud.nor(4))
    base_scores = api.score_pose_grid(c4, grid, rots, trans)
    base_props = _propose_all(scene, grid, rots, trans)
    assert sum(len(pr) for pr, _ in base_props) > 0
    W = {"dense_impl": "warp"}  # the first design of the dense search (one warp per pose); the default is the cell-binned one
    for opts in ({"score_impl": "coop"}, {"prune": "0"}, W, dict(W, prune="0"), dict(W, score_g="8"), dict(W, score_minb="4", score_warps="4"),
                 dict(W, score_warps="2"), dict(W, score_warps="1", score_minb="16"), dict(W, search="lane"), {"score_g": "8"}, {"search": "lane"},
                 {"dense_cap": "20000"}, {"dense_cap": "300000", "dense_bps": "1"}, {"dense_bps": "6", "prune": "0"}, {"dense_sub": "n"}, {"dense_sub": "o"},
                 {"dense_serial": "0"}):
        for k, v in opts.items():
            api.set_option(k, v)
        s = api.score_pose_grid(c4, grid, rots, trans)
        assert (s == base_scores).all(), f"dense scores differ under {opts}"
        for (pa, ia), (pb, ib) in zip(base_props, _propose_all(scene, grid, rots, trans)):
            assert (ia == ib).all() and (pa == pb).all(), f"proposals differ under {opts}"
        for k in opts:
            api.set_option(k, None)


@pytest.mark.parametrize("which", ["tiny", "small"])
def test_dense_binned_equals_warp_design(which):
    """the cell-binned, shared-memory-staged dense search (csrc/dense_binned.cuh) against the first design (one warp per pose,
    csrc/nearest_group.cuh): every score of the pose grid bit-identical, for every object, with one chunk and with many
    (dense_cap), on a sparse scene (blocks staged) and on a dense one (1 cm spacing: blocks above the staging capacity are
    read from global memory; the k = 64 cap binds), seeds on the scan's border included (home cell outside the grid)"""
    scene = common.tiny_scene() if which == "tiny" else common.small_scene()
    grid = api.HashGrid(scene.scan.pos(1), 0.05, normals=scene.scan.nor(1))
    rots, _ = common.rotation_xforms(9)
    trans = synth.translation_seeds(scene.scan, 96, seed=21)
    for i, o in enumerate(scene.objects):
        trans[i] = [o.pose[0, 3], 0.0, o.pose[2, 3]]
    lo, hi = scene.scan.pos(1).min(0), scene.scan.pos(1).max(0)
    trans[-4:] = [[lo[0], 0, lo[2]], [hi[0], 0, hi[2]], [hi[0] + 0.05, 0, 0.5 * (lo[2] + hi[2])], [0.5 * (lo[0] + hi[0]), 0, hi[2] + 0.08]]
    n_pos = 0
    for o in scene.objects:
        for lvl, k in ((4, 64), (3, 64), (4, 8)):  # level 3 as a stand-in for a big object; a small k makes the cap bind
            c = api.PointCloud(o.cloud.pos(lvl), o.cloud.nor(lvl))
            if len(c) > 4096:
                continue
            api.set_option("dense_impl", "warp")
            want = api.score_pose_grid(c, grid, rots, trans, max_n_neigh=k)
            api.set_option("dense_impl", None)
            for cap in (None, "5000", "70000"):
                api.set_option("dense_cap", cap)
                got = api.score_pose_grid(c, grid, rots, trans, max_n_neigh=k)
                api.set_option("dense_cap", None)
                assert (got == want).all(), (which, lvl, k, cap, int((got != want).sum()))
            n_pos += int((want > 0.25).sum())
    assert n_pos > 0


def test_icp_variants_bit_identical(scene):
    p2, n2 = scene.scan.pos(2), scene.scan.nor(2)
    grid = api.HashGrid(p2, 0.05, normals=n2)
    rng = np.random.default_rng(33)
    objs, starts = [], []
    for o in scene.objects:
        objs.append(api.PointCloud(o.cloud.pos(2), o.cloud.nor(2)))
        starts.append(np.stack([common.colmajor(m) for _, m in common.perturbed_poses(rng, type("S", (), {"objects": [o]})(), 5, 0.03, 0.08)]))
    ang = np.float32(np.deg2rad(60.0))
    base = api.icp_align_multi(objs, grid, [s.copy() for s in starts], 0.10, ang)
    # default = two launches per iteration, four iterations replayed as one CUDA graph; "split" = the same launches one by one;
    # "persistent" = one launch with a device work queue; "block" = one resident
    # block per alignment; "icp_ctas" = size of the persistent grid (1 block: every chunk and every solve on the same block)
    for opt, val in (("icp_impl", "block"), ("icp_impl", "split"), ("icp_impl", "persistent"), ("icp_ctas", "1"), ("icp_ctas", "7"), ("icp_ctas", "512")):
        api.set_option("icp_impl", "persistent")
        api.set_option(opt, val)
        alt = api.icp_align_multi(objs, grid, [s.copy() for s in starts], 0.10, ang)
        api.set_option(opt, None)
        api.set_option("icp_impl", None)
        for (Ta, ea, ia), (Tb, eb, ib) in zip(base, alt):
            assert (ia == ib).all() and (ea == eb).all() and (Ta == Tb).all(), (opt, val)
    # an iteration cap below the natural count stops both variants at the same iteration
    o = objs[0]
    a = api.icp_align(o, grid, starts[0].copy(), 0.10, ang, max_iter=7)
    for other in ("persistent", "split"):
        api.set_option("icp_impl", other)
        b = api.icp_align(o, grid, starts[0].copy(), 0.10, ang, max_iter=7)
        api.set_option("icp_impl", None)
        assert (a[2] == b[2]).all() and (a[0] == b[0]).all() and (a[1] == b[1]).all() and int(a[2].max()) <= 7, other
    assert max(int(i.max()) for _, _, i in base) > 6


def test_translation_order_does_not_change_proposals():
    """rsgpu_propose_opts_t.translation_ids: the translations handed over along a space-filling curve (or in any other
    order) with the caller's numbering give bit for bit the proposals, ids and order of the caller's own order"""
    from rescan_b200 import posegrid
    scene = common.small_scene()
    grid = api.HashGrid(scene.scan.pos(1), 0.05, normals=scene.scan.nor(1))
    rots, _ = common.rotation_xforms(12)
    trans = synth.translation_seeds(scene.scan, 300, seed=13)
    for i, o in enumerate(scene.objects):
        trans[i] = [o.pose[0, 3], 0.0, o.pose[2, 3]]
    rng = np.random.default_rng(3)
    n_checked = 0
    for o in scene.objects:
        if o.is_static:
            continue
        c4, c3, c2 = (api.PointCloud(o.cloud.pos(l), o.cloud.nor(l)) for l in (4, 3, 2))
        for top_k in (0, 5, 64):
            base, bid = api.propose_poses(c4, c3, c2, grid, rots, trans, top_k=top_k)
            for perm in (posegrid.spatial_order(trans), rng.permutation(len(trans)).astype(np.int64)):
                got, gid = api.propose_poses(c4, c3, c2, grid, rots, np.ascontiguousarray(trans[perm]), top_k=top_k, translation_ids=perm)
                assert (gid == bid).all() and (got == base).all()
                n_checked += len(base)
    assert n_checked > 0


@pytest.mark.parametrize("n_pts,radius", [(400_000, 0.10), (150_000, 0.25)])
def test_sub_cell_radius_search_equals_flat(n_pts, radius):
    """radius search over cells with hundreds of points: the sub-cell ranked kernel ("search_sub" = "1", the default choice there)
    returns the rows of the flat warp-per-query kernel bit for bit - indices, distances, counts, total - for k from 1 to 200,
    queries inside, on the border of and outside the cloud, and with fewer than k points in range"""
    import bench
    rng = np.random.default_rng(n_pts)
    cloud = bench.surface_cloud(n_pts, rng, (4.0, 2.5, 3.0))
    grid = api.HashGrid(cloud, np.float32(radius))
    q = np.concatenate([cloud[rng.integers(0, n_pts, 6000)] + rng.uniform(-radius, radius, (6000, 3)).astype(np.float32),
                        rng.uniform(-0.5, 4.5, (1500, 3)).astype(np.float32),      # mostly empty space
                        cloud[:500]]).astype(np.float32)                            # exact hits (distance 0)
    for k in (1, 16, 64, 200):
        for r in (radius, radius * 0.37):
            api.set_option("search_sub", "0")
            a = grid.radius_search(q, float(np.float32(r)), k)
            api.set_option("search_sub", "1")
            b = grid.radius_search(q, float(np.float32(r)), k)
            api.set_option("search_sub", None)
            assert a[3] == b[3] and (a[2] == b[2]).all(), (k, r)
            m = np.arange(k)[None, :] < a[2][:, None]
            assert (a[0][m] == b[0][m]).all() and (a[1][m] == b[1][m]).all(), (k, r)
            assert a[3] > 0
    grid.close()
