"""GPU: the C-ABI path against the committed golden vectors of the compiled reference (tests/golden/), i.e. parity
with the reference itself rather than with the oracle."""
import numpy as np
import pytest

from rescan_b200 import api, posegrid
from tests import common

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def G():
    z, scan, objs = common.golden()
    g1 = api.HashGrid(scan.pos(1), 0.05, normals=scan.nor(1))
    return z, scan, objs, g1


def test_grid_layout(G):
    z, scan, _, g = G
    info = g.info()
    assert (info["dims"] == z["grid_dims"]).all()
    assert info["cell_size"] == z["grid_cell"][0] and info["inv_cell_size"] == z["grid_cell"][1]
    assert (np.concatenate([info["min_pt"], info["max_pt"]]) == z["grid_minmax"]).all()
    assert [info["n_pts"], info["n_bins"], info["max_n_pts_in_bin"]] == list(z["grid_counts"])
    assert (g.data()[1] == z["grid_data_idx"]).all()


@pytest.mark.parametrize("tag", ["a", "b", "c", "d"])
def test_radius_search(G, tag):
    z, _, _, g = G
    r, k, tot = z[f"rs_{tag}_param"]
    i, d, n, t = g.radius_search(z["queries"], np.float32(r), int(k))
    assert t == int(tot)
    common.check_rows(i, d, n, z[f"rs_{tag}_idx"], z[f"rs_{tag}_d2"], z[f"rs_{tag}_n"], int(k))


def test_knn_search(G):
    z, _, _, g = G
    i, d, n, t = g.knn_search(z["knn_queries"], 8)
    common.check_rows(i, d, n, z["knn_idx"], z["knn_d2"], z["knn_n"], 8)


def test_pose_scores(G):
    z, _, objs, g = G
    clouds = {}
    got = []
    for x, oi, l in zip(z["score_poses"], z["score_obj"], z["score_lvl"]):
        c = clouds.setdefault((int(oi), int(l)), api.PointCloud(objs[oi].pos(l), objs[oi].nor(l)))
        got.append(api.compute_object_alignment_scores(c, g, x[None], 32 if l == 1 else 64, 0.10)[0])
    got = np.array(got, np.float32)
    assert np.allclose(got, z["score_ref"], rtol=1e-4, atol=1e-7)  # north_star tolerance: 1e-4 relative
    assert (got == z["score_ref"]).mean() > 0.9


def test_propose_poses_reference_grid(G):
    """the whole of mgs_propose_poses on the reference's own lattice: same survivors, same order, same poses"""
    z, _, objs, g = G
    rots = posegrid.rotation_xforms(angle_delta=np.float32(posegrid.MSH_TWO_PI / np.float32(10.0)))
    trans = posegrid.reference_translation_grid(z["scan_bbox"][:3], z["scan_bbox"][3:], 0.10)
    off = 0
    for i, cloud in enumerate(objs):
        cnt = int(z["propose_counts"][i])
        want = z["propose_flat"][off:off + cnt]
        off += cnt
        if int(z[f"obj{i}_meta"][2]):
            continue
        c4, c3, c2 = (api.PointCloud(cloud.pos(l), cloud.nor(l)) for l in (4, 3, 2))
        got, ids = api.propose_poses(c4, c3, c2, g, rots, trans)
        assert got.shape == want.shape
        assert (got[:, :16] == want[:, :16]).all()
        assert np.allclose(got[:, 16], want[:, 16], rtol=1e-4, atol=1e-7)
        assert ((got[:, 16] < 0) == (want[:, 16] < 0)).all()


def test_icp_align(G):
    z, scan, objs, _ = G
    g2 = api.HashGrid(scan.pos(2), 0.05, normals=scan.nor(2))
    ang = np.float32(np.deg2rad(60.0))
    jobs = sorted(set(int(o) for o in z["icp_obj"]))
    clouds = [api.PointCloud(objs[o].pos(2), objs[o].nor(2)) for o in jobs]
    outs = api.icp_align_multi(clouds, g2, [z["icp_start"][z["icp_obj"] == o] for o in jobs], 0.10, ang)
    for o, (T, err, it) in zip(jobs, outs):
        want_T, want_e = z["icp_end"][z["icp_obj"] == o], z["icp_err"][z["icp_obj"] == o]
        for a, b in zip(T, want_T):
            A, B = a.reshape(4, 4).T.astype(np.float64), b.reshape(4, 4).T.astype(np.float64)
            R = A[:3, :3] @ B[:3, :3].T
            assert np.linalg.norm(A[:3, 3] - B[:3, 3]) < 1e-5 and np.linalg.norm(R - R.T) / (2 * np.sqrt(2)) < 1e-5
        assert np.allclose(err, want_e, rtol=1e-4)
        assert (T == want_T).mean() > 0.9  # reference-order sums: the refined matrices are normally bit-identical


def test_labels_and_unary_terms(G):
    z, scan, objs, _ = G
    n = len(objs)
    meta = [z[f"obj{i}_meta"] for i in range(n)]
    order = sorted(range(n), key=lambda i: (int(meta[i][2]), int(meta[i][1])))
    poses = np.stack([z[f"obj{i}_pose"] for i in order])
    grids = [api.HashGrid(objs[i].pos(1), 0.05, normals=objs[i].nor(1)) for i in order]
    n_dyn = sum(1 for i in order if not int(meta[i][2]))
    V = scan.n(1)
    lab, mind = np.zeros(V, np.int8), np.full(V, 1e9, np.float32)
    api.assign_labels(scan.pos(1), scan.nor(1), poses, grids, 0, n_dyn, 0.05, lab, mind)
    api.assign_labels(scan.pos(1), scan.nor(1), poses, grids, n_dyn, n, 0.075, lab, mind)
    inst = np.array([1024] + [int(meta[i][0]) for i in order], np.int32)[lab]
    assert (inst == z["label_instance"]).all()
    L = int(z["gco_n_labels"][0])
    is_static = np.zeros(L, np.uint8)
    for i in range(n):
        if int(meta[i][2]):
            is_static[int(meta[i][0]) + 1] = 1
    dc = api.unary_costs(z["gco_init_labels"], is_static, L)
    assert (dc[:3000] == z["gco_data_cost_head"]).all() and (dc.sum(axis=1) == z["gco_data_cost_rowsum"]).all()


def test_neighborhood_edges(G):
    z, scan, _, g = G
    nbr, w = api.neighborhood(g, scan.pos(1), scan.nor(1))
    cand = {}
    for i in range(3000):
        for j, ww in zip(nbr[i], w[i]):
            if j >= 0:
                cand[(i, int(j))] = ww
    hits = 0
    for a, b, ww in zip(z["edges_a"], z["edges_b"], z["edges_w"]):
        if a < 3000:
            assert (int(a), int(b)) in cand
            assert abs(cand[(int(a), int(b))] - ww) <= 2e-6 * max(abs(ww), 1e-3)
            hits += 1
    assert hits > 3000
