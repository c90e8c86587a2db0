"""CPU: the oracle (oracle/rescan_oracle.c) against the committed golden vectors written by the compiled reference
(tests/golden/make_golden.py).  This is what pins the oracle on machines without /root/reference."""
import numpy as np
import pytest

from oracle import orcbind as O
from rescan_b200 import posegrid
from tests import common


@pytest.fixture(scope="module")
def G():
    return common.golden()


def test_grid_layout(G):
    z, scan, _ = G
    g = O.OrcGrid(scan.pos(1), 0.05)
    info = g.info()
    assert (info["dims"] == z["grid_dims"]).all()
    assert info["cell_size"] == z["grid_cell"][0] and info["inv_cell_size"] == z["grid_cell"][1]
    assert (np.concatenate([info["min_pt"], info["max_pt"]]) == z["grid_minmax"]).all()
    assert [info["n_pts"], info["n_bins"], info["max_n_pts_in_bin"]] == list(z["grid_counts"])
    assert (g.data()[1] == z["grid_data_idx"]).all()


@pytest.mark.parametrize("tag", ["a", "b", "c", "d"])
def test_radius_search(G, tag):
    z, scan, _ = G
    g = O.OrcGrid(scan.pos(1), 0.05)
    r, k, tot = z[f"rs_{tag}_param"]
    i, d, n, t = g.radius_search(z["queries"], np.float32(r), int(k))
    assert t == int(tot)
    common.check_rows(i, d, n, z[f"rs_{tag}_idx"], z[f"rs_{tag}_d2"], z[f"rs_{tag}_n"], int(k))


def test_knn_search(G):
    z, scan, _ = G
    g = O.OrcGrid(scan.pos(1), 0.05)
    i, d, n, t = g.knn_search(z["knn_queries"], 8)
    common.check_rows(i, d, n, z["knn_idx"], z["knn_d2"], z["knn_n"], 8)


def test_pose_scores(G):
    z, scan, objs = G
    g = O.OrcGrid(scan.pos(1), 0.05)
    got = []
    for x, oi, l in zip(z["score_poses"], z["score_obj"], z["score_lvl"]):
        s, _ = O.score_poses(objs[oi].pos(l), objs[oi].nor(l), g, scan.nor(1), x[None], 32 if l == 1 else 64, 0.10)
        got.append(s[0])
    got = np.array(got, np.float32)
    assert (z["score_ref"] > 0.3).sum() > 10
    assert (got == z["score_ref"]).all(), "oracle scores are expected to be bit-identical to the reference"


def test_propose_poses_reference_grid(G):
    """mgs_propose_poses on the reference's own lattice: 0.10 m spacing over the scan bbox, 10 rotations"""
    z, scan, objs = G
    g = O.OrcGrid(scan.pos(1), 0.05)
    ang = posegrid.rotation_angles(angle_delta=np.float32(posegrid.MSH_TWO_PI / np.float32(10.0)))
    rots = posegrid.rotation_xforms(angle_delta=np.float32(posegrid.MSH_TWO_PI / np.float32(10.0)))
    trans = posegrid.reference_translation_grid(z["scan_bbox"][:3], z["scan_bbox"][3:], 0.10)
    xf = posegrid.pose_grid(rots, trans)
    off = 0
    for i, cloud in enumerate(objs):
        cnt = int(z["propose_counts"][i])
        want = z["propose_flat"][off:off + cnt]
        off += cnt
        if int(z[f"obj{i}_meta"][2]):
            assert cnt == 0  # static classes are skipped (pose_proposal.cpp:198)
            continue
        sc, _ = O.score_poses(cloud.pos(4), cloud.nor(4), g, scan.nor(1), xf, 64, 0.10, n_threads=8)
        ot, orr, osc = O.select_proposals(sc.reshape(len(trans), len(ang)), 0.25)
        got = []
        for t, r, s in zip(ot, orr, osc):
            x = xf[t, r]
            for lvl, thr in ((3, 0.35), (2, 0.40)):
                if s > 0:
                    v, _ = O.score_poses(cloud.pos(lvl), cloud.nor(lvl), g, scan.nor(1), x[None], 64, 0.10)
                    s = v[0] if v[0] > thr else np.float32(-1.0)
            got.append(np.concatenate([x, [s]]).astype(np.float32))
        got = np.stack(got) if got else np.zeros((0, 17), np.float32)
        assert got.shape == want.shape and (got == want).all()


def test_icp_align(G):
    z, scan, objs = G
    for s, oi, end, err in zip(z["icp_start"], z["icp_obj"], z["icp_end"], z["icp_err"]):
        T, e, it = O.icp_align(objs[oi].pos(2), objs[oi].nor(2), scan.pos(2), scan.nor(2), s, 0.10, np.float32(np.deg2rad(60.0)))
        assert (T == end).all() and np.float32(e) == err


def test_labels_and_unary_terms(G):
    z, scan, objs = G
    n = len(objs)
    meta = [z[f"obj{i}_meta"] for i in range(n)]
    order = sorted(range(n), key=lambda i: (int(meta[i][2]), int(meta[i][1])))  # dynamic first (rs_pointcloud_filters.cpp:725-736)
    poses = np.stack([z[f"obj{i}_pose"] for i in order])
    grids = [O.OrcGrid(objs[i].pos(1), 0.05) for i in order]
    nors = [objs[i].nor(1) for i in order]
    n_dyn = sum(1 for i in order if not int(meta[i][2]))
    V = scan.n(1)
    lab, mind = np.zeros(V, np.int8), np.full(V, 1e9, np.float32)
    O.assign_labels(scan.pos(1), scan.nor(1), poses, grids, nors, 0, n_dyn, 0.05, lab, mind)
    O.assign_labels(scan.pos(1), scan.nor(1), poses, grids, nors, n_dyn, n, 0.075, lab, mind)
    inst = np.array([1024] + [int(meta[i][0]) for i in order], np.int32)[lab]
    assert (inst == z["label_instance"]).all()
    # data_cost of rspf_smooth_labels (:926-939) from the reference's own initial labels
    L = int(z["gco_n_labels"][0])
    labels = z["gco_init_labels"]
    is_static = np.zeros(L, np.uint8)
    for i in range(n):
        if int(meta[i][2]):
            is_static[int(meta[i][0]) + 1] = 1
    dc = O.unary_costs(labels, is_static, L)
    assert (dc[:3000] == z["gco_data_cost_head"]).all() and (dc.sum(axis=1) == z["gco_data_cost_rowsum"]).all()


def test_neighborhood_edges(G):
    z, scan, _ = G
    g = O.OrcGrid(scan.pos(1), 0.05)
    nbr, w = O.neighborhood(g, scan.pos(1), scan.nor(1))
    V = scan.n(1)
    cand = {}
    for i in range(3000):
        for j, ww in zip(nbr[i], w[i]):
            if j >= 0:
                cand[(i, int(j))] = ww
    # every reference edge touching the first 3000 vertices is one of the oracle's candidate edges with the same weight
    # (the reference de-duplicates through an int32 key that overflows, so only containment is checked)
    hits = 0
    for a, b, ww in zip(z["edges_a"], z["edges_b"], z["edges_w"]):
        if a < 3000:
            assert (int(a), int(b)) in cand
            assert abs(cand[(int(a), int(b))] - ww) <= 2e-6 * max(abs(ww), 1e-3)
            hits += 1
    assert hits > 3000


def test_overlap_and_nms_match_golden():
    """tests/golden/nms_golden.npz (written from the compiled reference by make_golden_nms.py)"""
    import os
    z, scan, objs = common.golden()
    g = dict(np.load(os.path.join(os.path.dirname(common.GOLDEN), "nms_golden.npz")))
    for i, o in enumerate(objs):
        props = g[f"nms{i}_props"]
        c = O.centroid(o.pos(0))
        assert (c == g[f"nms{i}_centroid"]).all()
        got = np.array([O.overlap_factor(o.pos(3), o.pos(1), props[0, :16], props[j, :16], 0.1, 1, 0) for j in range(len(props))], np.float32)
        assert (got == g[f"nms{i}_overlap"]).all()
        gotb = np.array([O.overlap_factor(o.pos(3), o.pos(1), props[0, :16], props[j, :16], 0.1, 0, 1) for j in range(16)], np.float32)
        assert (gotb == g[f"nms{i}_overlap_boundary"]).all()
        keep = O.nms(o.pos(3), o.pos(1), c, props, 0.2)
        assert (props[keep] == g[f"nms{i}_kept"]).all() and len(g[f"nms{i}_kept"]) == keep.sum()


def test_poisson_levels_match_golden():
    """rs_pointcloud__compute_level_poisson (rs_pointcloud.h:984-1037): the oracle's sample indices of levels 1-4 against
    the reference's (tests/golden/levels_golden.npz, written by make_golden_levels.py from oracle/_ref)"""
    import os
    g = np.load(os.path.join(os.path.dirname(common.GOLDEN), "levels_golden.npz"))
    for name in ("a", "b"):
        pos = g[f"{name}_pos0"]
        for lvl in range(1, 5):
            assert (O.poisson_level(pos, lvl) == g[f"{name}_idx{lvl}"]).all(), (name, lvl)


def test_coverage_grids_match_golden():
    """coverage term: the oracle's grid set-up and rasterisation against tests/golden/coverage_golden.npz
    (written from the reference's grid primitives by make_golden_coverage.py)"""
    import os
    g = np.load(os.path.join(os.path.dirname(common.GOLDEN), "coverage_golden.npz"))
    for voxel, tag in ((0.05, "v5"), (0.15, "v15")):
        res, origin, n = O.cov_grid(g["bbox_min"], g["bbox_max"], voxel)
        assert (res == g[f"{tag}_res"]).all() and (origin == g[f"{tag}_origin"]).all()
        lit = np.nonzero(O.cov_rasterize(g["scan_pos2"], None, res, origin, voxel))[0]
        assert (lit == g[f"{tag}_scan_lit"]).all()
    res, origin, n = O.cov_grid(g["bbox_min"], g["bbox_max"], 0.05)
    for k in range(int(g["n_poses"][0])):
        pts = g[f"obj{int(g[f'pose{k}_obj'][0])}_pos2"]
        lit = np.nonzero(O.cov_rasterize(pts, g[f"pose{k}"], res, origin, 0.05))[0]
        assert (lit == g[f"pose{k}_lit"]).all(), k


def test_plane_inlier_counts_match_golden():
    """evaluate_plane_model of the reference (tests/golden/make_golden_planes.py) incl. a NaN and a zero normal"""
    import os
    g = np.load(os.path.join(os.path.dirname(common.GOLDEN), "planes_golden.npz"))
    for thr in (0.033, 0.05):
        got = O.plane_inlier_counts(g["pts"], g["weights"] > 0.01, g["planes"], thr)
        assert (got == g[f"counts_{int(thr * 1000)}"]).all()
    assert g["counts_33"][7] == 0 and g["counts_33"][11] == int((g["weights"] > 0.01).sum())
