"""GPU, at BASELINE.json's full C2 size (200 K-point scan, 36 rotations x 2048 translation seeds per object): properties that
do not need the CPU oracle to run at that size - sortedness and brute-force spot checks of the searches, invariance of the
proposals under sharding / batching / code path, the defining properties of the Poisson-disk levels and of the NMS result -
plus an oracle spot check on a random sample of the dense pose grid."""
import numpy as np
import pytest

from oracle import orcbind as O
from rescan_b200 import api, pipeline, posegrid, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def C2():
    scene, rotations, translations = pipeline.make_workload("C2")
    models = pipeline.upload_objects(scene.objects)
    g1 = api.HashGrid(scene.scan.pos(1), 0.05, normals=scene.scan.nor(1))
    g2 = api.HashGrid(scene.scan.pos(2), 0.05, normals=scene.scan.nor(2))
    return dict(scene=scene, rot=rotations, trans=translations, models=models, g1=g1, g2=g2,
                dyn=[m for m in models if not m.is_static], dyn_obj=[o for o in scene.objects if not o.is_static])


def test_radius_search_rows_sorted_bounded_and_exact_on_a_sample(C2):
    pos = C2["scene"].scan.pos(1)
    rng = np.random.default_rng(2)
    q = np.ascontiguousarray(pos[rng.integers(0, len(pos), 200_000)] + rng.uniform(-0.03, 0.03, (200_000, 3)).astype(np.float32), np.float32)
    for r, k in ((0.10, 64), (0.05, 16), (0.05, 1)):
        idx, d2, nn, _ = C2["g1"].radius_search(q, r, k)
        r2 = np.float32(np.float64(np.float32(r)) * np.float64(np.float32(r)))
        m = np.arange(k)[None, :] < nn[:, None]
        assert (nn <= k).all() and (d2[m] < r2).all() and (idx[m] >= 0).all() and (idx[m] < len(pos)).all()
        dd = np.where(m, d2, np.float32(np.inf))
        with np.errstate(invalid="ignore"):
            assert (np.diff(dd, axis=1)[m[:, 1:]] >= 0).all(), "rows are not ascending"
        # brute force on a sample: the same float32 expression ((vx*vx + vy*vy) + vz*vz), all points of the scan
        for j in rng.integers(0, len(q), 40):
            v = pos - q[j]
            bf = (v[:, 0] * v[:, 0] + v[:, 1] * v[:, 1]) + v[:, 2] * v[:, 2]
            inside = np.sort(bf[bf < r2])
            assert nn[j] == min(k, len(inside)) and (d2[j, : nn[j]] == inside[: nn[j]]).all()


def test_pose_grid_path_equals_explicit_pose_path_and_oracle_on_a_sample(C2):
    rng = np.random.default_rng(5)
    m, o = C2["dyn"][1], C2["dyn_obj"][1]
    grid = api.score_pose_grid(m.levels[4], C2["g1"], C2["rot"], C2["trans"])  # 36 x 2048 scores, never pruned
    assert grid.shape == (len(C2["trans"]), len(C2["rot"])) and np.isfinite(grid).all() and (grid >= 0).all() and (grid <= 1).all()
    t = rng.integers(0, len(C2["trans"]), 512)
    r = rng.integers(0, len(C2["rot"]), 512)
    best = np.argsort(-grid.reshape(-1))[:64]  # make sure high-scoring poses are in the sample
    t[:64], r[:64] = best // len(C2["rot"]), best % len(C2["rot"])
    xf = posegrid.pose_grid(C2["rot"], C2["trans"])[t, r]
    explicit = api.compute_object_alignment_scores(m.levels[4], C2["g1"], xf, 64, 0.10)
    assert (explicit == grid[t, r]).all()
    og = O.OrcGrid(C2["scene"].scan.pos(1), 0.05)
    want, _ = O.score_poses(o.cloud.pos(4), o.cloud.nor(4), og, C2["scene"].scan.nor(1), xf[:128], 64, 0.10, n_threads=8)
    assert np.allclose(explicit[:128], want, rtol=1e-4, atol=1e-7)  # the stated score tolerance


def test_proposals_invariant_under_translation_sharding(C2):
    m = C2["dyn"][2]
    full, ids = api.propose_poses(m.levels[4], m.levels[3], m.levels[2], C2["g1"], C2["rot"], C2["trans"])
    parts, pids = [], []
    for lo, hi in ((0, 700), (700, 1500), (1500, 2048)):
        p, i = api.propose_poses(m.levels[4], m.levels[3], m.levels[2], C2["g1"], C2["rot"], C2["trans"][lo:hi])
        parts.append(p)
        pids.append(i + lo * len(C2["rot"]))
    assert (np.concatenate(pids) == ids).all() and (np.concatenate(parts) == full).all()
    # and the top-k of the whole equals the merge of the parts' top-k (what the ranks exchange)
    top, tid = api.propose_poses(m.levels[4], m.levels[3], m.levels[2], C2["g1"], C2["rot"], C2["trans"], top_k=64)
    loc = [pipeline.merge_topk([p], [i], 64) for p, i in zip(parts, pids)]
    mp_, mi = pipeline.merge_topk([l[0] for l in loc], [l[1] for l in loc], 64)
    assert (mi == tid).all() and (mp_ == top).all()


def test_icp_batch_invariance(C2):
    """one launch over several objects' alignments gives bit for bit what one launch per object gives"""
    rng = np.random.default_rng(9)
    objs, starts = [], []
    for m, o in zip(C2["dyn"][:4], C2["dyn_obj"][:4]):
        objs.append(m.levels[2])
        T = []
        for _ in range(5):
            d = synth.yaw_pose(rng.uniform(-0.08, 0.08), rng.uniform(-0.03, 0.03), rng.uniform(-0.03, 0.03), 0.0)
            T.append(np.ascontiguousarray((o.pose.astype(np.float64) @ d.astype(np.float64)).astype(np.float32).T.reshape(16)))  # perturbed in the object's own frame
        starts.append(np.stack(T))
    ang = np.float32(np.deg2rad(60.0))
    multi = api.icp_align_multi(objs, C2["g2"], [s.copy() for s in starts], 0.10, ang)
    n_close = 0
    for (Tm, em, im), ob, st, o in zip(multi, objs, starts, C2["dyn_obj"][:4]):
        Ts, es, it = api.icp_align(ob, C2["g2"], st.copy(), 0.10, ang)
        assert (Tm == Ts).all() and (em == es).all() and (im == it).all()
        for T in Ts:
            n_close += int(np.linalg.norm(T.reshape(4, 4).T[:3, 3] - o.pose[:3, 3]) < 0.01)
    assert n_close >= 14  # the perturbed starts (<= 3 cm, 0.08 rad) come back to the true placements


def test_poisson_levels_are_first_maximal_independent_sets(C2):
    from scipy.spatial import cKDTree
    p0 = C2["scene"].scan.pos(0)
    tree0 = cKDTree(p0.astype(np.float64))
    for lvl in (1, 2, 3, 4):
        r = float(np.float32(api.LEVEL_VOXEL[lvl]))
        idx = api.poisson_level(p0, lvl)
        assert idx[0] == 0 and (np.diff(idx) > 0).all()
        s = p0[idx].astype(np.float64)
        ts = cKDTree(s)
        # independent: no two samples closer than r; maximal: every point has a sample within r
        assert len(ts.query_pairs(r * (1 - 1e-5))) == 0
        d, _ = ts.query(p0.astype(np.float64), k=1)
        assert (d < r * (1 + 1e-5)).all()
        # "first": a non-sample is always covered by an EARLIER sample
        is_s = np.zeros(len(p0), bool)
        is_s[idx] = True
        rng = np.random.default_rng(lvl)
        for j in rng.choice(np.nonzero(~is_s)[0], 300, replace=False) if (~is_s).any() else []:
            nb = np.array(tree0.query_ball_point(p0[j].astype(np.float64), r * (1 - 1e-5)))
            assert (is_s[nb] & (nb < j)).any()


def test_nms_result_properties_at_top_k(C2):
    m, o = C2["dyn"][0], C2["dyn_obj"][0]
    props, _ = api.propose_poses(m.levels[4], m.levels[3], m.levels[2], C2["g1"], C2["rot"], C2["trans"], top_k=64)
    assert len(props) > 4
    keep = api.non_maxima_suppression(m.levels[3], m.levels[1], m.centroid, props, 0.2)
    kept = props[keep]
    assert keep[np.argmax(props[:, 16])]  # the best proposal always survives
    cen = np.stack([(p[:16].reshape(4, 4).T @ np.append(m.centroid, 1.0))[:3] for p in props])
    kc = cen[keep]
    for a in range(len(kept)):
        ov = api.overlap_factors(m.levels[3], m.levels[1], kept[a, :16], kept[:, :16])
        for b in range(len(kept)):
            if a != b and kept[a, 16] >= kept[b, 16]:
                assert ov[b] <= 0.5 and np.linalg.norm(kc[a] - kc[b]) >= 0.2 - 1e-5 and kept[b, 16] >= 0.01
    # every discarded proposal is suppressed by a kept one that scores at least as high
    for j in np.nonzero(~keep)[0]:
        if props[j, 16] < 0.01:
            continue
        ov = api.overlap_factors(m.levels[3], m.levels[1], props[j, :16], kept[:, :16])
        dist = np.linalg.norm(kc - cen[j], axis=1)
        assert (((ov > 0.5) | (dist < 0.2 + 1e-5)) & (kept[:, 16] >= props[j, 16])).any()
