"""Coverage term of the arrangement optimiser on the GPU: rasterisation and per-placement bit masks against
the CPU oracle (which is pinned to the reference's grid primitives).  Integer work: grids, masks and counts identical; the
score is the same float division."""
import numpy as np
import pytest

from oracle import orcbind as O
from rescan_b200 import api, synth
from tests import common

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("which,voxel", [("tiny", 0.05), ("small", 0.05), ("tiny", 0.15)])
def test_rasterisation_and_masks_match_oracle(which, voxel):
    scene = common.small_scene() if which == "small" else common.tiny_scene()
    mn, mx = scene.scan.pos(0).min(0), scene.scan.pos(0).max(0)
    res, origin = api.coverage_grid(mn, mx, voxel)
    ro, oo, n_cells = O.cov_grid(mn, mx, voxel)
    assert (res == ro).all() and (origin == oo).all()
    # the scan: rsao_rasterize_scene_to_grid
    scn = api.rasterize_points(scene.scan.pos(2), None, res, origin, voxel)
    want = O.cov_rasterize(scene.scan.pos(2), None, ro, oo, voxel)
    assert (scn == want).all() and len(scn) == n_cells
    # candidate placements: true poses, perturbed poses, poses that push part of the object out of the grid
    rng = np.random.default_rng(5)
    clouds, poses, owners = [], [], []
    for oi, o in enumerate(scene.objects):
        c2 = api.PointCloud(o.cloud.pos(2), o.cloud.nor(2))
        for j in range(6):
            d = np.eye(4, dtype=np.float32) if j == 0 else synth.yaw_pose(rng.uniform(0, 6.28), rng.uniform(-2, 2), rng.uniform(-2, 2), rng.uniform(-0.3, 0.3))
            clouds.append(c2)
            owners.append(oi)
            poses.append(common.colmajor((d.astype(np.float64) @ o.pose.astype(np.float64)).astype(np.float32)))
    masks, n_lit = api.coverage_masks(clouds, np.stack(poses), res, origin, scn, voxel)
    lit = np.nonzero(want)[0]
    assert n_lit == len(lit) and masks.shape == (len(poses), (n_lit + 31) // 32)
    grids = []
    for k, (oi, pose) in enumerate(zip(owners, poses)):
        g = O.cov_rasterize(scene.objects[oi].cloud.pos(2), pose, ro, oo, voxel)
        grids.append(g)
        bits = np.unpackbits(masks[k].view(np.uint8), bitorder="little")[:n_lit].astype(bool)
        assert (bits == (g[lit] > 0)).all(), k
        # the posed-object rasteriser is the same entry point with a pose
        assert (api.rasterize_points(scene.objects[oi].cloud.pos(2), pose, res, origin, voxel) == g).all()
    # coverage of random arrangements: popcount( OR of masks ) / n_lit == the reference's counting loop on the OR-ed grids
    for _ in range(20):
        pick = rng.choice(len(poses), size=rng.integers(1, 6), replace=False)
        arr = np.zeros(n_cells, np.uint8)
        for k in pick:
            arr |= grids[k]
        assert api.coverage_score(masks[pick], n_lit) == np.float32(O.cov_score(want, arr))


def test_coverage_edge_cases():
    res, origin = api.coverage_grid(np.zeros(3, np.float32), np.ones(3, np.float32), 0.05)
    empty = np.zeros(int(res[0]) * int(res[1]) * int(res[2]), np.uint8)
    c = api.PointCloud(np.full((5, 3), 0.5, np.float32), np.tile(np.array([0, 1, 0], np.float32), (5, 1)))
    eye = np.eye(4, dtype=np.float32).reshape(1, 16)
    masks, n_lit = api.coverage_masks([c], eye, res, origin, empty)  # nothing lit in the scan: no bits, score 0
    assert n_lit == 0 and masks.shape == (1, 0) and api.coverage_score(masks, n_lit) == 0
    g = api.rasterize_points(np.array([[0.5, 0.5, 0.5], [99.0, 0.5, 0.5], [-99.0, 0.5, 0.5]], np.float32), None, res, origin)
    assert g.sum() == 1  # points outside the grid are ignored
    masks, n_lit = api.coverage_masks([c], eye, res, origin, g)
    assert n_lit == 1 and masks[0, 0] == 1
    assert (api.rasterize_points(np.zeros((0, 3), np.float32), None, res, origin) == 0).all()


def test_masks_match_reference_golden():
    import os
    g = np.load(os.path.join(os.path.dirname(common.GOLDEN), "coverage_golden.npz"))
    res, origin = api.coverage_grid(g["bbox_min"], g["bbox_max"], 0.05)
    assert (res == g["v5_res"]).all() and (origin == g["v5_origin"]).all()
    scn = api.rasterize_points(g["scan_pos2"], None, res, origin, 0.05)
    lit = g["v5_scan_lit"]
    assert (np.nonzero(scn)[0] == lit).all()
    n = int(g["n_poses"][0])
    clouds = {}
    objs, poses = [], []
    for k in range(n):
        oi = int(g[f"pose{k}_obj"][0])
        if oi not in clouds:
            p = g[f"obj{oi}_pos2"]
            clouds[oi] = api.PointCloud(p, np.tile(np.array([0, 1, 0], np.float32), (len(p), 1)))
        objs.append(clouds[oi])
        poses.append(g[f"pose{k}"])
    masks, n_lit = api.coverage_masks(objs, np.stack(poses), res, origin, scn, 0.05)
    assert n_lit == len(lit)
    for k in range(n):
        bits = np.unpackbits(masks[k].view(np.uint8), bitorder="little")[:n_lit].astype(bool)
        assert (lit[bits] == np.intersect1d(g[f"pose{k}_lit"], lit)).all(), k
