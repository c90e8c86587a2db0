"""GPU: inlier counts of the plane detector's RANSAC rounds (rsgpu_plane_inlier_counts, include/rsgpu.h) against the CPU
oracle and against golden vectors written with the reference's own evaluate_plane_model (tests/golden/planes_golden.npz).
Integer work: counts identical."""
import os

import numpy as np
import pytest

from oracle import orcbind as O
from rescan_b200 import api
from tests import common

pytestmark = pytest.mark.gpu


def _triples(p, rng, n):
    idx = rng.integers(0, len(p), (n, 3))
    a, b, c = p[idx[:, 0]], p[idx[:, 1]], p[idx[:, 2]]
    nr = np.cross((b - a).astype(np.float32), (c - a).astype(np.float32)).astype(np.float32)
    with np.errstate(invalid="ignore", divide="ignore"):
        nr = (nr / np.sqrt((nr * nr).sum(1, keepdims=True), dtype=np.float32)).astype(np.float32)
    return np.concatenate([a, nr], 1).astype(np.float32)


def test_counts_match_reference_golden():
    g = np.load(os.path.join(os.path.dirname(common.GOLDEN), "planes_golden.npz"))
    for thr in (0.033, 0.05):
        got = api.plane_inlier_counts(g["pts"], g["weights"] > 0.01, g["planes"], thr)
        assert (got == g[f"counts_{int(thr * 1000)}"]).all()


@pytest.mark.parametrize("n_planes", [1, 15, 16, 17, 5000])
def test_counts_match_oracle(n_planes):
    scene = common.small_scene()
    p, n = scene.scan.pos(2), scene.scan.nor(2)
    rng = np.random.default_rng(n_planes)
    active = np.abs(n[:, 1]) < 0.2
    active &= rng.random(len(p)) > 0.2
    planes = _triples(p, rng, n_planes)
    got = api.plane_inlier_counts(p, active, planes, 0.033)
    want = O.plane_inlier_counts(p, active, planes, 0.033) if n_planes <= 64 else None
    if want is None:  # a round of the reference's size: the oracle on a sample of the candidates
        pick = rng.choice(n_planes, 64, replace=False)
        assert (got[pick] == O.plane_inlier_counts(p, active, planes[pick], 0.033)).all()
    else:
        assert (got == want).all()


def test_edge_cases():
    p = np.array([[0, 0, 0], [0, 0.01, 0], [0, 1, 0], [5, 0.02, 5]], np.float32)
    planes = np.array([[0, 0, 0, 0, 1, 0], [0, 0, 0, np.nan, np.nan, np.nan], [0, 0, 0, 0, 0, 0]], np.float32)
    assert list(api.plane_inlier_counts(p, np.ones(4, bool), planes, 0.033)) == [3, 0, 4]  # NaN normal counts nothing, zero normal everything
    assert list(api.plane_inlier_counts(p, np.array([1, 0, 1, 0], bool), planes, 0.033)) == [1, 0, 2]  # inactive points are skipped
    assert list(api.plane_inlier_counts(np.zeros((0, 3), np.float32), np.zeros(0, bool), planes, 0.033)) == [0, 0, 0]
    assert len(api.plane_inlier_counts(p, np.ones(4, bool), np.zeros((0, 6), np.float32), 0.033)) == 0
