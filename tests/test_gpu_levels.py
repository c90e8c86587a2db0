"""Level building on the GPU: rsgpu_poisson_level against the reference's golden level indices and the
CPU oracle's restatement of rs_pointcloud__compute_level_poisson (lib/rs/rs_pointcloud.h:984-1037).  Integer work:
the sample indices must be identical."""
import os

import numpy as np
import pytest

from oracle import orcbind as O
from rescan_b200 import api, synth
from tests import common

pytestmark = pytest.mark.gpu


def test_levels_match_reference_golden():
    g = np.load(os.path.join(os.path.dirname(common.GOLDEN), "levels_golden.npz"))
    for name in ("a", "b"):
        pos = g[f"{name}_pos0"]
        for lvl in range(1, 5):
            idx, rounds = api.poisson_level(pos, lvl, return_rounds=True)
            assert (idx == g[f"{name}_idx{lvl}"]).all(), (name, lvl)
            assert rounds >= 1


@pytest.mark.parametrize("which", ["tiny_scan", "small_scan", "object"])
def test_levels_match_oracle(which):
    scene = common.small_scene() if which == "small_scan" else common.tiny_scene()
    pos0 = scene.objects[0].cloud.pos(0) if which == "object" else scene.scan.pos(0)
    for lvl in range(1, 5):
        assert (api.poisson_level(pos0, lvl) == O.poisson_level(pos0, lvl)).all(), lvl


def test_compute_levels_rows():
    scene = common.tiny_scene()
    p0, n0 = scene.scan.pos(0), scene.scan.nor(0)
    lv = api.compute_levels(p0, n0)
    assert len(lv) == 5 and lv[0][0].shape == p0.shape
    for lvl in range(1, 5):
        idx = O.poisson_level(p0, lvl)
        assert (lv[lvl][0] == p0[idx]).all() and (lv[lvl][1] == n0[idx]).all()
        assert len(lv[lvl][0]) <= len(lv[lvl - 1][0])


def test_edge_cases():
    assert len(api.poisson_level(np.zeros((0, 3), np.float32), 1)) == 0
    assert list(api.poisson_level(np.array([[1, 2, 3]], np.float32), 3)) == [0]
    # exact duplicates: distance 0 < r^2, the later copy is marked by the earlier one
    p = np.array([[0, 0, 0], [0, 0, 0], [1, 0, 0], [0, 0, 0], [1, 0, 0.005]], np.float32)
    assert list(api.poisson_level(p, 1)) == [0, 2] == list(O.poisson_level(p, 1))
    # a chain along a line: every decision depends on the previous one (spacing 0.6 r)
    line = np.zeros((4000, 3), np.float32)
    line[:, 0] = np.arange(4000) * np.float32(0.006)
    idx, rounds = api.poisson_level(line, 1, return_rounds=True)
    assert (idx == O.poisson_level(line, 1)).all() and rounds > 1000
    # random order of the same line
    perm = np.random.default_rng(1).permutation(4000)
    assert (api.poisson_level(line[perm], 1) == O.poisson_level(line[perm], 1)).all()


def test_k_cap_fails_loudly():
    """a ball holding more points than the reference's max_n_neigh is outside what the propagation models"""
    rng = np.random.default_rng(8)
    u, v = np.meshgrid(np.arange(0, 0.32, 0.0035), np.arange(0, 0.32, 0.0035), indexing="ij")
    dense = (np.stack([u.reshape(-1), np.zeros(u.size), v.reshape(-1)], axis=1) + rng.normal(0, 0.0004, (u.size, 3))).astype(np.float32)
    assert (api.poisson_level(dense, 1) == O.poisson_level(dense, 1)).all()  # level 1: 26 points per ball, fine
    with pytest.raises(api.RsgpuError) as e:
        api.poisson_level(dense, 4)
    assert e.value.code == -4
