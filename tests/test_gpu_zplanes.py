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


ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PP_LEVELS = os.path.join(ROOT, "integration", "_build", "pose_proposal_rsgpu_levels")
ST_ALL = os.path.join(ROOT, "integration", "_build", "segment_transfer_rsgpu_all")


@pytest.mark.skipif(not (os.path.exists(PP_LEVELS) and os.path.exists(ST_ALL)),
                    reason="integration/_build/*_levels / *_all not built (needs /root/reference at build time)")
def test_dropin_executables_with_levels_and_planes_match_cpu_reference(tmp_path):
    """the drop-in executables with ALL optional shim objects linked (level building at its call site, plane detector rounds as
    one count call each, integration/rsgpu_dropin_levels.cpp / rsgpu_dropin_planes.cpp) against the same goldens as
    tests/test_gpu_dropin.py: the proposal .bin of the pure-CPU pose_proposal and the decisions of the pure-CPU segment_transfer"""
    import shutil
    import subprocess
    import sys
    from rescan_b200 import rsio
    sys.path.insert(0, os.path.join(ROOT, "integration"))
    import make_dropin_case
    db, scan, out, _ = make_dropin_case.write_case(str(tmp_path))
    r = subprocess.run([PP_LEVELS, db, scan, out, "-v"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "selected by rsgpu_poisson_level" in r.stdout
    got = rsio.read_proposals(os.path.join(str(tmp_path), "scan1_pp", "scan1_pp.bin"))
    want = rsio.read_proposals(os.path.join(ROOT, "tests", "golden", "dropin_pp.bin"))
    assert [len(g) for g in got] == [len(w) for w in want]
    for g, w in zip(got, want):
        for a, b in zip(g, w):
            assert np.abs(a[:16] - b[:16]).max() < 1e-5 and abs(a[16] - b[16]) <= 1e-4 * max(abs(b[16]), 1e-3)
    shutil.copy(os.path.join(ROOT, "tests", "golden", "dropin_pp.bin"), os.path.join(str(tmp_path), "scan1_pp", "scan1_pp.bin"))
    stdout, rows, ply = make_dropin_case.run_segment_transfer(ST_ALL, out, str(tmp_path))
    assert "candidates counted in one call each" in stdout and "selected by rsgpu_poisson_level" in stdout
    g = np.load(os.path.join(ROOT, "tests", "golden", "dropin_st.npz"))
    assert [r[0] for r in rows] == list(g["placement_uidx"]) and [r[2] for r in rows] == list(g["object_idx"])
    for r, pose in zip(rows, g["pose"]):
        assert np.abs(r[4][:3, 3] - pose[:3, 3]).max() < 1e-5 and np.abs(r[4][:3, :3] - pose[:3, :3]).max() < 1e-5
    assert len(ply) == len(g["x"]) and (np.asarray(ply["x"], np.float32) == g["x"]).all()
    assert (np.asarray(ply["class_idx"], np.int32) == g["class_idx"]).all()
    assert (np.asarray(ply["instance_idx"], np.int32) == g["instance_idx"]).all()
