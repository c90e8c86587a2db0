"""GPU: the reference's UNMODIFIED apps/pose_proposal/main.cpp linked against the rsgpu drop-in
(integration/_build/pose_proposal_rsgpu, built by integration/Makefile where /root/reference exists) is run on a
synthetic database + rescan and its proposal .bin is compared with the one the pure-CPU reference build wrote for the
same files (tests/golden/dropin_pp.bin, written by integration/make_dropin_case.py --golden)."""
import os
import subprocess

import numpy as np
import pytest

from rescan_b200 import rsio
from tests import common

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "integration", "_build", "pose_proposal_rsgpu")


@pytest.mark.skipif(not os.path.exists(EXE), reason="integration/_build/pose_proposal_rsgpu not built (needs /root/reference at build time)")
def test_dropin_executable_matches_cpu_reference(tmp_path):
    import sys
    sys.path.insert(0, os.path.join(ROOT, "integration"))
    import make_dropin_case
    db, scan, out, _ = make_dropin_case.write_case(str(tmp_path))
    r = subprocess.run([EXE, db, scan, out, "-v"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "(GPU)" in r.stdout
    got = rsio.read_proposals(os.path.join(str(tmp_path), "scan1_pp", "scan1_pp.bin"))
    want = rsio.read_proposals(os.path.join(ROOT, "tests", "golden", "dropin_pp.bin"))
    assert [len(g) for g in got] == [len(w) for w in want]
    for g, w in zip(got, want):
        for a, b in zip(g, w):
            A, B = a[:16].reshape(4, 4).T.astype(np.float64), b[:16].reshape(4, 4).T.astype(np.float64)
            R = A[:3, :3] @ B[:3, :3].T
            assert np.linalg.norm(A[:3, 3] - B[:3, 3]) < 1e-5 and np.linalg.norm(R - R.T) / (2 * np.sqrt(2)) < 1e-5
            assert abs(a[16] - b[16]) <= 1e-4 * max(abs(b[16]), 1e-3)


ST_EXE = os.path.join(ROOT, "integration", "_build", "segment_transfer_rsgpu")


@pytest.mark.skipif(not (os.path.exists(EXE) and os.path.exists(ST_EXE)),
                    reason="integration/_build/segment_transfer_rsgpu not built (needs /root/reference at build time)")
def test_segment_transfer_dropin_matches_cpu_reference(tmp_path):
    """The reference's UNMODIFIED apps/segment_transfer sources linked against integration/rsgpu_dropin_st.cpp: scan
    rasterisation, coverage term of the greedy / simulated-annealing optimiser, ICP refinement, label transfer, unary terms
    and neighbourhood weights on the GPU.  Input = the database the GPU pose_proposal drop-in writes for the synthetic case,
    with the proposal .bin of the pure-CPU reference (tests/golden/dropin_pp.bin) so that both optimisers start from the same
    bytes; expected = what the pure-CPU segment_transfer build decided (tests/golden/dropin_st.npz, written by
    integration/make_dropin_case.py --golden): the same placements with poses within 1e-5 m / 1e-5 rad and identical
    per-vertex class / instance labels."""
    import shutil
    import sys
    sys.path.insert(0, os.path.join(ROOT, "integration"))
    import make_dropin_case
    db, scan, out, _ = make_dropin_case.write_case(str(tmp_path))
    r = subprocess.run([EXE, db, scan, out, "-v"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    shutil.copy(os.path.join(ROOT, "tests", "golden", "dropin_pp.bin"), os.path.join(str(tmp_path), "scan1_pp", "scan1_pp.bin"))
    stdout, rows, ply = make_dropin_case.run_segment_transfer(ST_EXE, out, str(tmp_path))
    assert "(GPU)" in stdout
    g = np.load(os.path.join(ROOT, "tests", "golden", "dropin_st.npz"))
    assert [r[0] for r in rows] == list(g["placement_uidx"]) and [r[2] for r in rows] == list(g["object_idx"])
    for r, score, pose in zip(rows, g["score"], g["pose"]):
        A, B = r[4].astype(np.float64), pose.astype(np.float64)
        R = A[:3, :3] @ B[:3, :3].T
        assert np.linalg.norm(A[:3, 3] - B[:3, 3]) < 1e-5 and np.linalg.norm(R - R.T) / (2 * np.sqrt(2)) < 1e-5
        assert abs(r[3] - score) <= 1e-4 * max(abs(score), 1e-3)
    assert len(ply) == len(g["x"]) and (np.asarray(ply["x"], np.float32) == g["x"]).all()
    assert (np.asarray(ply["class_idx"], np.int32) == g["class_idx"]).all()
    assert (np.asarray(ply["instance_idx"], np.int32) == g["instance_idx"]).all()
