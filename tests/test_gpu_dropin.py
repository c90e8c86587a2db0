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
