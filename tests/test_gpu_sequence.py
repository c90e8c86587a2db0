"""C4 in miniature: two rescans of a small room through pose_proposal + the unary path (scripts/run_sequence.py), with the
label transfer and data_cost of every scan checked against the CPU oracle."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_rescan_sequence_small():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "run_sequence.py"), "--scans", "2", "--objects", "6", "--static", "2",
                          "--room", "4,2.2,3.5", "--spacing", "0.03", "--seeds", "128", "--rot", "8", "--check"],
                         capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [json.loads(l) for l in out.stdout.splitlines() if l.startswith("{")]
    scans = [l for l in lines if "scan" in l]
    assert len(scans) == 2
    for s in scans:
        assert s["labels_match_oracle"] and s["data_cost_match_oracle"]
        assert s["labelled_vertices"] > 1000 and s["edges"] > 0
    assert lines[-1]["summary"] == "C4" and lines[-1]["pose_evaluations_per_s"] > 0
