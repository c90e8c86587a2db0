"""Writes the file set of the drop-in test (a first-scan database + a rescan PLY) into a directory.  With --golden it
also runs the pure-CPU reference build (integration/_build/pose_proposal_ref) on it and stores the resulting proposal
.bin as tests/golden/dropin_pp.bin, together with the .rsdb and one object PLY the reference wrote (build container only;
the .rsdb holds the absolute paths of the folder it was written in, default /tmp/rsgpu_dropin_case), then runs the pure-CPU
reference build of segment_transfer (integration/_build/segment_transfer_ref, gco replaced by the pass-through stand-in) on
that output and stores what it decided as tests/golden/dropin_st.npz: the optimised and refined arrangement and the
per-vertex class / instance labels of the level-1 scan."""
import os
import shutil
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rescan_b200 import rsio, synth  # noqa: E402


SMALL = dict(n_objects=3, n_static=1, room=(3.2, 1.6, 2.8), spacing=0.035)


def write_case(folder, **scene):
    """first scan + database and a rescan of the same objects; `scene` = synth.make_scene arguments (default: the small test case)"""
    os.makedirs(folder, exist_ok=True)
    scene = scene or SMALL
    scan0 = synth.make_scene(seed=synth.SEED + 99, **scene)
    scan1 = synth.make_scene(seed=synth.SEED + 100, objects=scan0.objects, **scene)
    p0 = os.path.join(folder, "scan0.ply")
    p1 = os.path.join(folder, "scan1.ply")
    rsio.write_ply(p0, scan0.scan.pos(0), scan0.scan.nor(0), scan0.scan_class, scan0.scan_instance)
    rsio.write_ply(p1, scan1.scan.pos(0), scan1.scan.nor(0), scan1.scan_class, scan1.scan_instance)
    db = rsio.write_database(folder, "scan0", scan0, p0, [(i, o.pose) for i, o in enumerate(scan0.objects)])
    return db, p1, os.path.join(folder, "scan1_pp.rsdb"), scan1


def write_pose_proposal_output(folder, db, scan, proposals_bin):
    """What `pose_proposal <db> <scan> <folder>/scan1_pp.rsdb` leaves behind, without running it: the first-scan database with
    the rescan appended as scene 1 (rs_database.h:539-611 writes the same lines) and a proposal .bin (copied from
    `proposals_bin`).  The object models stay where the first-scan database has them.  Returns the .rsdb path."""
    out = os.path.join(folder, "scan1_pp.rsdb")
    pp = os.path.join(folder, "scan1_pp")
    os.makedirs(pp, exist_ok=True)
    text = []
    for ln in open(db).read().splitlines():
        text.append(ln)
        if ln.startswith("scene 0"):
            text.append(f"scene 1 1 {scan} {os.path.join(pp, 'scan1_pp.bin')} ")
    with open(out, "w") as f:
        f.write("\n".join(text) + "\n")
    shutil.copy(proposals_bin, os.path.join(pp, "scan1_pp.bin"))
    return out


def run_segment_transfer(exe, pp_rsdb, folder):
    """`segment_transfer <pp.rsdb> -o <folder>/out/scan1_st.rsdb` -> (stdout, arrangement rows of the last scene, labelled level-1 PLY)"""
    out_dir = os.path.join(folder, "out")
    os.makedirs(out_dir, exist_ok=True)
    out = os.path.join(out_dir, "scan1_st.rsdb")
    r = subprocess.run([exe, pp_rsdb, "-o", out], capture_output=True, text=True, timeout=1200)
    if r.returncode != 0:
        raise RuntimeError(f"{exe} failed ({r.returncode}): {r.stdout[-2000:]}{r.stderr[-2000:]}")
    db = rsio.read_database(out)
    last = max(p[1] for p in db["poses"])
    rows = [p for p in db["poses"] if p[1] == last]
    ply = rsio.read_ply(os.path.join(out_dir, "predictions", "scan1_st.ply"))
    return r.stdout, rows, ply


def main():
    folder = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("--") else "/tmp/rsgpu_dropin_case"
    db, scan, out, _ = write_case(folder)
    print(db, scan, out)
    if "--golden" in sys.argv:
        exe = os.path.join(ROOT, "integration", "_build", "pose_proposal_ref")
        subprocess.check_call([exe, db, scan, out, "-v"], stdout=subprocess.DEVNULL)
        gold = os.path.join(ROOT, "tests", "golden")
        shutil.copy(os.path.join(folder, "scan1_pp", "scan1_pp.bin"), os.path.join(gold, "dropin_pp.bin"))
        # the database and one object model exactly as the reference wrote them: fixtures of the rsio readers (tests/test_host_logic.py)
        shutil.copy(out, os.path.join(gold, "dropin_pp.rsdb"))
        shutil.copy(os.path.join(folder, "scan1_pp", "obj_102.ply"), os.path.join(gold, "dropin_obj_102.ply"))
        print("wrote tests/golden/dropin_pp.bin", os.path.getsize(os.path.join(gold, "dropin_pp.bin")), "bytes, dropin_pp.rsdb, dropin_obj_102.ply")
        exe = os.path.join(ROOT, "integration", "_build", "segment_transfer_ref")
        _, rows, ply = run_segment_transfer(exe, out, folder)
        np.savez_compressed(os.path.join(gold, "dropin_st.npz"),
                            placement_uidx=np.array([r[0] for r in rows], np.int32), object_idx=np.array([r[2] for r in rows], np.int32),
                            score=np.array([r[3] for r in rows], np.float32), pose=np.stack([r[4] for r in rows]).astype(np.float32),
                            class_idx=np.asarray(ply["class_idx"], np.int32), instance_idx=np.asarray(ply["instance_idx"], np.int32),
                            x=np.asarray(ply["x"], np.float32))
        print("wrote tests/golden/dropin_st.npz:", len(rows), "placements,", len(ply), "labelled vertices")


if __name__ == "__main__":
    main()
