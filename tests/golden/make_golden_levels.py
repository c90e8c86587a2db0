"""Writes tests/golden/levels_golden.npz from the UNMODIFIED reference compiled in place (oracle/_ref): the level-0
indices of the points rs_pointcloud_compute_levels (lib/rs/rs_pointcloud.h:1305, Poisson-disk sampling :984-1037)
puts into levels 1-4 of two small clouds.  Run in the build container (needs /root/reference):

    python tests/golden/make_golden_levels.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import refbind as R  # noqa: E402


def clouds():
    rng = np.random.default_rng(20191027)
    # (a) a floor patch and a box on it, jittered 6 mm lattice in row-major order (dense enough that level 1 thins it)
    def plane(o, eu, ev, lu, lv, s):
        u, v = np.meshgrid(np.arange(0, lu, s), np.arange(0, lv, s), indexing="ij")
        p = o + u.reshape(-1, 1) * eu + v.reshape(-1, 1) * ev
        return p + rng.normal(0, 0.0007, p.shape)
    ex, ey, ez = np.eye(3)
    a = np.concatenate([plane(np.zeros(3), ex, ez, 0.9, 0.7, 0.006),
                        plane(np.array([0.3, 0.0, 0.2]), ex, ey, 0.3, 0.25, 0.006),
                        plane(np.array([0.3, 0.25, 0.2]), ex, ez, 0.3, 0.3, 0.006)]).astype(np.float32)
    # (b) the same kind of surface in random order plus a sparse volume cloud (isolated points, empty windows)
    b = np.concatenate([plane(np.zeros(3), ex, ez, 0.5, 0.5, 0.007), rng.uniform(0, 0.6, (1500, 3))]).astype(np.float32)
    b = b[rng.permutation(len(b))]
    return {"a": np.ascontiguousarray(a), "b": np.ascontiguousarray(b)}


def main():
    out = {}
    for name, pos in clouds().items():
        nor = np.tile(np.array([0, 1, 0], np.float32), (len(pos), 1))
        rc = R.RefCloud.from_level0(pos, nor)
        out[f"{name}_pos0"] = pos
        key = {p.tobytes(): i for i, p in enumerate(pos)}
        assert len(key) == len(pos), "duplicate points: rows cannot be matched back to indices"
        for lvl in range(1, 5):
            lp, _ = rc.level(lvl)
            idx = np.array([key[p.tobytes()] for p in lp], np.int32)
            assert (np.diff(idx) > 0).all()
            out[f"{name}_idx{lvl}"] = idx
            print(name, "level", lvl, len(idx), "of", len(pos))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "levels_golden.npz"), **out)


if __name__ == "__main__":
    main()
