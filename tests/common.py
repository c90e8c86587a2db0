"""Shared scene builders for the tests (CPU side: synthetic data + oracle handles)."""
from __future__ import annotations

import functools
import numpy as np

from oracle import orcbind as O
from rescan_b200 import synth


@functools.lru_cache(maxsize=None)
def tiny_scene():
    return synth.make_scene(**synth.CONFIGS["tiny"]["scene"])


@functools.lru_cache(maxsize=None)
def small_scene():
    """denser than `tiny` (1.5 cm lattice -> level 1 at ~1 cm spacing is hit, bins hold tens of points)"""
    return synth.make_scene(n_objects=4, n_static=1, room=(3.0, 2.2, 2.5), spacing=0.012, seed=synth.SEED + 7)


def perturbed_poses(rng, scene, n_per_object=6, dt=0.04, dyaw=0.1):
    out = []
    for o in scene.objects:
        for _ in range(n_per_object):
            d = synth.yaw_pose(rng.uniform(-dyaw, dyaw), rng.uniform(-dt, dt), rng.uniform(-dt, dt), rng.uniform(-0.01, 0.01))
            out.append((o, (d.astype(np.float64) @ o.pose.astype(np.float64)).astype(np.float32)))
    return out


def colmajor(m4):
    return np.ascontiguousarray(np.asarray(m4, np.float32).T.reshape(16))


def rotation_xforms(n_rot):
    """the reference's rotation set: msh_rotate(I, y_angle, +Y) with float-accumulated angles"""
    ang = synth.rotation_angles(n_rot)
    return np.stack([O.make_pose(a, 0, 0, 0) for a in ang]).astype(np.float32), ang


# ------------------------------------------------------------------------------------------------ golden fixture
import os

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "rescan_golden.npz")


@functools.lru_cache(maxsize=None)
def golden():
    """reference outputs written by tests/golden/make_golden.py + the inputs rebuilt from the stored level 0"""
    z = dict(np.load(GOLDEN))
    scan = synth.make_cloud(z["scan_pos0"], z["scan_nor0"])
    objs = []
    for i in range(int(z["n_objects"][0])):
        objs.append(synth.make_cloud(z[f"obj{i}_pos0"], z[f"obj{i}_nor0"]))
    return z, scan, objs


def check_rows(gi, gd, gn, oi, od, on, k):
    """NN parity: counts and squared distances bit-identical, indices identical wherever the distance is not tied"""
    assert (gn == on).all(), f"neighbour counts differ for {(gn != on).sum()} queries"
    m = np.arange(k)[None, :] < on[:, None]
    assert (gd[m] == od[m]).all(), "squared distances are not bit-identical"
    diff = m & (gi != oi)
    for r in np.unique(np.nonzero(diff)[0]):
        d = od[r, : on[r]]
        for j in np.nonzero(diff[r])[0]:
            assert (d == d[j]).sum() > 1 or j == k - 1, f"index mismatch without a distance tie (row {r}, col {j})"
