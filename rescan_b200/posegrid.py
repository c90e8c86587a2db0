"""Host-side construction of candidate pose grids with the reference's own float arithmetic.

The dense search of reference apps/pose_proposal/pose_proposal.cpp:213-222 builds every candidate as
``msh_rotate(identity, y_angle, +Y)`` with the translation column overwritten, using float loop counters that
accumulate rounding (``for(float y = 0; y < 2pi; y += inc)``).  Poses handed to the GPU must carry exactly
those bits, so this module restates ``msh_rotate`` (lib/msh/msh_vec_math.h:2089-2136) in float32 scalar
arithmetic with the C library's ``cosf`` / ``sinf`` — the same functions the reference's host code calls.
Pure host code: no GPU, no oracle.
"""
from __future__ import annotations

import ctypes
import ctypes.util
import numpy as np

_libm = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")
_libm.cosf.restype = ctypes.c_float
_libm.cosf.argtypes = [ctypes.c_float]
_libm.sinf.restype = ctypes.c_float
_libm.sinf.argtypes = [ctypes.c_float]

F = np.float32
MSH_TWO_PI = 6.2831853072  # msh_std.h:619 (a double literal, exactly this many digits)


def rotate(m16, angle, axis):
    """msh_rotate(m, angle, axis): column-major float32[16] in, new float32[16] out"""
    m = np.asarray(m16, F).copy()
    c, s = F(_libm.cosf(float(F(angle)))), F(_libm.sinf(float(F(angle))))
    t = F(F(1.0) - c)
    ax, ay, az = (F(v) for v in axis)
    inv = F(F(1.0) / F(np.sqrt(F(F(F(ax * ax) + F(ay * ay)) + F(az * az)))))
    ax, ay, az = F(ax * inv), F(ay * inv), F(az * inv)
    R = np.zeros(16, F)
    R[0] = F(c + F(F(ax * ax) * t))
    R[5] = F(c + F(F(ay * ay) * t))
    R[10] = F(c + F(F(az * az) * t))
    a, b = F(F(ax * ay) * t), F(az * s)
    R[1], R[4] = F(a + b), F(a - b)
    a, b = F(F(ax * az) * t), F(ay * s)
    R[2], R[8] = F(a - b), F(a + b)
    a, b = F(F(ay * az) * t), F(ax * s)
    R[6], R[9] = F(a + b), F(a - b)
    o = m.copy()
    for j in range(3):
        for r in range(4):
            o[4 * j + r] = F(F(m[r] * R[4 * j]) + F(F(m[4 + r] * R[4 * j + 1]) + F(m[8 + r] * R[4 * j + 2])))
    return o


def identity():
    m = np.zeros(16, F)
    m[0] = m[5] = m[10] = m[15] = F(1.0)
    return m


def make_pose(y_angle, tx, ty, tz):
    """the candidate of pose_proposal.cpp:221-222"""
    m = rotate(identity(), y_angle, (0.0, 1.0, 0.0))
    m[12], m[13], m[14], m[15] = F(tx), F(ty), F(tz), F(1.0)
    return m


def rotation_angles(n_rot=None, angle_delta=None):
    """the float-accumulated yaw angles of `for(float y = 0.0f; y < MSH_TWO_PI; y += inc)` (pose_proposal.cpp:219);
    inc = MSH_TWO_PI / n_rot as mgs_init_opts does for n_rot = 10 (pose_proposal.cpp:24-34), or given directly"""
    inc = F(angle_delta) if angle_delta is not None else F(MSH_TWO_PI / float(n_rot))
    out, a = [], F(0.0)
    while float(a) < MSH_TWO_PI:  # float promoted to double for the comparison
        out.append(a)
        a = F(a + inc)
    return np.asarray(out, F)


def rotation_xforms(n_rot=None, angle_delta=None):
    """[R, 16] column-major rotations about +Y with zero translation.  With `angle_delta` the list is exactly the
    reference loop's (which may run one step past a full turn through accumulated rounding); with `n_rot` it is
    the first n_rot angles of that loop, so that a workload quoted as "36 rotations" has 36."""
    ang = rotation_angles(n_rot, angle_delta)
    if angle_delta is None:
        ang = ang[:n_rot]
    return np.stack([make_pose(a, 0.0, 0.0, 0.0) for a in ang]).astype(F)


def reference_translation_grid(bbox_min, bbox_max, spacing=0.10):
    """the reference's xz lattice over the scan bbox +- one step at y = 0, ox-major / oz-minor, with its
    float-accumulating loop counters (pose_proposal.cpp:203-222) -> float32 [T, 3]"""
    sp = F(spacing)
    lx = F(F(bbox_max[0]) - F(bbox_min[0]))
    lz = F(F(bbox_max[2]) - F(bbox_min[2]))

    def axis(length):
        out, o = [], F(-sp)
        while o < F(length + sp):
            out.append(o)
            o = F(o + sp)
        return out

    oxs, ozs = axis(lx), axis(lz)
    t = np.zeros((len(oxs) * len(ozs), 3), F)
    k = 0
    for ox in oxs:
        for oz in ozs:
            t[k, 0] = F(F(bbox_min[0]) + ox)
            t[k, 2] = F(F(bbox_min[2]) + oz)
            k += 1
    return t


def pose_grid(rotations, translations):
    """explicit [T, R, 16] xforms of the dense grid (what the GPU builds on the fly)"""
    r = np.asarray(rotations, F).reshape(-1, 16)
    t = np.asarray(translations, F).reshape(-1, 3)
    out = np.broadcast_to(r[None, :, :], (len(t), len(r), 16)).copy()
    out[:, :, 12:15] = t[:, None, :]
    out[:, :, 15] = 1.0
    return out


def cloud_centroid(pos0):
    """rs_pointcloud_centroid( pc, 0 ): fp64 accumulation in point order, divided, narrowed to float
    (reference lib/rs/rs_pointcloud.h:1319-1339).  np.cumsum accumulates sequentially, np.sum would not."""
    p = np.ascontiguousarray(pos0, np.float32).reshape(-1, 3)
    if len(p) == 0:
        return np.zeros(3, np.float32)
    c = np.array([np.cumsum(p[:, a].astype(np.float64))[-1] for a in range(3)])
    return (c / float(len(p))).astype(np.float32)


def spatial_order(translations, cell=0.1):
    """permutation that walks the translations along a Z-order curve over (x, z) in steps of `cell`: consecutive poses of a
    dense launch then search neighbouring scan cells (cache locality); used with rsgpu_propose_opts_t.translation_ids, which
    keeps ids, order and ties in the caller's numbering"""
    t = np.asarray(translations, np.float32).reshape(-1, 3)
    if len(t) == 0:
        return np.zeros(0, np.int64)
    q = np.floor((t[:, [0, 2]].astype(np.float64) - t[:, [0, 2]].astype(np.float64).min(0)) / cell)
    q = np.clip(q, 0, 65535).astype(np.uint64)

    def spread(v):
        v = (v | (v << np.uint64(8))) & np.uint64(0x00FF00FF)
        v = (v | (v << np.uint64(4))) & np.uint64(0x0F0F0F0F)
        v = (v | (v << np.uint64(2))) & np.uint64(0x33333333)
        v = (v | (v << np.uint64(1))) & np.uint64(0x55555555)
        return v
    return np.argsort(spread(q[:, 0]) | (spread(q[:, 1]) << np.uint64(1)), kind="stable").astype(np.int64)
