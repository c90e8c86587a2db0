"""TEST INFRASTRUCTURE ONLY — ctypes view of oracle/_build/librescan_oracle.so (our plain-C CPU
restatement, oracle/rescan_oracle.c).  Importable only from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs — never from rescan_b200/.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "librescan_oracle.so")
_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
_u64p = np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
_i8p = np.ctypeslib.ndpointer(np.int8, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")

_lib = None


def build():
    subprocess.check_call(["make", "-s", "-f", os.path.join(HERE, "Makefile"), "oracle"])


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(os.path.join(HERE, "rescan_oracle.c")):
        build()
    L = C.CDLL(LIB)
    vp, i32, i64, f32, sz = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_size_t
    sig = {
        "orc_grid_build": (vp, [_f32p, i32, f32]),
        "orc_grid_free": (None, [vp]),
        "orc_grid_info": (None, [vp, _i64p, _f64p, _f32p, _i64p]),
        "orc_grid_data": (None, [vp, _f32p, _i32p]),
        "orc_radius_search": (sz, [vp, _f32p, sz, f32, sz, C.c_int, _f32p, _i32p, _u64p]),
        "orc_knn_search": (sz, [vp, _f32p, sz, sz, C.c_int, _f32p, _i32p, _u64p]),
        "orc_xf_apply": (None, [_f32p, _f32p, C.c_int, _f32p]),
        "orc_xf_mul": (None, [_f32p, _f32p, _f32p]),
        "orc_xf_inverse": (None, [_f32p, _f32p]),
        "orc_make_pose": (None, [f32, f32, f32, f32, _f32p]),
        "orc_score_pose": (f32, [_f32p, _f32p, i32, vp, _f32p, _f32p, i32, f32]),
        "orc_score_batch": (C.c_double, [_f32p, _f32p, i32, vp, _f32p, _f32p, i64, i32, f32, _f32p, C.c_int]),
        "orc_score_threshold": (f32, [C.c_int]),
        "orc_select_proposals": (i32, [_f32p, i32, i32, f32, _i32p, _i32p, _f32p]),
        "orc_icp_find_corrs": (i32, [_f32p, _f32p, i32, vp, _f32p, _f32p, _f32p, _f32p, f32, f32,
                                     _f32p, _f32p, _f32p, _f32p, _f32p]),
        "orc_icp_pt2pl": (f32, [_f32p, _f32p, _f32p, _f32p, i32, _f32p]),
        "orc_icp_align": (f32, [_f32p, _f32p, i32, _f32p, _f32p, i32, _f32p, _f32p, f32, f32, C.POINTER(i32)]),
        "orc_assign_labels": (None, [_f32p, _f32p, i32, _f32p, C.POINTER(vp), C.POINTER(vp), i32, i32, f32, _i8p, _f32p]),
        "orc_unary_costs": (None, [_i32p, _u8p, i32, i32, _i32p]),
        "orc_posed_bbox": (None, [_f32p, i32, _f32p, _f32p]),
        "orc_overlap_factor": (f32, [_f32p, i32, _f32p, i32, _f32p, _f32p, f32, C.c_int, C.c_int]),
        "orc_nms": (None, [_f32p, i32, _f32p, i32, _f32p, _f32p, i32, f32, _u8p]),
        "orc_poisson_level": (i32, [_f32p, i32, f32, i32, _i32p]),
        "orc_cov_grid": (i32, [_f32p, _f32p, f32, _i32p, _f32p]),
        "orc_cov_rasterize": (None, [_f32p, i32, vp, _f32p, _i32p, f32, _u8p]),
        "orc_cov_score": (f32, [_u8p, _u8p, i32]),
        "orc_neighborhood": (None, [vp, _f32p, _f32p, i32, i32, f32, f32, f32, _i32p, _f32p]),
        "orc_plane_inlier_counts": (None, [_f32p, _u8p, i32, _f32p, i32, f32, _i32p]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def _f32(a):
    return np.ascontiguousarray(a, np.float32)


IDENTITY = np.eye(4, dtype=np.float32).reshape(16)


class OrcGrid:
    def __init__(self, pts, radius):
        self.L = load()
        self.pts = _f32(pts)
        self.h = self.L.orc_grid_build(self.pts.reshape(-1), len(self.pts), radius)

    def info(self):
        dims = np.zeros(3, np.int64)
        cell = np.zeros(2, np.float64)
        mm = np.zeros(6, np.float32)
        cnt = np.zeros(3, np.int64)
        self.L.orc_grid_info(self.h, dims, cell, mm, cnt)
        return dict(dims=dims, cell_size=cell[0], inv_cell_size=cell[1], min_pt=mm[:3], max_pt=mm[3:],
                    n_pts=int(cnt[0]), n_bins=int(cnt[1]), max_n_pts_in_bin=int(cnt[2]))

    def data(self):
        n = len(self.pts)
        xyz = np.zeros((n, 3), np.float32)
        idx = np.zeros(n, np.int32)
        self.L.orc_grid_data(self.h, xyz.reshape(-1), idx)
        return xyz, idx

    def radius_search(self, q, radius, k, sort=1):
        q = _f32(q)
        nq = len(q)
        d2 = np.full((nq, k), np.nan, np.float32)
        idx = np.full((nq, k), -1, np.int32)
        nn = np.zeros(nq, np.uint64)
        tot = self.L.orc_radius_search(self.h, q.reshape(-1), nq, radius, k, sort, d2.reshape(-1), idx.reshape(-1), nn)
        return idx, d2, nn.astype(np.int64), int(tot)

    def knn_search(self, q, k, sort=1):
        q = _f32(q)
        nq = len(q)
        d2 = np.full((nq, k), np.nan, np.float32)
        idx = np.full((nq, k), -1, np.int32)
        nn = np.zeros(nq, np.uint64)
        tot = self.L.orc_knn_search(self.h, q.reshape(-1), nq, k, sort, d2.reshape(-1), idx.reshape(-1), nn)
        return idx, d2, nn.astype(np.int64), int(tot)

    def close(self):
        if self.h:
            self.L.orc_grid_free(self.h)
        self.h = None


def make_pose(angle, tx, ty, tz):
    out = np.zeros(16, np.float32)
    load().orc_make_pose(angle, tx, ty, tz, out)
    return out


def make_pose_grid(angles, translations):
    """[T, R, 16] column-major xforms, rotation about +Y then translation column (pose_proposal.cpp:221-222)"""
    angles = np.asarray(angles, np.float32)
    tr = _f32(translations)
    out = np.zeros((len(tr), len(angles), 16), np.float32)
    for r, a in enumerate(angles):
        base = make_pose(a, 0, 0, 0)
        out[:, r, :] = base
        out[:, r, 12:15] = tr
    return out


def score_poses(obj_pos, obj_nor, scan_grid: OrcGrid, scan_nor, xforms, k=64, radius=0.10, n_threads=1):
    L = load()
    x = _f32(xforms).reshape(-1, 16)
    out = np.zeros(len(x), np.float32)
    op, on, sn = _f32(obj_pos), _f32(obj_nor), _f32(scan_nor)
    dt = L.orc_score_batch(op.reshape(-1), on.reshape(-1), len(op), scan_grid.h, sn.reshape(-1), x.reshape(-1), len(x),
                           k, radius, out, n_threads)
    return out, float(dt)


def select_proposals(scores_TR, threshold):
    s = _f32(scores_TR)
    T, R = s.shape
    ot, orr, os_ = np.zeros(T, np.int32), np.zeros(T, np.int32), np.zeros(T, np.float32)
    n = load().orc_select_proposals(s.reshape(-1), T, R, threshold, ot, orr, os_)
    return ot[:n], orr[:n], os_[:n]


def icp_find_corrs(p1, n1, grid2: OrcGrid, p2, n2, T1, max_dist, max_angle, T2=None):
    L = load()
    p1, n1, p2, n2 = _f32(p1), _f32(n1), _f32(p2), _f32(n2)
    n = len(p1)
    outs = [np.zeros((max(n, 1), 3), np.float32) for _ in range(4)]
    w = np.zeros(max(n, 1), np.float32)
    nc = L.orc_icp_find_corrs(p1.reshape(-1), n1.reshape(-1), n, grid2.h, p2.reshape(-1), n2.reshape(-1), _f32(T1),
                              _f32(T2) if T2 is not None else IDENTITY, max_dist, max_angle,
                              *[o.reshape(-1) for o in outs], w)
    return [o[:nc] for o in outs] + [w[:nc]]


def icp_pt2pl(cp1, cp2, cn2, w, T1):
    T = _f32(T1).copy()
    err = load().orc_icp_pt2pl(_f32(cp1).reshape(-1), _f32(cp2).reshape(-1), _f32(cn2).reshape(-1), _f32(w), len(w), T)
    return T, float(err)


def icp_align(p1, n1, p2, n2, T1, max_dist, max_angle, T2=None):
    L = load()
    T = _f32(T1).copy()
    p1, n1, p2, n2 = _f32(p1), _f32(n1), _f32(p2), _f32(n2)
    it = C.c_int32(0)
    err = L.orc_icp_align(p1.reshape(-1), n1.reshape(-1), len(p1), p2.reshape(-1), n2.reshape(-1), len(p2), T,
                          _f32(T2) if T2 is not None else IDENTITY, max_dist, max_angle, C.byref(it))
    return T, float(err), int(it.value)


def assign_labels(scan_pos, scan_nor, poses, obj_grids, obj_normals, first, last, radius, labels, min_d):
    L = load()
    sp, sn = _f32(scan_pos), _f32(scan_nor)
    A = len(obj_grids)
    gh = (C.c_void_p * A)(*[g.h for g in obj_grids])
    keep = [_f32(n) for n in obj_normals]
    nh = (C.c_void_p * A)(*[k.ctypes.data for k in keep])
    L.orc_assign_labels(sp.reshape(-1), sn.reshape(-1), len(sp), _f32(poses).reshape(-1), gh, nh, first, last, radius,
                        labels, min_d)


def unary_costs(labels, label_is_static, n_labels):
    lab = np.ascontiguousarray(labels, np.int32)
    st = np.ascontiguousarray(label_is_static, np.uint8)
    out = np.zeros((len(lab), n_labels), np.int32)
    load().orc_unary_costs(lab, st, len(lab), n_labels, out.reshape(-1))
    return out


def neighborhood(grid: OrcGrid, pos, nor, max_nn=8, radius_sq=np.float32(0.05) * np.float32(0.05), dist_exp=15.0, angle_exp=16.0):
    p, n = _f32(pos), _f32(nor)
    nbr = np.zeros((len(p), max_nn), np.int32)
    w = np.zeros((len(p), max_nn), np.float32)
    load().orc_neighborhood(grid.h, p.reshape(-1), n.reshape(-1), len(p), max_nn, radius_sq, dist_exp, angle_exp,
                            nbr.reshape(-1), w.reshape(-1))
    return nbr, w


def overlap_factor(pos3, pos1, pose_a, pose_b, voxel=0.1, inside=1, normalize_by_smaller=0):
    """isect_get_overlap_factor restated: level-3 points give the boxes, level-1 points the occupancy"""
    p3, p1 = _f32(pos3).reshape(-1, 3), _f32(pos1).reshape(-1, 3)
    return float(load().orc_overlap_factor(p3.reshape(-1), len(p3), p1.reshape(-1), len(p1), _f32(pose_a).reshape(16),
                                           _f32(pose_b).reshape(16), voxel, inside, normalize_by_smaller))


def centroid(pos0):
    """rs_pointcloud_centroid: double accumulation in point order, narrowed to float (rs_pointcloud.h:1319-1339)"""
    p = _f32(pos0).reshape(-1, 3)
    c = np.zeros(3, np.float64)
    for a in range(3):
        c[a] = np.cumsum(p[:, a].astype(np.float64))[-1] if len(p) else 0.0
    return (c / float(len(p))).astype(np.float32)


def nms(pos3, pos1, centroid_xyz, proposals, dist_threshold=0.2):
    """mgs_non_maxima_suppresion for one object -> keep flags [n]"""
    p3, p1 = _f32(pos3).reshape(-1, 3), _f32(pos1).reshape(-1, 3)
    pr = _f32(proposals).reshape(-1, 17)
    keep = np.zeros(len(pr), np.uint8)
    load().orc_nms(p3.reshape(-1), len(p3), p1.reshape(-1), len(p1), _f32(centroid_xyz).reshape(3), pr.reshape(-1), len(pr),
                   dist_threshold, keep)
    return keep.astype(bool)


LEVEL_VOXEL = (0.005, 0.01, 0.02, 0.04, 0.08)  # rs_pointcloud_init (rs_pointcloud.h:145)


def poisson_level(pos0, level):
    """rs_pointcloud__compute_level_poisson (rs_pointcloud.h:984-1037) -> ascending level-0 indices of the level's points"""
    p = _f32(pos0).reshape(-1, 3)
    out = np.zeros(len(p), np.int32)
    n = load().orc_poisson_level(p.reshape(-1), len(p), np.float32(LEVEL_VOXEL[level]), level, out)
    return out[:n].copy()


def cov_grid(bbox_min, bbox_max, voxel=0.05):
    """isect_grid3d_init over a bbox -> (res int32[3], origin float32[3], n_cells)"""
    res, origin = np.zeros(3, np.int32), np.zeros(3, np.float32)
    n = load().orc_cov_grid(_f32(bbox_min).reshape(3), _f32(bbox_max).reshape(3), np.float32(voxel), res, origin)
    return res, origin, int(n)


def cov_rasterize(pts, pose, res, origin, voxel=0.05, grid=None):
    """cells lit by the points (under `pose`, column-major 16 floats, or None) -> uint8 grid [n_cells] (OR-ed into `grid`)"""
    p = _f32(pts).reshape(-1, 3)
    res = np.ascontiguousarray(res, np.int32)
    if grid is None:
        grid = np.zeros(int(res[0]) * int(res[1]) * int(res[2]), np.uint8)
    ps = _f32(pose).reshape(16) if pose is not None else None
    load().orc_cov_rasterize(p.reshape(-1), len(p), ps.ctypes.data if ps is not None else None, _f32(origin).reshape(3), res, np.float32(voxel), grid)
    return grid


def cov_score(scn_grid, arr_grid):
    return float(load().orc_cov_score(np.ascontiguousarray(scn_grid, np.uint8), np.ascontiguousarray(arr_grid, np.uint8), len(scn_grid)))


def plane_inlier_counts(pts, active, planes, dist_threshold):
    """evaluate_plane_model for a list of candidate planes [P, 6] = {center, normal} -> int32 counts [P]"""
    p = _f32(pts).reshape(-1, 3)
    a = np.ascontiguousarray(active, np.uint8)
    pl = _f32(planes).reshape(-1, 6)
    out = np.zeros(len(pl), np.int32)
    load().orc_plane_inlier_counts(p.reshape(-1), a, len(p), pl.reshape(-1), len(pl), np.float32(dist_threshold), out)
    return out
