"""TEST INFRASTRUCTURE ONLY — ctypes view of oracle/_ref/librescan_ref*.so (the UNMODIFIED reference
compiled in place by oracle/Makefile; see oracle/ref_harness.cpp for what each entry point wraps).

Importable only from tests/, __graft_entry__.smoke() and bench.py's CPU-baseline / --impl reference legs.
"""
from __future__ import annotations

import ctypes as C
import os
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
_u64p = np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")


def lib_path(openmp=False):
    return os.path.join(HERE, "_ref", "librescan_ref_omp.so" if openmp else "librescan_ref.so")


def available(openmp=False):
    return os.path.exists(lib_path(openmp))


_libs = {}


def load(openmp=False):
    if openmp in _libs:
        return _libs[openmp]
    L = C.CDLL(lib_path(openmp))
    vp, i32, i64, f32, sz = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_size_t
    sig = {
        "ref_sizeof_hash_grid": (C.c_int, []),
        "ref_openmp_threads": (C.c_int, []),
        "ref_grid_build": (vp, [_f32p, i32, f32]),
        "ref_grid_free": (None, [vp]),
        "ref_grid_info": (None, [vp, _i64p, _f64p, _f32p, _i64p]),
        "ref_grid_data": (None, [vp, _f32p, _i32p]),
        "ref_grid_set_threads": (None, [vp, C.c_int]),
        "ref_grid_cell": (None, [vp, i64, _i64p]),
        "ref_radius_search": (sz, [vp, _f32p, sz, f32, sz, C.c_int, _f32p, _i32p, _u64p]),
        "ref_knn_search": (sz, [vp, _f32p, sz, sz, C.c_int, _f32p, _i32p, _u64p]),
        "ref_cloud_from_level0": (vp, [_f32p, _f32p, vp, vp, i32]),
        "ref_cloud_create": (vp, []),
        "ref_cloud_set_level": (None, [vp, C.c_int, _f32p, _f32p, i32, C.c_int]),
        "ref_cloud_n": (i32, [vp, C.c_int]),
        "ref_cloud_get_level": (None, [vp, C.c_int, _f32p, _f32p]),
        "ref_cloud_get_ids": (None, [vp, C.c_int, _i32p, _i32p]),
        "ref_cloud_bbox": (None, [vp, _f32p]),
        "ref_cloud_grid": (vp, [vp, C.c_int]),
        "ref_cloud_free": (None, [vp]),
        "ref_score": (f32, [vp, vp, C.c_int, C.c_int, _f32p, C.c_int]),
        "ref_score_batch": (C.c_double, [vp, vp, C.c_int, C.c_int, _f32p, i64, C.c_int, _f32p, C.c_int]),
        "ref_db_create": (vp, []),
        "ref_db_n_classes": (C.c_int, []),
        "ref_db_class_name": (C.c_char_p, [C.c_int]),
        "ref_db_is_class_static": (C.c_int, [vp, C.c_int]),
        "ref_db_add_object": (C.c_int, [vp, vp, C.c_int, C.c_int]),
        "ref_propose_poses": (C.c_int, [vp, vp, f32, f32, _i32p, C.POINTER(C.POINTER(C.c_float))]),
        "ref_free": (None, [vp]),
        "ref_make_pose": (None, [f32, f32, f32, f32, _f32p]),
        "ref_icp_align": (f32, [_f32p, _f32p, i32, _f32p, _f32p, i32, _f32p, _f32p, f32, f32]),
        "ref_icp_find_corrs": (i32, [_f32p, _f32p, i32, _f32p, _f32p, i32, _f32p, _f32p, f32, f32, f32,
                                     _f32p, _f32p, _f32p, _f32p, _f32p]),
        "ref_icp_pt2pl": (f32, [_f32p, _f32p, _f32p, _f32p, i32, _f32p]),
        "ref_arrangement_to_labels": (None, [vp, vp, _i32p, _i32p, _f32p, i32, f32, C.c_int]),
        "ref_smooth_labels_capture": (C.c_int, [vp, vp]),
        "ref_capture_n_edges": (i64, []),
        "ref_capture_get": (None, [vp, vp, vp, vp, vp, vp]),
        "ref_overlap_factor": (f32, [vp, _f32p, _f32p, f32, C.c_int, C.c_int]),
        "ref_cloud_centroid": (None, [vp, _f32p]),
        "ref_cov_grid": (i32, [_f32p, _f32p, f32, _i32p, _f32p]),
        "ref_cov_rasterize": (None, [_f32p, _f32p, f32, _f32p, i32, vp, _u8p]),
        "ref_nms": (i32, [vp, C.c_int, _f32p, i32, f32, _f32p]),
        "ref_compute_neighborhood": (i64, [vp, C.c_int, C.c_int, f32, f32, f32, C.POINTER(C.POINTER(C.c_int32)),
                                          C.POINTER(C.POINTER(C.c_int32)), C.POINTER(C.POINTER(C.c_float))]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _libs[openmp] = L
    return L


def _f32(a):
    return np.ascontiguousarray(a, np.float32)


def mat_to_colmajor(m4):
    """math-convention 4x4 -> the 16 floats of msh_mat4_t (column-major, translation at 12..14)"""
    return _f32(np.asarray(m4, np.float32).T.reshape(16))


def colmajor_to_mat(f16):
    return np.asarray(f16, np.float32).reshape(4, 4).T.copy()


class RefGrid:
    def __init__(self, pts, radius, openmp=False, handle=None):
        self.L = load(openmp)
        self.pts = _f32(pts)
        self.owned = handle is None
        self.h = handle if handle is not None else self.L.ref_grid_build(self.pts.reshape(-1), len(self.pts), radius)

    def info(self):
        dims = np.zeros(3, np.int64)
        cell = np.zeros(2, np.float64)
        mm = np.zeros(6, np.float32)
        cnt = np.zeros(3, np.int64)
        self.L.ref_grid_info(self.h, dims, cell, mm, cnt)
        return dict(dims=dims, cell_size=cell[0], inv_cell_size=cell[1], min_pt=mm[:3], max_pt=mm[3:],
                    n_pts=int(cnt[0]), n_bins=int(cnt[1]), max_n_pts_in_bin=int(cnt[2]))

    def data(self):
        n = self.info()["n_pts"]
        xyz = np.zeros((n, 3), np.float32)
        idx = np.zeros(n, np.int32)
        self.L.ref_grid_data(self.h, xyz.reshape(-1), idx)
        return xyz, idx

    def cell(self, c):
        ol = np.zeros(2, np.int64)
        self.L.ref_grid_cell(self.h, int(c), ol)
        return int(ol[0]), int(ol[1])

    def radius_search(self, q, radius, k, sort=1):
        q = _f32(q)
        nq = len(q)
        d2 = np.full((nq, k), np.nan, np.float32)
        idx = np.full((nq, k), -1, np.int32)
        nn = np.zeros(nq, np.uint64)
        tot = self.L.ref_radius_search(self.h, q.reshape(-1), nq, radius, k, sort, d2.reshape(-1), idx.reshape(-1), nn)
        return idx, d2, nn.astype(np.int64), int(tot)

    def knn_search(self, q, k, sort=1):
        q = _f32(q)
        nq = len(q)
        d2 = np.full((nq, k), np.nan, np.float32)
        idx = np.full((nq, k), -1, np.int32)
        nn = np.zeros(nq, np.uint64)
        tot = self.L.ref_knn_search(self.h, q.reshape(-1), nq, k, sort, d2.reshape(-1), idx.reshape(-1), nn)
        return idx, d2, nn.astype(np.int64), int(tot)

    def close(self):
        if self.owned and self.h:
            self.L.ref_grid_free(self.h)
        self.h = None


class RefCloud:
    """rs_pointcloud_t built either from level 0 (reference computes levels) or from explicit levels."""

    def __init__(self, openmp=False):
        self.L = load(openmp)
        self.openmp = openmp
        self.h = None

    @classmethod
    def from_level0(cls, pos, nor, class_ids=None, instance_ids=None, openmp=False):
        self = cls(openmp)
        pos, nor = _f32(pos), _f32(nor)
        ci = np.ascontiguousarray(class_ids, np.int32) if class_ids is not None else None
        ii = np.ascontiguousarray(instance_ids, np.int32) if instance_ids is not None else None
        self.h = self.L.ref_cloud_from_level0(pos.reshape(-1), nor.reshape(-1),
                                              ci.ctypes.data if ci is not None else None,
                                              ii.ctypes.data if ii is not None else None, len(pos))
        return self

    @classmethod
    def from_levels(cls, levels, openmp=False, bbox_level=0):
        """levels: dict/list lvl -> (pos, nor); bbox from `bbox_level` like rs_pointcloud__compute_bbox"""
        self = cls(openmp)
        self.h = self.L.ref_cloud_create()
        items = levels.items() if isinstance(levels, dict) else enumerate(levels)
        for lvl, (p, n) in items:
            p, n = _f32(p), _f32(n)
            self.L.ref_cloud_set_level(self.h, lvl, p.reshape(-1), n.reshape(-1), len(p), int(lvl == bbox_level))
        return self

    def n(self, lvl):
        return self.L.ref_cloud_n(self.h, lvl)

    def level(self, lvl):
        n = self.n(lvl)
        p = np.zeros((n, 3), np.float32)
        q = np.zeros((n, 3), np.float32)
        self.L.ref_cloud_get_level(self.h, lvl, p.reshape(-1), q.reshape(-1))
        return p, q

    def ids(self, lvl):
        n = self.n(lvl)
        a = np.zeros(n, np.int32)
        b = np.zeros(n, np.int32)
        self.L.ref_cloud_get_ids(self.h, lvl, a, b)
        return a, b

    def bbox(self):
        b = np.zeros(6, np.float32)
        self.L.ref_cloud_bbox(self.h, b)
        return b[:3].copy(), b[3:].copy()

    def grid(self, lvl):
        return RefGrid(self.level(lvl)[0], 0.05, self.openmp, handle=self.L.ref_cloud_grid(self.h, lvl))

    def to_synth_cloud(self):
        from rescan_b200 import synth
        lv = [self.level(l) for l in range(5)]
        mn, mx = self.bbox()
        return synth.Cloud(lv, mn, mx)


def score(obj: RefCloud, scene: RefCloud, xform_colmajor, query_lvl, search_lvl=1, k=64):
    return float(obj.L.ref_score(obj.h, scene.h, search_lvl, query_lvl, _f32(xform_colmajor), k))


def score_batch(obj: RefCloud, scene: RefCloud, xforms_colmajor, query_lvl, search_lvl=1, k=64, n_threads=1):
    x = _f32(xforms_colmajor).reshape(-1, 16)
    out = np.zeros(len(x), np.float32)
    dt = obj.L.ref_score_batch(obj.h, scene.h, search_lvl, query_lvl, x.reshape(-1), len(x), k, out, n_threads)
    return out, float(dt)


def make_pose(L, angle, tx, ty, tz):
    out = np.zeros(16, np.float32)
    L.ref_make_pose(angle, tx, ty, tz, out)
    return out


class RefDB:
    def __init__(self, openmp=False):
        self.L = load(openmp)
        self.h = self.L.ref_db_create()
        self.objects = []

    def add_object(self, cloud: RefCloud, uidx, class_idx):
        self.objects.append(cloud)
        return self.L.ref_db_add_object(self.h, cloud.h, uidx, class_idx)

    def is_class_static(self, c):
        return bool(self.L.ref_db_is_class_static(self.h, c))

    def propose_poses(self, scan: RefCloud, spacing=0.0, angle_delta=0.0):
        counts = np.zeros(len(self.objects), np.int32)
        flat = C.POINTER(C.c_float)()
        n = self.L.ref_propose_poses(self.h, scan.h, spacing, angle_delta, counts, C.byref(flat))
        assert n == len(self.objects)
        tot = int(counts.sum())
        arr = np.ctypeslib.as_array(flat, shape=(max(tot, 1) * 17,)).copy()[: tot * 17].reshape(tot, 17)
        self.L.ref_free(flat)
        out, o = [], 0
        for c in counts:
            out.append(arr[o:o + c].copy())
            o += c
        return out

    def nms(self, object_idx, proposals, dist_threshold=0.2):
        """mgs_non_maxima_suppresion for one object -> surviving proposals [m, 17] in the reference's order"""
        pr = _f32(proposals).reshape(-1, 17)
        kept = np.zeros_like(pr)
        m = self.L.ref_nms(self.h, int(object_idx), pr.reshape(-1), len(pr), dist_threshold, kept.reshape(-1))
        return kept[:m].copy()

    def arrangement_to_labels(self, scan: RefCloud, object_idx, uidx, poses_colmajor, radius, prioritize_static=False):
        oi = np.ascontiguousarray(object_idx, np.int32)
        ui = np.ascontiguousarray(uidx, np.int32)
        ps = _f32(poses_colmajor).reshape(-1)
        self.L.ref_arrangement_to_labels(self.h, scan.h, oi, ui, ps, len(oi), radius, int(prioritize_static))
        return scan.ids(1)

    def smooth_labels_capture(self, scan: RefCloud):
        n_labels = self.L.ref_smooth_labels_capture(self.h, scan.h)
        v = scan.n(1)
        ne = self.L.ref_capture_n_edges()
        dc = np.zeros((v, n_labels), np.int32)
        sc = np.zeros((n_labels, n_labels), np.int32)
        il = np.zeros(v, np.int32)
        ea, eb, ew = (np.zeros(max(ne, 1), np.int32) for _ in range(3))
        self.L.ref_capture_get(dc.ctypes.data, sc.ctypes.data, il.ctypes.data, ea.ctypes.data, eb.ctypes.data, ew.ctypes.data)
        return dict(n_labels=n_labels, data_cost=dc, smooth_cost=sc, init_labels=il, edges=(ea[:ne], eb[:ne], ew[:ne]))


def compute_neighborhood(cloud: RefCloud, lvl=1, max_nn=8, radius_sq=0.05 * 0.05, dist_exp=15.0, angle_exp=16.0):
    a, b, w = C.POINTER(C.c_int32)(), C.POINTER(C.c_int32)(), C.POINTER(C.c_float)()
    n = cloud.L.ref_compute_neighborhood(cloud.h, lvl, max_nn, np.float32(radius_sq), dist_exp, angle_exp,
                                         C.byref(a), C.byref(b), C.byref(w))
    ea = np.ctypeslib.as_array(a, shape=(max(n, 1),))[:n].copy()
    eb = np.ctypeslib.as_array(b, shape=(max(n, 1),))[:n].copy()
    ew = np.ctypeslib.as_array(w, shape=(max(n, 1),))[:n].copy()
    for p in (a, b, w):
        cloud.L.ref_free(p)
    return ea, eb, ew


def icp_align(p1, n1, p2, n2, T1_colmajor, max_dist, max_angle, T2_colmajor=None, openmp=False):
    L = load(openmp)
    T1 = _f32(T1_colmajor).copy()
    T2 = _f32(T2_colmajor) if T2_colmajor is not None else _f32(np.eye(4).reshape(16))
    p1, n1, p2, n2 = _f32(p1), _f32(n1), _f32(p2), _f32(n2)
    err = L.ref_icp_align(p1.reshape(-1), n1.reshape(-1), len(p1), p2.reshape(-1), n2.reshape(-1), len(p2),
                          T1, T2, max_dist, max_angle)
    return T1, float(err)


def icp_find_corrs(p1, n1, p2, n2, T1_colmajor, grid_radius, max_dist, max_angle, T2_colmajor=None):
    L = load(False)
    T1 = _f32(T1_colmajor)
    T2 = _f32(T2_colmajor) if T2_colmajor is not None else _f32(np.eye(4).reshape(16))
    p1, n1, p2, n2 = _f32(p1), _f32(n1), _f32(p2), _f32(n2)
    n = len(p1)
    outs = [np.zeros((n, 3), np.float32) for _ in range(4)]
    w = np.zeros(n, np.float32)
    nc = L.ref_icp_find_corrs(p1.reshape(-1), n1.reshape(-1), n, p2.reshape(-1), n2.reshape(-1), len(p2), T1, T2,
                              grid_radius, max_dist, max_angle, *[o.reshape(-1) for o in outs], w)
    return [o[:nc] for o in outs] + [w[:nc]]


def icp_pt2pl(cp1, cp2, cn2, w, T1_colmajor):
    L = load(False)
    T1 = _f32(T1_colmajor).copy()
    err = L.ref_icp_pt2pl(_f32(cp1).reshape(-1), _f32(cp2).reshape(-1), _f32(cn2).reshape(-1), _f32(w), len(w), T1)
    return T1, float(err)


def overlap_factor(cloud: RefCloud, pose_a, pose_b, voxel=0.1, inside=1, normalize_by_smaller=0):
    """isect_get_overlap_factor of one cloud under two poses (column-major float32[16])"""
    return float(cloud.L.ref_overlap_factor(cloud.h, _f32(pose_a).reshape(16), _f32(pose_b).reshape(16), voxel, inside, normalize_by_smaller))


def cloud_centroid(cloud: RefCloud):
    c = np.zeros(3, np.float32)
    cloud.L.ref_cloud_centroid(cloud.h, c)
    return c


def cov_grid(bbox_min, bbox_max, voxel=0.05):
    res, origin = np.zeros(3, np.int32), np.zeros(3, np.float32)
    n = load(False).ref_cov_grid(_f32(bbox_min).reshape(3), _f32(bbox_max).reshape(3), np.float32(voxel), res, origin)
    return res, origin, int(n)


def cov_rasterize(bbox_min, bbox_max, pts, pose, n_cells, voxel=0.05, grid=None):
    p = _f32(pts).reshape(-1, 3)
    if grid is None:
        grid = np.zeros(n_cells, np.uint8)
    ps = _f32(pose).reshape(16) if pose is not None else None
    load(False).ref_cov_rasterize(_f32(bbox_min).reshape(3), _f32(bbox_max).reshape(3), np.float32(voxel), p.reshape(-1), len(p),
                                  ps.ctypes.data if ps is not None else None, grid)
    return grid


def plane_inlier_counts(pts, weights, planes, dist_threshold):
    """the reference's own evaluate_plane_model (lib/rs/rs_pointcloud_filters.cpp:117-134, C++ linkage: called through its
    mangled name) once per candidate plane [P, 6] = {center, normal}; `weights` are the detector's doubles (> 0.01 = active)"""
    L = load()
    fn = getattr(L, "_Z20evaluate_plane_modelP16rspf_plane_modelPdP5vec3fmf")
    fn.restype = None
    fn.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_float]
    p = np.ascontiguousarray(pts, np.float32).reshape(-1, 3)
    w = np.ascontiguousarray(weights, np.float64)
    out = np.zeros(len(planes), np.int64)
    model = np.zeros(512, np.uint8)  # rspf_plane_model_t: plane {center, normal} at byte 0, size_t n_inliers at byte 24
    for k, pl in enumerate(np.ascontiguousarray(planes, np.float32).reshape(-1, 6)):
        model[:] = 0
        model[:24] = pl.view(np.uint8)
        fn(model.ctypes.data, w.ctypes.data, p.ctypes.data, len(p), np.float32(dist_threshold))
        out[k] = model[24:32].view(np.uint64)[0]
    return out
