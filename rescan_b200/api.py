"""Host-side mirror of the reference's hot-path interface, over the C ABI of ``include/rsgpu.h``.

Everything here is a thin ctypes view of ``rescan_b200/librsgpu.so`` (hand-written sm_100a CUDA, built by
``rescan_b200/csrc/Makefile``).  Names and argument meaning follow the reference entry points they replace:

  ================================  ==========================================================
  here                              reference
  ================================  ==========================================================
  ``HashGrid(pts, radius)``         ``msh_hash_grid_init_3d``        (lib/msh/msh_hash_grid.h:222)
  ``HashGrid.radius_search``        ``msh_hash_grid_radius_search``  (msh_hash_grid.h:227)
  ``HashGrid.knn_search``           ``msh_hash_grid_knn_search``     (msh_hash_grid.h:229)
  ``compute_object_alignment_scores`` ``mgs_compute_object_alignment_score`` (apps/pose_proposal/pose_proposal.h:46)
  ``propose_poses``                 ``mgs_propose_poses``            (pose_proposal.h:51)
  ``icp_align``                     ``icp_align``                    (lib/rs/icp.h:84)
  ``assign_labels`` / ``unary_costs`` / ``neighborhood``
                                    ``rspf__assign_temporary_labels`` / data_cost block / ``rspf_compute_neighborhood``
                                    (lib/rs/rs_pointcloud_filters.cpp:738, 926, 674)
  ================================  ==========================================================

There is no CPU fallback: if the library is missing or no CUDA device is present every call raises.
Nothing in this module imports or calls ``oracle/``.
"""
from __future__ import annotations

import os as _os

# tens of streams per process (one object chain per lane): more hardware channels than the default 8, so that unrelated
# streams do not queue behind one another.  Effective only before the process creates its CUDA context.
_os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

import ctypes as C
import os
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "librsgpu.so")

RSGPU_MAX_K = 512
POSE_FLOATS = 17


class RsgpuError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"rsgpu error {code}: {msg}")
        self.code = code


class GridInfo(C.Structure):
    _fields_ = [("width", C.c_int64), ("height", C.c_int64), ("depth", C.c_int64),
                ("cell_size", C.c_double), ("inv_cell_size", C.c_double),
                ("min_pt", C.c_float * 3), ("max_pt", C.c_float * 3),
                ("n_pts", C.c_int64), ("n_bins", C.c_int64), ("max_n_pts_in_bin", C.c_int64)]


class SearchDesc(C.Structure):
    _fields_ = [("query_pts", C.c_void_p), ("n_query_pts", C.c_size_t), ("distances_sq", C.c_void_p),
                ("indices", C.c_void_p), ("n_neighbors", C.c_void_p), ("radius", C.c_float),
                ("k", C.c_size_t), ("sort", C.c_int)]


class IcpJob(C.Structure):
    _fields_ = [("object", C.c_void_p), ("T1", C.c_void_p), ("n_batch", C.c_int32), ("errs", C.c_void_p), ("iters", C.c_void_p)]


class ProposeOpts(C.Structure):
    _fields_ = [("max_n_neigh", C.c_int32), ("radius", C.c_float), ("thresholds", C.c_float * 3), ("top_k", C.c_int32),
                ("translation_ids", C.c_void_p)]


_lib = None

# every symbol include/rsgpu.h declares: (restype, argtypes)
_vp, _i32, _i64, _f32, _sz, _int = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_size_t, C.c_int
SIGNATURES = {
    "rsgpu_device_count": (_int, []),
    "rsgpu_set_device": (_int, [_int]),
    "rsgpu_set_stream": (_int, [_vp]),
    "rsgpu_synchronize": (_int, []),
    "rsgpu_lane_count": (_int, []),
    "rsgpu_thread_attach": (_int, [_int]),
    "rsgpu_last_error": (C.c_char_p, []),
    "rsgpu_version": (C.c_char_p, []),
    "rsgpu_set_option": (_int, [C.c_char_p, C.c_char_p]),
    "rsgpu_profile_enable": (_int, [_int]),
    "rsgpu_profile_reset": (_int, []),
    "rsgpu_profile_get": (_int, [C.c_char_p, C.POINTER(C.c_double), C.POINTER(_i64)]),
    "rsgpu_launch_count": (_i64, []),
    "rsgpu_grid_create": (_int, [_vp, _i32, _f32, C.POINTER(_vp)]),
    "rsgpu_grid_create_dev": (_int, [_vp, _i32, _f32, C.POINTER(_vp)]),
    "rsgpu_grid_destroy": (None, [_vp]),
    "rsgpu_grid_set_normals": (_int, [_vp, _vp]),
    "rsgpu_grid_set_normals_dev": (_int, [_vp, _vp]),
    "rsgpu_grid_get_info": (_int, [_vp, C.POINTER(GridInfo)]),
    "rsgpu_grid_get_data": (_int, [_vp, _vp, _vp]),
    "rsgpu_grid_radius_search": (_int, [_vp, C.POINTER(SearchDesc), C.POINTER(_sz)]),
    "rsgpu_grid_knn_search": (_int, [_vp, C.POINTER(SearchDesc), C.POINTER(_sz)]),
    "rsgpu_grid_radius_search_dev": (_int, [_vp, C.POINTER(SearchDesc), C.POINTER(_sz)]),
    "rsgpu_grid_knn_search_dev": (_int, [_vp, C.POINTER(SearchDesc), C.POINTER(_sz)]),
    "rsgpu_grid_search_census_dev": (_int, [_vp, _vp, _sz, _f32, C.POINTER(_i64)]),
    "rsgpu_cloud_create": (_int, [_vp, _vp, _i32, C.POINTER(_vp)]),
    "rsgpu_cloud_destroy": (None, [_vp]),
    "rsgpu_cloud_size": (_i32, [_vp]),
    "rsgpu_score_poses": (_int, [_vp, _vp, _vp, _i64, _i32, _f32, _vp]),
    "rsgpu_score_poses_dev": (_int, [_vp, _vp, _vp, _i64, _i32, _f32, _vp]),
    "rsgpu_score_pose_grid": (_int, [_vp, _vp, _vp, _i32, _vp, _i64, _i32, _f32, _vp]),
    "rsgpu_score_pose_grid_dev": (_int, [_vp, _vp, _vp, _i32, _vp, _i64, _i32, _f32, _vp]),
    "rsgpu_score_pose_grid_count": (_int, [_vp, _vp, _vp, _i32, _vp, _i64, _i32, _f32, C.POINTER(_i64)]),
    "rsgpu_propose_default_opts": (None, [C.POINTER(ProposeOpts)]),
    "rsgpu_propose_poses": (_int, [_vp, _vp, _vp, _vp, _vp, _i32, _vp, _i64, C.POINTER(ProposeOpts), _vp, _vp, _i64,
                                   C.POINTER(_i64)]),
    "rsgpu_icp_align_batch": (_int, [_vp, _vp, _vp, _i32, _vp, _f32, _f32, _vp, _vp]),
    "rsgpu_icp_align_multi": (_int, [C.POINTER(IcpJob), _i32, _vp, _vp, _f32, _f32]),
    "rsgpu_icp_align_batch_ex": (_int, [_vp, _vp, _vp, _i32, _vp, _f32, _f32, _i32, _vp, _vp]),
    "rsgpu_assign_labels": (_int, [_vp, _vp, _i32, _vp, C.POINTER(_vp), _i32, _i32, _f32, _vp, _vp]),
    "rsgpu_unary_costs": (_int, [_vp, _vp, _i32, _i32, _vp]),
    "rsgpu_overlap_factors": (_int, [_vp, _vp, _vp, _vp, _i32, _f32, _i32, _i32, _vp]),
    "rsgpu_nms": (_int, [_vp, _vp, _vp, _vp, _i32, _f32, _vp]),
    "rsgpu_poisson_level": (_int, [_vp, _i32, _f32, _i32, _vp, C.POINTER(_i32), C.POINTER(_i32)]),
    "rsgpu_rasterize_points": (_int, [_vp, _i32, _vp, _vp, _vp, _f32, _vp]),
    "rsgpu_coverage_masks": (_int, [C.POINTER(_vp), _vp, _i32, _vp, _vp, _f32, _vp, _vp, _i32, C.POINTER(_i32)]),
    "rsgpu_neighborhood": (_int, [_vp, _vp, _vp, _i32, _i32, _f32, _f32, _f32, _vp, _vp]),
    "rsgpu_plane_inlier_counts": (_int, [_vp, _vp, _i32, _vp, _i32, _f32, _vp]),
    "rsgpu_peer_handle_bytes": (_int, []),
    "rsgpu_peer_init": (_int, [_i32, _i32, _i32, _i64, _vp]),
    "rsgpu_peer_open": (_int, [_vp]),
    "rsgpu_peer_allgather": (_int, [_i32, C.c_uint32, _vp, _i64, _vp, C.c_double]),
    "rsgpu_peer_put": (_int, [_i32, C.c_uint32, C.c_uint32, _vp, _i64]),
    "rsgpu_peer_get": (_int, [_i32, C.c_uint32, C.c_uint32, _vp, _i64, _vp, C.c_double]),
    "rsgpu_peer_close": (_int, []),
}


def lib():
    """Load librsgpu.so (once).  Raises if it has not been built: there is no fallback path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RsgpuError(-1, f"{LIB_PATH} is missing - build it with `make -f rescan_b200/csrc/Makefile` "
                             f"(or __graft_entry__.build()); there is no CPU fallback")
    L = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def _check(code):
    if code != 0:
        raise RsgpuError(code, lib().rsgpu_last_error().decode("utf-8", "replace"))


def _f32(a):
    return np.ascontiguousarray(a, np.float32)


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def device_count():
    return lib().rsgpu_device_count()


def set_device(i):
    _check(lib().rsgpu_set_device(int(i)))


def synchronize():
    _check(lib().rsgpu_synchronize())


def lane_count():
    return int(lib().rsgpu_lane_count())


def thread_attach(lane):
    """bind the calling host thread to a lane (its own stream); lane < 0 detaches"""
    _check(lib().rsgpu_thread_attach(int(lane)))


def launch_count():
    return int(lib().rsgpu_launch_count())


def set_option(name, value=None):
    """tuning / A-B knob of the library (see include/rsgpu.h); value None restores the default"""
    _check(lib().rsgpu_set_option(name.encode(), None if value is None else str(value).encode()))


def profile_enable(on=True):
    lib().rsgpu_profile_enable(int(bool(on)))


def profile_reset():
    lib().rsgpu_profile_reset()


def profile_get(name):
    ms, n = C.c_double(0), C.c_int64(0)
    lib().rsgpu_profile_get(name.encode(), C.byref(ms), C.byref(n))
    return float(ms.value), int(n.value)


class PointCloud:
    """One sampling level of an object model resident in HBM (positions + normals)."""

    def __init__(self, pos, nor):
        self.pos, self.nor = _f32(pos).reshape(-1, 3), _f32(nor).reshape(-1, 3)
        assert self.pos.shape == self.nor.shape
        h = C.c_void_p()
        _check(lib().rsgpu_cloud_create(_ptr(self.pos), _ptr(self.nor), len(self.pos), C.byref(h)))
        self.h = h

    def __len__(self):
        return len(self.pos)

    def close(self):
        if getattr(self, "h", None) and _lib is not None:  # at interpreter shutdown the module globals may already be gone
            _lib.rsgpu_cloud_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class HashGrid:
    """msh_hash_grid_t on the GPU.  ``normals`` (original point order) are needed by scoring / ICP / labels."""

    def __init__(self, pts=None, radius=0.05, normals=None, device_ptr=None, n_pts=None):
        h = C.c_void_p()
        if device_ptr is not None:
            _check(lib().rsgpu_grid_create_dev(C.c_void_p(device_ptr), int(n_pts), radius, C.byref(h)))
            self.pts = None
        else:
            self.pts = _f32(pts).reshape(-1, 3)
            _check(lib().rsgpu_grid_create(_ptr(self.pts), len(self.pts), radius, C.byref(h)))
        self.h = h
        if normals is not None:
            self.set_normals(normals)

    def set_normals(self, normals):
        n = _f32(normals).reshape(-1, 3)
        assert len(n) == self.info()["n_pts"]
        _check(lib().rsgpu_grid_set_normals(self.h, _ptr(n)))

    def info(self):
        gi = GridInfo()
        _check(lib().rsgpu_grid_get_info(self.h, C.byref(gi)))
        return dict(dims=np.array([gi.width, gi.height, gi.depth], np.int64), cell_size=gi.cell_size,
                    inv_cell_size=gi.inv_cell_size, min_pt=np.array(gi.min_pt[:], np.float32),
                    max_pt=np.array(gi.max_pt[:], np.float32), n_pts=int(gi.n_pts), n_bins=int(gi.n_bins),
                    max_n_pts_in_bin=int(gi.max_n_pts_in_bin))

    def data(self):
        n = self.info()["n_pts"]
        xyz = np.zeros((n, 3), np.float32)
        idx = np.zeros(n, np.int32)
        _check(lib().rsgpu_grid_get_data(self.h, _ptr(xyz), _ptr(idx)))
        return xyz, idx

    def _search(self, fn, q, radius, k, sort):
        q = _f32(q).reshape(-1, 3)
        nq = len(q)
        d2 = np.full((nq, k), np.nan, np.float32)
        idx = np.full((nq, k), -1, np.int32)
        nn = np.zeros(nq, np.uint64)
        sd = SearchDesc(q.ctypes.data, nq, d2.ctypes.data, idx.ctypes.data, nn.ctypes.data, radius, k, sort)
        tot = C.c_size_t(0)
        _check(fn(self.h, C.byref(sd), C.byref(tot)))
        return idx, d2, nn.astype(np.int64), int(tot.value)

    def radius_search(self, q, radius, k, sort=1):
        """-> indices [nq,k], distances_sq [nq,k], n_neighbors [nq], total (entries past a row's count are unspecified)"""
        return self._search(lib().rsgpu_grid_radius_search, q, radius, k, sort)

    def knn_search(self, q, k, sort=1):
        return self._search(lib().rsgpu_grid_knn_search, q, 0.0, k, sort)

    def radius_search_host(self, q, radius, k, d2, idx, nn, sort=1):
        """msh_hash_grid_radius_search with caller-allocated HOST buffers exactly as the reference takes them
        (msh_hash_grid.h:196-216: q float32 [nq,3], d2 float32 [nq,k], idx int32 [nq,k], nn uint64/int64 [nq]) -> total"""
        nq = len(q)
        assert q.dtype == np.float32 and d2.dtype == np.float32 and idx.dtype == np.int32 and nn.itemsize == 8
        assert d2.shape == (nq, k) and idx.shape == (nq, k) and len(nn) == nq
        sd = SearchDesc(q.ctypes.data, nq, d2.ctypes.data, idx.ctypes.data, nn.ctypes.data, radius, k, sort)
        tot = C.c_size_t(0)
        _check(lib().rsgpu_grid_radius_search(self.h, C.byref(sd), C.byref(tot)))
        return int(tot.value)

    def radius_search_dev(self, q_ptr, nq, radius, k, d2_ptr, idx_ptr, nn_ptr=None):
        """same search with every buffer already in HBM (raw device pointers; n_neighbors is uint64) -> total"""
        sd = SearchDesc(q_ptr, nq, d2_ptr, idx_ptr, nn_ptr, radius, k, 1)
        tot = C.c_size_t(0)
        _check(lib().rsgpu_grid_radius_search_dev(self.h, C.byref(sd), C.byref(tot)))
        return int(tot.value)

    def search_census_dev(self, q_ptr, nq, radius):
        """(non-empty cells, points) the reference's radius search reads for these device-resident queries"""
        c = (C.c_int64 * 2)()
        _check(lib().rsgpu_grid_search_census_dev(self.h, C.c_void_p(q_ptr), nq, radius, c))
        return int(c[0]), int(c[1])

    def close(self):
        if getattr(self, "h", None):
            lib().rsgpu_grid_destroy(self.h)
            self.h = None

    __del__ = close


def compute_object_alignment_scores(obj: PointCloud, scene: HashGrid, xforms, max_n_neigh=64, radius=0.10):
    """mgs_compute_object_alignment_score for a batch of column-major 4x4 poses -> float32 [P]"""
    x = _f32(xforms).reshape(-1, 16)
    out = np.zeros(len(x), np.float32)
    _check(lib().rsgpu_score_poses(obj.h, scene.h, _ptr(x), len(x), max_n_neigh, radius, _ptr(out)))
    return out


def score_pose_grid(obj: PointCloud, scene: HashGrid, rotations, translations, max_n_neigh=64, radius=0.10):
    """scores [T, R] of the dense pose grid rotations x translations (pose_proposal.cpp:213-236)"""
    r, t = _f32(rotations).reshape(-1, 16), _f32(translations).reshape(-1, 3)
    out = np.zeros((len(t), len(r)), np.float32)
    _check(lib().rsgpu_score_pose_grid(obj.h, scene.h, _ptr(r), len(r), _ptr(t), len(t), max_n_neigh, radius, _ptr(out)))
    return out


def score_pose_grid_count(obj: PointCloud, scene: HashGrid, rotations, translations, max_n_neigh=64, radius=0.10):
    """exact algorithmic-byte census of one dense scoring call -> dict(queries, B, C, T, bytes)"""
    r, t = _f32(rotations).reshape(-1, 16), _f32(translations).reshape(-1, 3)
    c = (C.c_int64 * 4)()
    _check(lib().rsgpu_score_pose_grid_count(obj.h, scene.h, _ptr(r), len(r), _ptr(t), len(t), max_n_neigh, radius, c))
    q, b, cc, tt = (int(v) for v in c)
    n_poses = len(r) * len(t)
    return dict(queries=q, B=b, C=cc, T=tt, poses=n_poses, bytes=8 * b + 16 * cc + 12 * tt + 68 * n_poses + 24 * len(obj))


def propose_poses(obj_lvl4: PointCloud, obj_lvl3: PointCloud, obj_lvl2: PointCloud, scene: HashGrid, rotations,
                  translations, max_n_neigh=64, radius=0.10, thresholds=(0.25, 0.35, 0.40), top_k=0, cap=None, translation_ids=None):
    """mgs_propose_poses for one object -> (proposals float32 [n,17] = 16 xform + score, pose_ids int64 [n]).
    translation_ids: the callers' numbering of the translations when they are passed in another order (see rsgpu.h)."""
    r, t = _f32(rotations).reshape(-1, 16), _f32(translations).reshape(-1, 3)
    tid = np.ascontiguousarray(translation_ids, np.int64) if translation_ids is not None else None
    assert tid is None or len(tid) == len(t)
    opts = ProposeOpts(max_n_neigh, radius, (C.c_float * 3)(*thresholds), top_k, tid.ctypes.data if tid is not None else None)
    cap = int(cap if cap is not None else max(len(t), 1))
    while True:
        out = np.zeros((cap, POSE_FLOATS), np.float32)
        ids = np.zeros(cap, np.int64)
        n = C.c_int64(0)
        _check(lib().rsgpu_propose_poses(obj_lvl4.h, obj_lvl3.h, obj_lvl2.h, scene.h, _ptr(r), len(r), _ptr(t), len(t),
                                         C.byref(opts), _ptr(out), _ptr(ids), cap, C.byref(n)))
        if n.value <= cap:
            return out[: n.value].copy(), ids[: n.value].copy()
        cap = int(n.value)


def icp_align(obj: PointCloud, scan: HashGrid, T1, max_dist, max_angle, T2=None, max_iter=0):
    """icp_align for a batch of starting poses -> (T1 refined [B,16], err [B], iters [B])"""
    T = _f32(T1).reshape(-1, 16).copy()
    T2 = _f32(T2 if T2 is not None else np.eye(4).reshape(16)).reshape(16)
    err = np.zeros(len(T), np.float32)
    it = np.zeros(len(T), np.int32)
    _check(lib().rsgpu_icp_align_batch_ex(obj.h, scan.h, _ptr(T), len(T), _ptr(T2), max_dist, max_angle, max_iter,
                                          _ptr(err), _ptr(it)))
    return T, err, it


def icp_align_multi(objects, scan: HashGrid, T1_list, max_dist, max_angle, T2=None):
    """icp_align for several objects against one scan in a single launch -> list of (T1 [B,16], err [B], iters [B])"""
    T2 = _f32(T2 if T2 is not None else np.eye(4).reshape(16)).reshape(16)
    outs = []
    jobs = (IcpJob * len(objects))()
    for j, (o, T1) in enumerate(zip(objects, T1_list)):
        T = _f32(T1).reshape(-1, 16).copy()
        err = np.zeros(len(T), np.float32)
        it = np.zeros(len(T), np.int32)
        outs.append((T, err, it))
        jobs[j] = IcpJob(o.h.value, T.ctypes.data, len(T), err.ctypes.data, it.ctypes.data)
    _check(lib().rsgpu_icp_align_multi(jobs, len(objects), scan.h, _ptr(T2), max_dist, max_angle))
    return outs


def assign_labels(scan_pos, scan_nor, poses, object_grids, first, last, radius, labels, min_dists):
    """rspf__assign_temporary_labels over placements [first, last); labels (int8) / min_dists (float32) updated in place"""
    sp, sn = _f32(scan_pos).reshape(-1, 3), _f32(scan_nor).reshape(-1, 3)
    ps = _f32(poses).reshape(-1, 16)
    assert labels.dtype == np.int8 and min_dists.dtype == np.float32 and labels.flags.c_contiguous
    gh = (C.c_void_p * len(object_grids))(*[g.h for g in object_grids])
    _check(lib().rsgpu_assign_labels(_ptr(sp), _ptr(sn), len(sp), _ptr(ps), gh, first, last, radius, _ptr(labels),
                                     _ptr(min_dists)))


def unary_costs(labels, label_is_static, n_labels):
    lab = np.ascontiguousarray(labels, np.int32)
    st = np.ascontiguousarray(label_is_static, np.uint8)
    out = np.zeros((len(lab), n_labels), np.int32)
    _check(lib().rsgpu_unary_costs(_ptr(lab), _ptr(st), len(lab), n_labels, _ptr(out)))
    return out


def neighborhood(grid: HashGrid, pos, nor, max_nn=8, radius_sq=np.float32(0.05) * np.float32(0.05), dist_exp=15.0,
                 angle_exp=16.0):
    p, n = _f32(pos).reshape(-1, 3), _f32(nor).reshape(-1, 3)
    nbr = np.zeros((len(p), max_nn), np.int32)
    w = np.zeros((len(p), max_nn), np.float32)
    _check(lib().rsgpu_neighborhood(grid.h, _ptr(p), _ptr(n), len(p), max_nn, radius_sq, dist_exp, angle_exp, _ptr(nbr),
                                    _ptr(w)))
    return nbr, w


def overlap_factors(obj_lvl3: PointCloud, obj_lvl1: PointCloud, pose_ref, poses, voxel=0.1, inside=1, normalize_by_smaller=0):
    """isect_get_overlap_factor (lib/rs/intersect.h:309) of one object under ``pose_ref`` against every pose of ``poses``"""
    pr = _f32(pose_ref).reshape(16)
    ps = _f32(poses).reshape(-1, 16)
    out = np.zeros(len(ps), np.float32)
    _check(lib().rsgpu_overlap_factors(obj_lvl3.h, obj_lvl1.h, _ptr(pr), _ptr(ps), len(ps), voxel, inside, normalize_by_smaller, _ptr(out)))
    return out


def non_maxima_suppression(obj_lvl3: PointCloud, obj_lvl1: PointCloud, centroid, proposals, dist_threshold=0.2):
    """mgs_non_maxima_suppresion (apps/pose_proposal/pose_proposal.h:56) for one object -> keep flags [n]"""
    pr = _f32(proposals).reshape(-1, POSE_FLOATS)
    c = _f32(centroid).reshape(3)
    keep = np.zeros(len(pr), np.uint8)
    _check(lib().rsgpu_nms(obj_lvl3.h, obj_lvl1.h, _ptr(c), _ptr(pr), len(pr), dist_threshold, _ptr(keep)))
    return keep.astype(bool)


# ------------------------------------------------------------------------------------------------ level building
LEVEL_VOXEL = (0.005, 0.01, 0.02, 0.04, 0.08)  # rs_pointcloud_init (reference lib/rs/rs_pointcloud.h:145)


def level_max_n_neigh(level):
    """the k of the reference's sampling search: (size_t)(1024 * (level / 4.0f)), 256 when 0 (rs_pointcloud.h:994-995)"""
    k = int(np.float32(1024.0) * (np.float32(level) / np.float32(4.0)))
    return k if k else 256


def poisson_level(pos0, level, voxel=None, max_n_neigh=None, return_rounds=False):
    """rs_pointcloud__compute_level_poisson's sample selection -> ascending level-0 indices of the level's points"""
    p = _f32(pos0).reshape(-1, 3)
    voxel = np.float32(LEVEL_VOXEL[level] if voxel is None else voxel)
    k = level_max_n_neigh(level) if max_n_neigh is None else int(max_n_neigh)
    out = np.zeros(len(p), np.int32)
    n, rounds = C.c_int32(0), C.c_int32(0)
    _check(lib().rsgpu_poisson_level(_ptr(p), len(p), voxel, k, _ptr(out), C.byref(n), C.byref(rounds)))
    idx = out[: n.value].copy()
    return (idx, rounds.value) if return_rounds else idx


def compute_levels(pos0, nor0):
    """rs_pointcloud_compute_levels (rs_pointcloud.h:1305-1316) without the per-level grids: levels 1-4 are the rows of
    level 0 picked by poisson_level -> list of five (pos, nor) pairs"""
    p, n = _f32(pos0).reshape(-1, 3), _f32(nor0).reshape(-1, 3)
    levels = [(p, n)]
    for lvl in range(1, 5):
        idx = poisson_level(p, lvl)
        levels.append((np.ascontiguousarray(p[idx]), np.ascontiguousarray(n[idx])))
    return levels


# ------------------------------------------------------------------------------------------------ coverage term (8 f3)
def coverage_grid(bbox_min, bbox_max, voxel=0.05):
    """isect_grid3d_init (reference lib/rs/intersect.h:57-75) over a bbox -> (res int32[3], origin float32[3]): the bbox
    fattened by 0.3, resolution ceilf( extent / voxel ) + 1 per axis - host arithmetic in float32, like the reference"""
    mn = _f32(bbox_min).reshape(3) - np.float32(0.3)
    mx = _f32(bbox_max).reshape(3) + np.float32(0.3)
    res = (np.ceil((mx - mn) / np.float32(voxel)).astype(np.int32) + 1).astype(np.int32)
    return res, mn.astype(np.float32)


def rasterize_points(pts, pose, res, origin, voxel=0.05, grid=None):
    """rsao_rasterize_scene_to_grid's loop: cells lit by the points (under `pose` when given) -> uint8 grid (OR-ed into `grid`)"""
    p = _f32(pts).reshape(-1, 3)
    res = np.ascontiguousarray(res, np.int32)
    if grid is None:
        grid = np.zeros(int(res[0]) * int(res[1]) * int(res[2]), np.uint8)
    ps = _f32(pose).reshape(16) if pose is not None else None
    _check(lib().rsgpu_rasterize_points(_ptr(p), len(p), _ptr(ps) if ps is not None else None, _ptr(_f32(origin).reshape(3)), _ptr(res),
                                        np.float32(voxel), _ptr(grid)))
    return grid


def coverage_masks(objects_lvl2, poses, res, origin, scene_grid, voxel=0.05):
    """bit masks of the candidate placements over the scan's lit cells -> (masks uint32 [n, n_words], n_lit)"""
    ps = _f32(poses).reshape(-1, 16)
    assert len(ps) == len(objects_lvl2)
    res = np.ascontiguousarray(res, np.int32)
    sg = np.ascontiguousarray(scene_grid, np.uint8)
    n_lit = int(np.count_nonzero(sg))
    n_words = (n_lit + 31) // 32
    masks = np.zeros((len(ps), n_words), np.uint32)
    oh = (C.c_void_p * max(len(ps), 1))(*[o.h for o in objects_lvl2])
    got = C.c_int32(0)
    _check(lib().rsgpu_coverage_masks(oh, _ptr(ps), len(ps), _ptr(_f32(origin).reshape(3)), _ptr(res), np.float32(voxel), _ptr(sg),
                                      _ptr(masks) if masks.size else None, n_words, C.byref(got)))
    assert got.value == n_lit
    return masks, n_lit


def coverage_score(masks, n_lit):
    """rsao__compute_scene_coverage_score of the arrangement formed by the placements whose masks are given"""
    if n_lit == 0 or len(masks) == 0:
        return np.float32(0.0)
    u = np.bitwise_or.reduce(np.ascontiguousarray(masks, np.uint32), axis=0)
    agree = int(np.unpackbits(u.view(np.uint8)).sum())
    return np.float32(agree) / np.float32(n_lit)


# ------------------------------------------------------------------------------------------------ plane detection
def plane_inlier_counts(pts, active, planes, dist_threshold):
    """evaluate_plane_model (reference lib/rs/rs_pointcloud_filters.cpp:117-134) for one RANSAC round's candidate planes
    [P, 6] = {center, normal} over the points still active -> int32 counts [P]"""
    p = _f32(pts).reshape(-1, 3)
    a = np.ascontiguousarray(active, np.uint8)
    assert len(a) == len(p)
    pl = _f32(planes).reshape(-1, 6)
    out = np.zeros(len(pl), np.int32)
    _check(lib().rsgpu_plane_inlier_counts(_ptr(p), _ptr(a), len(p), _ptr(pl), len(pl), np.float32(dist_threshold), _ptr(out)))
    return out
