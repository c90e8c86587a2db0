"""A CPU stand-in for rescan_b200.api, ONLY for the host-logic tests of pipeline.run_step (lanes, object groups, exchanges):
deterministic fake kernels whose results depend only on their inputs, so that a pose-sharded run must reproduce the
single-rank run exactly.  Nothing here computes anything Rescan-related."""
import threading

import numpy as np

POSE_FLOATS = 17
_tl = threading.local()
calls = []
_lock = threading.Lock()


class HashGrid:
    def __init__(self, pts=None, radius=0.05, normals=None, device_ptr=None, n_pts=None):
        self.n = len(pts) if pts is not None else n_pts

    def close(self):
        pass


class PointCloud:
    def __init__(self, pos, nor):
        self.pos = np.asarray(pos, np.float32)

    def __len__(self):
        return len(self.pos)


def lane_count():
    return 8


def thread_attach(lane):
    _tl.lane = lane


def set_option(name, value=None):
    pass


def _score_of(obj_key, t):
    """a pseudo-random but input-determined score per (object, translation)"""
    h = np.sin(t[:, 0] * 12.9898 + t[:, 2] * 78.233 + obj_key * 3.7) * 43758.5453
    return (h - np.floor(h)).astype(np.float32)


def propose_poses(l4, l3, l2, scene, rotations, translations, max_n_neigh=64, radius=0.1, thresholds=(0.25, 0.35, 0.4), top_k=0, cap=None,
                  translation_ids=None):
    t = np.asarray(translations, np.float32).reshape(-1, 3)
    ids_t = np.asarray(translation_ids, np.int64) if translation_ids is not None else np.arange(len(t), dtype=np.int64)
    n_rot = len(rotations)
    s = _score_of(len(l4), t)
    r = (np.floor(s * 1000).astype(np.int64)) % n_rot
    emit = s > 0.55
    s = np.where(s > 0.8, s, np.float32(-1.0))  # "verification": most survivors fail
    order = np.argsort(ids_t[emit], kind="stable")  # the caller's translation order
    ids = (ids_t[emit] * n_rot + r[emit])[order]
    props = np.zeros((len(ids), POSE_FLOATS), np.float32)
    props[:, [0, 5, 10, 15]] = 1.0
    props[:, 12:15] = t[emit][order]
    props[:, 16] = s[emit][order]
    if top_k > 0:
        sel = np.lexsort((ids, -props[:, 16].astype(np.float64)))[:top_k]
        props, ids = props[sel], ids[sel]
    with _lock:
        calls.append(("propose", len(l4), len(t)))
    return props, ids


def non_maxima_suppression(l3, l1, centroid, proposals, dist_threshold=0.2):
    p = np.asarray(proposals, np.float32).reshape(-1, POSE_FLOATS)
    keep = np.zeros(len(p), bool)
    done = np.zeros(len(p), bool)
    while not done.all():
        cand = np.where(done, -np.inf, p[:, 16])
        b = int(np.argmax(cand))
        keep[b] = done[b] = True
        d = np.linalg.norm(p[:, 12:15] - p[b, 12:15], axis=1)
        done |= (~done) & ((d < 1.0) | (p[:, 16] < 0.01))
    return keep


def icp_align(obj, scan, T1, max_dist, max_angle, T2=None, max_iter=0):
    T = np.asarray(T1, np.float32).reshape(-1, 16).copy()
    T[:, 12] = np.round(T[:, 12], 1)  # "refinement": snap x to a decimetre
    return T, np.full(len(T), 0.01, np.float32), np.full(len(T), 7, np.int32)


def compute_object_alignment_scores(obj, scene, xforms, max_n_neigh=64, radius=0.1):
    x = np.asarray(xforms, np.float32).reshape(-1, 16)
    return (0.5 + 0.4 * _score_of(len(obj), x[:, 12:15])).astype(np.float32)


class FilePeer:
    """stand-in for rescan_b200.peerx.PeerExchange on the CPU tier: the same slot / use / put / get protocol, the payloads as
    files in a directory every rank sees (put = atomic rename, get = poll), so that the owner-per-object schedule of
    pipeline.run_step can run with world_size > 1 without CUDA IPC"""

    def __init__(self, root, rank, world, n_slots=64, timeout_s=60.0):
        import os
        self.root, self.rank, self.world, self.n_slots, self.timeout_s = root, int(rank), int(world), int(n_slots), timeout_s
        self._seq = [0] * self.n_slots
        os.makedirs(root, exist_ok=True)

    def begin_use(self, slot):
        self._seq[slot % self.n_slots] += 1
        return self._seq[slot % self.n_slots]

    def _path(self, slot, seq, src, dst):
        import os
        return os.path.join(self.root, f"s{slot}_u{seq}_from{src}_to{dst}.bin")

    def put(self, slot, seq, dst_ranks, buf):
        import os
        data = np.ascontiguousarray(buf).view(np.uint8).reshape(-1).tobytes()
        for d in dst_ranks:
            p = self._path(slot, seq, self.rank, d)
            with open(p + ".tmp", "wb") as f:
                f.write(data)
            os.replace(p + ".tmp", p)

    def get(self, slot, seq, src_ranks):
        import os
        import time
        out, t0 = {}, time.time()
        for s in src_ranks:
            p = self._path(slot, seq, s, self.rank)
            while not os.path.exists(p):
                if time.time() - t0 > self.timeout_s:
                    raise TimeoutError(p)
                time.sleep(0.001)
            out[int(s)] = np.fromfile(p, np.uint8)
        return out

    def slot(self, k):
        raise NotImplementedError

    def close(self):
        pass
