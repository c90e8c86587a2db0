"""Deterministic synthetic rescans (there is no dataset offline).

Generates what the reference's loaders would hand to the hot path: a scan point cloud (positions +
unit normals) with the five sampling levels of ``rs_pointcloud_t`` (reference lib/rs/rs_pointcloud.h:77-97,
voxel sizes 5 mm / 1 / 2 / 4 / 8 cm at :145), a set of object models in their own frames (dynamic objects
centred on their xz-centroid and resting on y = 0, as reference apps/seg2rsdb/main.cpp:118-126 leaves
them) and the ground-truth placements.  Shapes follow SURVEY.md §8(d): a box room (floor + 4 walls,
optional ceiling / partition walls) sampled on a jittered lattice with 1 mm Gaussian noise, closed
boxes / L-shapes of 0.3–1.5 m without bottom faces, rejection-sampled non-overlapping.

Level building here is a voxel-thinning stand-in (first point per voxel, ascending level-0 index order
like rs_pointcloud.h:984-1106 keeps it); the reference's greedy Poisson-disk sampler stays host code and is
not on the path, so parity tests that need reference-true levels get them from oracle/_ref instead.

Only numpy; no GPU, no oracle.
"""
from __future__ import annotations

import dataclasses
import numpy as np

SEED = 20191027
LEVEL_VOXEL = (0.005, 0.01, 0.02, 0.04, 0.08)  # rs_pointcloud.h:145
N_LEVELS = 5

# class table shared with oracle/ref_harness.cpp (k_class_names); static ones by NAME (rs_database.h:257-288)
CLASS_NAMES = ("unlabelled", "wall", "floor", "ceiling", "cabinet", "chair", "table", "sofa", "box")
STATIC_CLASSES = frozenset(("unlabelled", "wall", "floor", "ceiling", "cabinet"))
DYNAMIC_CLASS_IDS = tuple(i for i, n in enumerate(CLASS_NAMES) if n not in STATIC_CLASSES)


@dataclasses.dataclass
class Cloud:
    """Five-level SoA cloud: levels[l] = (positions float32 [n,3], normals float32 [n,3])."""
    levels: list
    bbox_min: np.ndarray
    bbox_max: np.ndarray

    def pos(self, lvl):
        return self.levels[lvl][0]

    def nor(self, lvl):
        return self.levels[lvl][1]

    def n(self, lvl):
        return self.levels[lvl][0].shape[0]


@dataclasses.dataclass
class SceneObject:
    uidx: int
    class_idx: int
    cloud: Cloud
    pose: np.ndarray  # ground-truth 4x4 (float32, row/col = math convention) in the CURRENT scan

    @property
    def is_static(self):
        return CLASS_NAMES[self.class_idx] in STATIC_CLASSES


@dataclasses.dataclass
class Scene:
    scan: Cloud
    objects: list
    room: tuple
    spacing: float
    scan_class: np.ndarray  # level-0 per-point ground-truth class / instance (for label tests)
    scan_instance: np.ndarray


def _plane(rng, origin, eu, ev, lu, lv, normal, spacing, noise):
    nu = max(int(np.ceil(lu / spacing)), 1)
    nv = max(int(np.ceil(lv / spacing)), 1)
    u = (np.arange(nu) + 0.5) * (lu / nu)
    v = (np.arange(nv) + 0.5) * (lv / nv)
    uu, vv = np.meshgrid(u, v, indexing="ij")
    uu = uu.ravel() + rng.uniform(-0.3, 0.3, uu.size) * (lu / nu)
    vv = vv.ravel() + rng.uniform(-0.3, 0.3, vv.size) * (lv / nv)
    p = origin[None, :] + uu[:, None] * eu[None, :] + vv[:, None] * ev[None, :]
    p = p + rng.normal(0.0, noise, p.shape)
    n = np.broadcast_to(np.asarray(normal, np.float64), p.shape)
    return p, n


def _box_faces(rng, lo, hi, spacing, noise, bottom=False, inward=False):
    """Sampled faces of an axis-aligned box; normals outward (or inward for a room)."""
    lo = np.asarray(lo, np.float64)
    hi = np.asarray(hi, np.float64)
    d = hi - lo
    ex, ey, ez = np.eye(3)
    s = -1.0 if inward else 1.0
    faces = [
        (lo, ey, ez, d[1], d[2], -s * ex),
        (np.array([hi[0], lo[1], lo[2]]), ey, ez, d[1], d[2], s * ex),
        (lo, ex, ey, d[0], d[1], -s * ez),
        (np.array([lo[0], lo[1], hi[2]]), ex, ey, d[0], d[1], s * ez),
        (np.array([lo[0], hi[1], lo[2]]), ex, ez, d[0], d[2], s * ey),
    ]
    if bottom:
        faces.append((lo, ex, ez, d[0], d[2], -s * ey))
    ps, ns = [], []
    for o, eu, ev, lu, lv, nrm in faces:
        p, n = _plane(rng, o, eu, ev, lu, lv, nrm, spacing, noise)
        ps.append(p)
        ns.append(n)
    return np.concatenate(ps), np.concatenate(ns)


def thin_levels(pos, nor):
    """levels 1..4 = first point (ascending index) per voxel of LEVEL_VOXEL[l]; level 0 = input."""
    levels = [(np.ascontiguousarray(pos, np.float32), np.ascontiguousarray(nor, np.float32))]
    p64 = levels[0][0].astype(np.float64)
    base = p64.min(axis=0) if len(p64) else np.zeros(3)
    for lvl in range(1, N_LEVELS):
        if len(p64) == 0:
            levels.append(levels[0])
            continue
        key = np.floor((p64 - base) / LEVEL_VOXEL[lvl]).astype(np.int64)
        dims = key.max(axis=0) + 1
        lin = (key[:, 2] * dims[1] + key[:, 1]) * dims[0] + key[:, 0]
        _, first = np.unique(lin, return_index=True)
        first.sort()
        levels.append((np.ascontiguousarray(levels[0][0][first]), np.ascontiguousarray(levels[0][1][first])))
    return levels


def perturb_normals(rng, nor, sigma):
    """scanned normals are estimated, never exact: isotropic Gaussian tilt of `sigma` radians, renormalised"""
    n = np.asarray(nor, np.float64) + rng.normal(0.0, sigma, np.shape(nor))
    return n / np.linalg.norm(n, axis=1, keepdims=True)


def make_cloud(pos, nor, levels=None):
    pos = np.ascontiguousarray(pos, np.float32)
    nor = np.ascontiguousarray(nor, np.float32)
    lv = levels if levels is not None else thin_levels(pos, nor)
    return Cloud(lv, pos.min(axis=0), pos.max(axis=0))


def yaw_pose(yaw, tx, tz, ty=0.0):
    c, s = np.cos(yaw), np.sin(yaw)
    m = np.eye(4, dtype=np.float64)
    m[0, 0], m[0, 2], m[2, 0], m[2, 2] = c, s, -s, c
    m[0, 3], m[1, 3], m[2, 3] = tx, ty, tz
    return m.astype(np.float32)


def apply_pose(m, pos, nor):
    m = m.astype(np.float64)
    return pos @ m[:3, :3].T + m[:3, 3], nor @ m[:3, :3].T


def _make_object_model(rng, spacing, noise):
    """closed box or L-shape, 0.3-1.5 m, no bottom face, centred in xz, resting on y=0"""
    w, h, d = rng.uniform(0.3, 1.5), rng.uniform(0.3, 1.2), rng.uniform(0.3, 1.5)
    p, n = _box_faces(rng, (-w / 2, 0, -d / 2), (w / 2, h, d / 2), spacing, noise)
    if rng.uniform() < 0.4:  # L-shape: add a taller back part
        w2, h2, d2 = w, h + rng.uniform(0.2, 0.6), d * rng.uniform(0.2, 0.4)
        p2, n2 = _box_faces(rng, (-w2 / 2, 0, -d / 2), (w2 / 2, h2, -d / 2 + d2), spacing, noise)
        inside2 = (np.abs(p[:, 0]) < w2 / 2 - 1e-3) & (p[:, 1] < h2 - 1e-3) & (p[:, 2] < -d / 2 + d2 - 1e-3) & (p[:, 2] > -d / 2 + 1e-3)
        inside1 = (np.abs(p2[:, 0]) < w / 2 - 1e-3) & (p2[:, 1] < h - 1e-3) & (np.abs(p2[:, 2]) < d / 2 - 1e-3)
        p, n = np.concatenate([p[~inside2], p2[~inside1]]), np.concatenate([n[~inside2], n2[~inside1]])
    c = p.mean(axis=0)
    p = p - np.array([c[0], 0.0, c[2]])
    radius = float(np.sqrt((p[:, 0] ** 2 + p[:, 2] ** 2).max()))
    return p, n, radius


def make_scene(n_objects=10, n_static=2, room=(7.0, 2.6, 5.0), spacing=0.025, noise=0.001, ceiling=False,
               partitions=0, seed=SEED, target_points=None, objects=None, normal_noise=0.03):
    """One synthetic scan.  ``objects`` re-uses models of a previous scan with fresh poses (a "rescan")."""
    rng = np.random.default_rng(seed)
    if target_points is not None:
        # one trial pass to calibrate the lattice spacing to the requested level-0 size
        trial = make_scene(n_objects, n_static, room, 0.05, noise, ceiling, partitions, seed, None, objects, normal_noise)
        spacing = 0.05 * np.sqrt(trial.scan.n(0) / float(target_points))
        rng = np.random.default_rng(seed)
    rx, ry, rz = room
    ps, ns, cls, ins = [], [], [], []

    def add(p, n, c, i):
        ps.append(p)
        ns.append(n)
        cls.append(np.full(len(p), c, np.int32))
        ins.append(np.full(len(p), i, np.int32))

    wall_c, floor_c, ceil_c = CLASS_NAMES.index("wall"), CLASS_NAMES.index("floor"), CLASS_NAMES.index("ceiling")
    ex, ey, ez = np.eye(3)
    o = np.zeros(3)
    add(*_plane(rng, o, ex, ez, rx, rz, ey, spacing, noise), floor_c, 1)
    add(*_plane(rng, o, ey, ez, ry, rz, ex, spacing, noise), wall_c, 2)
    add(*_plane(rng, np.array([rx, 0, 0.0]), ey, ez, ry, rz, -ex, spacing, noise), wall_c, 3)
    add(*_plane(rng, o, ex, ey, rx, ry, ez, spacing, noise), wall_c, 4)
    add(*_plane(rng, np.array([0, 0, rz]), ex, ey, rx, ry, -ez, spacing, noise), wall_c, 5)
    if ceiling:
        add(*_plane(rng, np.array([0, ry, 0.0]), ex, ez, rx, rz, -ey, spacing, noise), ceil_c, 6)
    blocked = []
    for k in range(partitions):  # two-sided partition walls parallel to z, leaving a door gap
        x = rx * (k + 1) / (partitions + 1)
        z0, z1 = (0.0, rz * 0.7) if k % 2 == 0 else (rz * 0.3, rz)
        add(*_plane(rng, np.array([x - 0.05, 0, z0]), ey, ez, ry, z1 - z0, -ex, spacing, noise), wall_c, 7 + k)
        add(*_plane(rng, np.array([x + 0.05, 0, z0]), ey, ez, ry, z1 - z0, ex, spacing, noise), wall_c, 7 + k)
        blocked.append((x, z0, z1))

    objs = []
    placed = []
    dyn = list(DYNAMIC_CLASS_IDS)
    for i in range(n_objects):
        if objects is not None:
            prev = objects[i]
            mp, mn = prev.cloud.pos(0).astype(np.float64), prev.cloud.nor(0).astype(np.float64)
            radius = float(np.sqrt((mp[:, 0] ** 2 + mp[:, 2] ** 2).max()))
            class_idx, uidx, cloud = prev.class_idx, prev.uidx, prev.cloud
            keep_pose = rng.uniform() < 0.3
        else:
            mp, mn, radius = _make_object_model(rng, spacing, noise)
            class_idx = CLASS_NAMES.index("cabinet") if i < n_static else dyn[i % len(dyn)]
            uidx, cloud, keep_pose = 100 + i, None, False
        for _ in range(1000):
            if keep_pose:
                yaw = float(np.arctan2(-prev.pose[2, 0], prev.pose[0, 0]))
                tx, tz = float(prev.pose[0, 3]), float(prev.pose[2, 3])
            else:
                yaw = rng.uniform(0, 2 * np.pi)
                tx, tz = rng.uniform(radius + 0.05, rx - radius - 0.05), rng.uniform(radius + 0.05, rz - radius - 0.05)
            ok = all((tx - a) ** 2 + (tz - b) ** 2 > (radius + r + 0.1) ** 2 for a, b, r in placed)
            ok = ok and all(abs(tx - x) > radius + 0.15 for x, _, _ in blocked)
            if ok:
                break
            keep_pose = False
        placed.append((tx, tz, radius))
        pose = yaw_pose(yaw, tx, tz)
        if cloud is None:
            cloud = make_cloud(mp, perturb_normals(rng, mn, normal_noise))
        objs.append(SceneObject(uidx, class_idx, cloud, pose))
        # the scan sees a fresh sampling of the same surfaces, not the model's own points
        sp, sn, _ = (mp, mn, None) if objects is None else (mp, mn, None)
        jitter = rng.normal(0.0, noise, sp.shape)
        wp, wn = apply_pose(pose, sp + jitter, sn)
        add(wp, wn, class_idx, uidx)

    pos = np.concatenate(ps)
    nor = np.concatenate(ns)
    nor = perturb_normals(rng, nor, normal_noise)
    perm = rng.permutation(len(pos))  # scans are not surface-ordered
    pos, nor = pos[perm], nor[perm]
    scan = make_cloud(pos, nor)
    return Scene(scan, objs, room, float(spacing), np.concatenate(cls)[perm], np.concatenate(ins)[perm])


def rotation_angles(n_rot):
    """float32-accumulated yaw angles exactly like `for(y=0; y<2pi; y+=inc)` (pose_proposal.cpp:219)"""
    inc = np.float32(np.float32(6.2831853072) / np.float32(n_rot))
    out, a = [], np.float32(0.0)
    while a < np.float32(6.2831853072):
        out.append(a)
        a = np.float32(a + inc)
    return np.asarray(out, np.float32)


def translation_seeds(scan: Cloud, n_seeds, seed=SEED + 1, level=3):
    """T seeds = seeded sample of scan level-`level` points projected to y = 0."""
    rng = np.random.default_rng(seed)
    p = scan.pos(level)
    idx = rng.choice(len(p), size=n_seeds, replace=len(p) < n_seeds)
    t = p[idx].copy()
    t[:, 1] = 0.0
    return np.ascontiguousarray(t, np.float32)


def reference_translation_grid(scan: Cloud, spacing=0.10):
    """The reference's own xz lattice over the scan bbox +- one step, y = 0, ox-major / oz-minor, with
    its float32-accumulating loop counters (pose_proposal.cpp:203-222)."""
    sp = np.float32(spacing)
    lx = np.float32(scan.bbox_max[0] - scan.bbox_min[0])
    lz = np.float32(scan.bbox_max[2] - scan.bbox_min[2])
    ox_list, oz_list = [], []
    ox = np.float32(-sp)
    while ox < np.float32(lx + sp):
        ox_list.append(ox)
        ox = np.float32(ox + sp)
    oz = np.float32(-sp)
    while oz < np.float32(lz + sp):
        oz_list.append(oz)
        oz = np.float32(oz + sp)
    t = np.zeros((len(ox_list) * len(oz_list), 3), np.float32)
    k = 0
    for ox in ox_list:
        for oz in oz_list:
            t[k, 0] = np.float32(scan.bbox_min[0]) + ox
            t[k, 2] = np.float32(scan.bbox_min[2]) + oz
            k += 1
    return t


CONFIGS = {
    # BASELINE.json configs[0]/[1]: ~200 K-point scan, 10 objects, 36 rotations x 2 K translation seeds
    "C2": dict(scene=dict(n_objects=10, n_static=2, room=(7.0, 2.6, 5.0), target_points=200_000),
               n_rot=36, n_seeds=2048),
    # configs[2]: ~2 M points, 60 objects, 72 x 20 K
    "C3": dict(scene=dict(n_objects=60, n_static=8, room=(20.0, 2.8, 15.0), target_points=2_000_000,
                          ceiling=True, partitions=3), n_rot=72, n_seeds=20_000),
    # small case for tests / smoke
    "tiny": dict(scene=dict(n_objects=3, n_static=1, room=(3.0, 2.0, 2.5), spacing=0.03), n_rot=8, n_seeds=64),
}
