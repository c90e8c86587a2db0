"""Writers / readers for the reference's on-disk formats, so that synthetic scenes can be handed to the reference's
own executables (and to the drop-in build of integration/): binary PLY clouds with the 12 vertex properties
rs_pointcloud__load_ply reads (reference lib/rs/rs_pointcloud.h:598-781), the `.rsdb` text database
(lib/rs/rs_database.h:291-441 parser, :532-611 writer) and the proposal `.bin` (apps/pose_proposal/main.cpp:61-89).
Host-side only: numpy, no GPU, no oracle.
"""
from __future__ import annotations

import os
import numpy as np

from . import synth


def write_ply(path, pos, nor, class_idx=0, instance_idx=0, radius=0.01):
    """face-less binary little-endian PLY: level 0 is then exactly these points (no area resampling, :1268-1281)"""
    pos, nor = np.asarray(pos, np.float32), np.asarray(nor, np.float32)
    n = len(pos)
    dt = np.dtype([("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("nx", "<f4"), ("ny", "<f4"), ("nz", "<f4"),
                   ("red", "u1"), ("green", "u1"), ("blue", "u1"), ("radius", "<f4"), ("class_idx", "<i4"), ("instance_idx", "<i4")])
    v = np.zeros(n, dt)
    v["x"], v["y"], v["z"] = pos[:, 0], pos[:, 1], pos[:, 2]
    v["nx"], v["ny"], v["nz"] = nor[:, 0], nor[:, 1], nor[:, 2]
    v["red"] = v["green"] = v["blue"] = 128
    v["radius"] = radius
    v["class_idx"] = class_idx
    v["instance_idx"] = instance_idx
    hdr = ["ply", "format binary_little_endian 1.0", f"element vertex {n}",
           "property float x", "property float y", "property float z",
           "property float nx", "property float ny", "property float nz",
           "property uchar red", "property uchar green", "property uchar blue",
           "property float radius", "property int class_idx", "property int instance_idx", "end_header"]
    with open(path, "wb") as f:
        f.write(("\n".join(hdr) + "\n").encode("ascii"))
        f.write(v.tobytes())


def pose_to_rsdb_row_major(m4):
    """the 16 numbers of a `pose` line: row-major although msh_mat4_t is column-major in memory (rs_database.h:601-606)"""
    return " ".join(f"{float(x):.9g}" for x in np.asarray(m4, np.float32).reshape(4, 4).reshape(-1))


def write_database(folder, name, scene: synth.Scene, scan_ply, placements=None):
    """`<folder>/<name>.rsdb` + `<folder>/<name>/obj_XXX.ply` object models; returns the .rsdb path.
    placements: optional list of (object_idx, 4x4 pose) forming arrangement 0 (the previous scan's arrangement)."""
    model_folder = os.path.join(folder, name)
    os.makedirs(model_folder, exist_ok=True)
    lines = ["rsdb 1.0", f"model_folder {model_folder}"]
    for i, c in enumerate(synth.CLASS_NAMES):
        lines.append(f"class {c} {i}")
    lines.append(f"scene 0 0 {scan_ply} none")
    for i, o in enumerate(scene.objects):
        fn = f"obj_{o.uidx:03d}.ply"
        write_ply(os.path.join(model_folder, fn), o.cloud.pos(0), o.cloud.nor(0), o.class_idx, o.uidx)
        lines.append(f"object {fn} {o.uidx} {o.class_idx}")
    lines.append("n_arrangements 1")
    for k, (oi, pose) in enumerate(placements or []):
        lines.append(f"pose {k} 0 {oi} 1.0  {pose_to_rsdb_row_major(pose)}")
    path = os.path.join(folder, name + ".rsdb")
    with open(path, "w") as f:
        f.write("\n".join(lines) + "\n")
    return path


def read_proposals(path):
    """proposal .bin -> list (per object) of float32 [n, 17] (column-major xform + score), as written (descending score)"""
    raw = np.fromfile(path, np.uint8)
    n_obj = int(raw[:4].view("<i4")[0])
    counts = raw[4:4 + 4 * n_obj].view("<i4")
    body = raw[4 + 4 * n_obj:].view("<f4").reshape(-1, 17)
    out, o = [], 0
    for c in counts:
        out.append(body[o:o + c].copy())
        o += c
    return out


def write_proposals(path, per_object):
    """list (per database object, static ones included) of float32 [n, 17] -> proposal .bin as save_pose_proposals writes it
    (reference apps/pose_proposal/main.cpp:61-89; readers: apps/segment_transfer/main.cpp:143-193, rsdb_viewer):
    int32 n_objects, int32 count[n_objects], then per object count x {16 float32 column-major xform, float32 score}"""
    per_object = [np.ascontiguousarray(p, "<f4").reshape(-1, 17) for p in per_object]
    with open(path, "wb") as f:
        f.write(np.array([len(per_object)], "<i4").tobytes())
        f.write(np.array([len(p) for p in per_object], "<i4").tobytes())
        for p in per_object:
            f.write(p.tobytes())
    return path


_PLY_TYPES = {"char": "i1", "uchar": "u1", "int8": "i1", "uint8": "u1", "short": "i2", "ushort": "u2", "int16": "i2", "uint16": "u2",
              "int": "i4", "uint": "u4", "int32": "i4", "uint32": "u4", "float": "f4", "float32": "f4", "double": "f8", "float64": "f8"}


def read_ply(path):
    """vertex element of a PLY (binary little-endian or ASCII, the two the reference reads and writes:
    lib/rs/rs_pointcloud.h:598-836) -> structured array with the file's own property names.  Faces are ignored (a scan with
    faces is area-resampled by the reference, :1268-1281; the fixtures are face-less)."""
    with open(path, "rb") as f:
        if f.readline().strip() != b"ply":
            raise ValueError(f"{path}: not a PLY file")
        fmt, n_vertex, props, in_vertex = None, 0, [], False
        while True:
            line = f.readline()
            if not line:
                raise ValueError(f"{path}: header without end_header")
            tok = line.decode("ascii", "replace").split()
            if not tok or tok[0] == "comment":
                continue
            if tok[0] == "format":
                fmt = tok[1]
            elif tok[0] == "element":
                in_vertex = tok[1] == "vertex"
                if in_vertex:
                    n_vertex = int(tok[2])
            elif tok[0] == "property" and in_vertex:
                if tok[1] == "list":
                    raise ValueError(f"{path}: list property on the vertex element")
                props.append((tok[2], _PLY_TYPES[tok[1]]))
            elif tok[0] == "end_header":
                break
        if fmt == "binary_little_endian":
            dt = np.dtype([(n, "<" + t) for n, t in props])
            return np.frombuffer(f.read(dt.itemsize * n_vertex), dt, n_vertex).copy()
        if fmt == "ascii":
            dt = np.dtype([(n, t) for n, t in props])
            out = np.zeros(n_vertex, dt)
            for i in range(n_vertex):
                vals = f.readline().split()
                for (n, _), v in zip(props, vals):
                    out[n][i] = float(v)
            return out
        raise ValueError(f"{path}: unsupported PLY format {fmt}")


def read_database(path):
    """`.rsdb` text database (parser of the reference: lib/rs/rs_database.h:291-441) -> dict with model_folder, classes
    [(name, idx)], scenes [(uidx, arrangement_idx, scan ply, proposals .bin or None)], objects [(file, uidx, class_idx,
    is_shape_prior)], n_arrangements and poses [(placement uidx, arrangement_idx, object_idx, score, 4x4 float32)].
    The 16 numbers of a pose line are row-major (:601-606); the returned matrix is the mathematical 4x4."""
    db = dict(version=None, model_folder=None, classes=[], scenes=[], objects=[], n_arrangements=0, poses=[])
    with open(path) as f:
        for line in f:
            tok = line.split()
            if not tok:
                continue
            cmd = tok[0]
            if cmd == "rsdb":
                db["version"] = tok[1]
            elif cmd == "model_folder":
                db["model_folder"] = tok[1]
            elif cmd == "class":
                db["classes"].append((tok[1], int(tok[2])))
            elif cmd == "scene":
                db["scenes"].append((int(tok[1]), int(tok[2]), tok[3], None if len(tok) < 5 or tok[4] == "none" else tok[4]))
            elif cmd in ("object", "shape_prior"):
                db["objects"].append((tok[1], int(tok[2]), int(tok[3]), cmd == "shape_prior"))
            elif cmd == "n_arrangements":
                db["n_arrangements"] = int(tok[1])
            elif cmd == "pose":
                m = np.array([float(x) for x in tok[5:21]], np.float32).reshape(4, 4)
                db["poses"].append((int(tok[1]), int(tok[2]), int(tok[3]), float(tok[4]), m))
    if db["version"] is None:
        raise ValueError(f"{path}: no 'rsdb <version>' line")
    return db
