"""Writes tests/golden/rescan_golden.npz from the UNMODIFIED reference (oracle/_ref/librescan_ref.so, compiled in
place from /root/reference by oracle/Makefile).  Run in the build container only:

    python tests/golden/make_golden.py

The fixture holds small seeded inputs and the reference's own outputs for every function on the hot path, so the
oracle (and the GPU path) can be pinned on machines where /root/reference does not exist.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import refbind as R  # noqa: E402
from rescan_b200 import synth, posegrid  # noqa: E402


def colmajor(m4):
    return np.ascontiguousarray(np.asarray(m4, np.float32).T.reshape(16))


def main():
    assert R.available(), "build oracle/_ref first (make -f oracle/Makefile ref)"
    out = {}
    scene = synth.make_scene(n_objects=3, n_static=1, room=(3.2, 1.6, 2.8), spacing=0.035, seed=synth.SEED + 99)
    rng = np.random.default_rng(424242)
    # only level 0 is stored: levels 1-4 are synth.thin_levels(level 0), a pure function the tests re-run
    out["scan_pos0"], out["scan_nor0"] = scene.scan.pos(0), scene.scan.nor(0)
    for i, o in enumerate(scene.objects):
        out[f"obj{i}_pos0"], out[f"obj{i}_nor0"] = o.cloud.pos(0), o.cloud.nor(0)
        out[f"obj{i}_pose"] = colmajor(o.pose)
        out[f"obj{i}_meta"] = np.array([o.uidx, o.class_idx, int(o.is_static)], np.int32)
    out["n_objects"] = np.array([len(scene.objects)], np.int32)

    scan = R.RefCloud.from_levels({l: (scene.scan.pos(l), scene.scan.nor(l)) for l in range(5)})
    objs = [R.RefCloud.from_levels({l: (o.cloud.pos(l), o.cloud.nor(l)) for l in range(5)}) for o in scene.objects]
    p1 = scene.scan.pos(1)

    # ---- hash grid + searches (msh_hash_grid.h)
    g = scan.grid(1)
    info = g.info()
    out["grid_dims"] = info["dims"]
    out["grid_cell"] = np.array([info["cell_size"], info["inv_cell_size"]], np.float64)
    out["grid_minmax"] = np.concatenate([info["min_pt"], info["max_pt"]])
    out["grid_counts"] = np.array([info["n_pts"], info["n_bins"], info["max_n_pts_in_bin"]], np.int64)
    xyz, idx = g.data()
    out["grid_data_idx"] = idx
    q = p1[rng.choice(len(p1), 400)] + rng.uniform(-0.03, 0.03, (400, 3)).astype(np.float32)
    q[:10] += 4.0
    out["queries"] = np.ascontiguousarray(q, np.float32)
    for tag, r, k in (("a", 0.10, 64), ("b", 0.05, 16), ("c", 0.05, 1), ("d", 0.075, 8)):
        i, d, n, t = g.radius_search(out["queries"], r, k)
        m = np.arange(k)[None, :] < n[:, None]
        out[f"rs_{tag}_idx"], out[f"rs_{tag}_d2"], out[f"rs_{tag}_n"] = np.where(m, i, -1), np.where(m, d, 0).astype(np.float32), n
        out[f"rs_{tag}_param"] = np.array([r, k, t], np.float64)
    qk = p1[rng.choice(len(p1), 300)] + rng.uniform(-0.005, 0.005, (300, 3)).astype(np.float32)
    out["knn_queries"] = np.ascontiguousarray(qk, np.float32)
    i, d, n, t = g.knn_search(out["knn_queries"], 8)
    out["knn_idx"], out["knn_d2"], out["knn_n"] = i, d, n

    # ---- pose scoring (pose_proposal.cpp:93-158)
    poses, which, lvls = [], [], []
    for oi, o in enumerate(scene.objects):
        for j in range(24):
            d = synth.yaw_pose(rng.uniform(-0.2, 0.2), rng.uniform(-0.08, 0.08), rng.uniform(-0.08, 0.08), rng.uniform(-0.01, 0.01))
            m = (d.astype(np.float64) @ o.pose.astype(np.float64)).astype(np.float32) if j % 4 else synth.yaw_pose(
                rng.uniform(0, 6.28), rng.uniform(0, 3.2), rng.uniform(0, 2.8))
            poses.append(colmajor(m))
            which.append(oi)
            lvls.append((4, 3, 2, 1)[j % 4])
    scores = [R.score(objs[oi], scan, x, query_lvl=l, k=(32 if l == 1 else 64)) for x, oi, l in zip(poses, which, lvls)]
    out["score_poses"], out["score_obj"], out["score_lvl"] = np.stack(poses), np.array(which, np.int32), np.array(lvls, np.int32)
    out["score_ref"] = np.array(scores, np.float32)

    # ---- mgs_propose_poses on the reference's own 0.10 m x 10-rotation grid (pose_proposal.cpp:325-369)
    db = R.RefDB()
    for o, c in zip(scene.objects, objs):
        db.add_object(c, o.uidx, o.class_idx)
    props = db.propose_poses(scan)
    out["propose_counts"] = np.array([len(p) for p in props], np.int32)
    out["propose_flat"] = np.concatenate(props) if sum(len(p) for p in props) else np.zeros((0, 17), np.float32)
    mn, mx = scan.bbox()
    out["scan_bbox"] = np.concatenate([mn, mx])

    # ---- icp_align (icp.h:416-500)
    p2, n2 = scene.scan.pos(2), scene.scan.nor(2)
    starts, sw, ends, errs = [], [], [], []
    for oi, o in enumerate(scene.objects):
        for j in range(4):
            d = synth.yaw_pose(rng.uniform(-0.08, 0.08), rng.uniform(-0.03, 0.03), rng.uniform(-0.03, 0.03), rng.uniform(-0.01, 0.01))
            s = colmajor((d.astype(np.float64) @ o.pose.astype(np.float64)).astype(np.float32))
            T, e = R.icp_align(o.cloud.pos(2), o.cloud.nor(2), p2, n2, s, 0.10, np.float32(np.deg2rad(60.0)))
            starts.append(s); sw.append(oi); ends.append(T); errs.append(e)
    out["icp_start"], out["icp_obj"], out["icp_end"], out["icp_err"] = np.stack(starts), np.array(sw, np.int32), np.stack(ends), np.array(errs, np.float32)

    # ---- labels, unary terms, 8-NN edges (rs_pointcloud_filters.cpp:674-989) with the recording gco stub
    oi = np.arange(len(scene.objects), dtype=np.int32)
    ui = np.array([o.uidx for o in scene.objects], np.int32)
    ps = np.stack([colmajor(o.pose) for o in scene.objects])
    ci, ii = db.arrangement_to_labels(scan, oi, ui, ps, 0.05)
    out["label_class"], out["label_instance"] = ci, ii
    cap = db.smooth_labels_capture(scan)
    out["gco_n_labels"] = np.array([cap["n_labels"]], np.int32)
    out["gco_data_cost_head"] = cap["data_cost"][:3000].astype(np.int32)  # first 3000 vertices
    out["gco_data_cost_rowsum"] = cap["data_cost"].sum(axis=1).astype(np.int64)
    out["gco_init_labels"] = cap["init_labels"]
    ea, eb, ew = R.compute_neighborhood(scan, 1, 8, np.float32(0.05) * np.float32(0.05), 15.0, 16.0)
    keep = np.minimum(ea, eb) < 3000  # edges touching the first 3000 vertices
    out["edges_a"], out["edges_b"], out["edges_w"] = ea[keep], eb[keep], ew[keep]
    out["edges_total"] = np.array([len(ea)], np.int64)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "rescan_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB;", {k: v.shape for k, v in list(out.items())[:4]})


if __name__ == "__main__":
    main()
