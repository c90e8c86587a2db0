"""Writes tests/golden/coverage_golden.npz from the UNMODIFIED reference's grid primitives compiled in place (oracle/_ref:
isect_grid3d_init, msh_mat4_vec3_mul, isect_grid3d_cell_from_world_space driven like rsao_rasterize_scene_to_grid /
rsao__rasterize_arrangement_to_grid, apps/segment_transfer/arrangement_optimization.cpp:1064-1106): the grid of a small
scan, its lit cells, and the lit cells of a few posed objects.  Build container only:

    python tests/golden/make_golden_coverage.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import refbind as R  # noqa: E402
from rescan_b200 import synth  # noqa: E402


def main():
    scene = synth.make_scene(**synth.CONFIGS["tiny"]["scene"])
    mn, mx = scene.scan.pos(0).min(0), scene.scan.pos(0).max(0)
    out = dict(bbox_min=mn, bbox_max=mx, scan_pos2=scene.scan.pos(2))
    rng = np.random.default_rng(20191027)
    for voxel, tag in ((0.05, "v5"), (0.15, "v15")):
        res, origin, n = R.cov_grid(mn, mx, voxel)
        out[f"{tag}_res"], out[f"{tag}_origin"] = res, origin
        out[f"{tag}_scan_lit"] = np.nonzero(R.cov_rasterize(mn, mx, scene.scan.pos(2), None, n, voxel))[0].astype(np.int32)
    res, origin, n = R.cov_grid(mn, mx, 0.05)
    k = 0
    for oi, o in enumerate(scene.objects):
        out[f"obj{oi}_pos2"] = o.cloud.pos(2)
        for j in range(3):
            d = np.eye(4, dtype=np.float32) if j == 0 else synth.yaw_pose(rng.uniform(0, 6.28), rng.uniform(-2, 2), rng.uniform(-2, 2), rng.uniform(-0.3, 0.3))
            pose = np.ascontiguousarray((d.astype(np.float64) @ o.pose.astype(np.float64)).astype(np.float32).T.reshape(16))
            out[f"pose{k}"], out[f"pose{k}_obj"] = pose, np.array([oi], np.int32)
            out[f"pose{k}_lit"] = np.nonzero(R.cov_rasterize(mn, mx, o.cloud.pos(2), pose, n, 0.05))[0].astype(np.int32)
            k += 1
    out["n_poses"] = np.array([k], np.int32)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "coverage_golden.npz"), **out)
    print("wrote coverage_golden.npz:", k, "poses,", len(out["v5_scan_lit"]), "lit scan cells")


if __name__ == "__main__":
    main()
