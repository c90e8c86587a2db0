"""Writes tests/golden/nms_golden.npz from the UNMODIFIED reference (oracle/_ref/librescan_ref.so): overlap factors
(lib/rs/intersect.h:309-368) and survivors of mgs_non_maxima_suppresion (apps/pose_proposal/pose_proposal.cpp:371-452) for
the objects of the main golden fixture (tests/golden/rescan_golden.npz provides the clouds).  Build container only:

    python tests/golden/make_golden_nms.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import refbind as R  # noqa: E402
from rescan_b200 import synth  # noqa: E402
from tests import common  # noqa: E402


def proposals_for(pose, rng, n=48):
    props = np.zeros((n, 17), np.float32)
    for j in range(n):
        if j % 5 == 4:
            m = synth.yaw_pose(rng.uniform(0, 6.28), rng.uniform(0.5, 2.5), rng.uniform(0.5, 2.0))
        else:
            d = synth.yaw_pose(rng.uniform(-0.6, 0.6), rng.uniform(-0.5, 0.5), rng.uniform(-0.5, 0.5), rng.uniform(-0.02, 0.02))
            m = (d.astype(np.float64) @ pose.astype(np.float64)).astype(np.float32)
        props[j, :16] = common.colmajor(m)
        props[j, 16] = np.float32(rng.uniform(0.3, 0.99)) if j % 7 else np.float32(-1.0)
    props[3, 16] = props[2, 16]
    return props


def main():
    assert R.available()
    z, scan, objs = common.golden()
    rng = np.random.default_rng(515151)
    db = R.RefDB()
    clouds = []
    for i, o in enumerate(objs):
        rc = R.RefCloud.from_levels({l: (o.pos(l), o.nor(l)) for l in range(5)})
        clouds.append(rc)
        db.add_object(rc, int(z[f"obj{i}_meta"][0]), int(z[f"obj{i}_meta"][1]))
    out = {}
    for i, (o, rc) in enumerate(zip(objs, clouds)):
        pose = z[f"obj{i}_pose"].reshape(4, 4).T
        props = proposals_for(pose, rng)
        out[f"nms{i}_props"] = props
        out[f"nms{i}_centroid"] = R.cloud_centroid(rc)
        out[f"nms{i}_overlap"] = np.array([R.overlap_factor(rc, props[0, :16], props[j, :16], 0.1, 1, 0) for j in range(len(props))], np.float32)
        out[f"nms{i}_overlap_boundary"] = np.array([R.overlap_factor(rc, props[0, :16], props[j, :16], 0.1, 0, 1) for j in range(16)], np.float32)
        out[f"nms{i}_kept"] = db.nms(i, props, 0.2)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "nms_golden.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
