"""a few pose_proposal steps of a named workload (ncu captures, A/B runs of kernel variants via RSGPU_* env vars):
   python scripts/one_step.py [C2] [n_steps]   -> per-kernel ms/step and a digest of the proposals
   STEP_LANES=n runs n objects at once on their own lanes (bench.py's default is 4); STEP_NMS=1 adds the two NMS passes of main.cpp:161/205 to the step (bench.py's default)"""
import hashlib
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rescan_b200 import api, pipeline  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "C2"
n_steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
api.set_device(0)
LANES = int(os.environ.get("STEP_LANES", "1"))
NMS = 0.2 if os.environ.get("STEP_NMS", "0") == "1" else None
scene, rotations, translations = pipeline.make_workload(name)
models = pipeline.upload_objects(scene.objects)
args = ((scene.scan.pos(1), scene.scan.nor(1)), (scene.scan.pos(2), scene.scan.nor(2)), models, rotations, translations)
if n_steps > 1:
    pipeline.run_step(*args, top_k=64, nms_dist=NMS, lanes=LANES)  # warm-up
api.profile_reset()
api.profile_enable(os.environ.get("STEP_PROFILE", "1") != "0")
t0 = time.perf_counter()
for _ in range(n_steps):
    res = pipeline.run_step(*args, top_k=64, nms_dist=NMS, lanes=LANES)
dt = (time.perf_counter() - t0) / n_steps
api.profile_enable(False)
h = hashlib.sha1()
for p, i in zip(res.proposals, res.pose_ids):
    h.update(np.ascontiguousarray(p).tobytes())
    h.update(np.ascontiguousarray(i).tobytes())
prof = {n: round(api.profile_get(n)[0] / n_steps, 3) for n in ("grid_build", "score_dense", "score", "icp", "icp_search", "icp_solve", "overlap")}
env = {k: v for k, v in os.environ.items() if k.startswith("RSGPU_")}
print(f"env {env} wall {dt * 1e3:.2f} ms/step kernels {prof} evaluations {res.n_evaluations} launches {api.launch_count()} digest {h.hexdigest()[:12]}")
if os.environ.get("STEP_TRACE") == "1":
    tr = pipeline.run_step(*args, top_k=64, nms_dist=NMS, lanes=LANES, trace=True).trace
    for k, stage, a, b in tr:
        print(f"  obj {k:2d} {stage:7s} {a * 1e3:7.2f} -> {b * 1e3:7.2f} ms  ({(b - a) * 1e3:6.2f})")
