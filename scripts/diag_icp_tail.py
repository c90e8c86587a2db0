"""distribution of ICP work in one C2 step: per starting pose the object size and the iteration count (explains the launch's tail)"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rescan_b200 import api, pipeline
name = sys.argv[1] if len(sys.argv) > 1 else "C2"
api.set_device(0)
scene, rotations, translations = pipeline.make_workload(name)
models = pipeline.upload_objects(scene.objects)
res = pipeline.run_step((scene.scan.pos(1), scene.scan.nor(1)), (scene.scan.pos(2), scene.scan.nor(2)), models, rotations, translations, top_k=64, do_icp=False)
g2 = api.HashGrid(scene.scan.pos(2), 0.05, normals=scene.scan.nor(2))
dyn = [m for m in models if not m.is_static]
jobs = [(m, p[p[:, 16] > 0, :16]) for m, p in zip(dyn, res.proposals) if (p[:, 16] > 0).any()]
for rep in range(2):
    api.profile_reset(); api.profile_enable(True)
    out = api.icp_align_multi([m.levels[2] for m, _ in jobs], g2, [t.copy() for _, t in jobs], 0.10, np.float32(np.deg2rad(60.0)))
    api.profile_enable(False)
    print("icp launch ms", api.profile_get("icp"))
tot = 0
for (m, t), (T, err, it) in zip(jobs, out):
    n = len(m.levels[2]); tot += n * it.sum()
    print(f"object n2={n:6d} poses={len(t):3d} iters: min {it.min()} median {int(np.median(it))} max {it.max()}  n*iters max {n * it.max():9d} sum {n * it.sum():10d}")
print("total point-iterations", tot)
