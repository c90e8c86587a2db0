"""one (or a few) pose_proposal steps of a named workload, for ncu captures:  python scripts/one_step.py [C2] [n_steps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rescan_b200 import api, pipeline  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "C2"
n_steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
api.set_device(0)
scene, rotations, translations = pipeline.make_workload(name)
models = pipeline.upload_objects(scene.objects)
for _ in range(n_steps):
    res = pipeline.run_step((scene.scan.pos(1), scene.scan.nor(1)), (scene.scan.pos(2), scene.scan.nor(2)), models, rotations,
                            translations, top_k=64)
print("evaluations", res.n_evaluations, "launches", api.launch_count())
