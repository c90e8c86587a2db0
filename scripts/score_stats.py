"""work census of the dense scoring kernel (needs a library built with EXTRA=-DRS_SCORE_STATS; prints to stderr):
   make -f rescan_b200/csrc/Makefile EXTRA=-DRS_SCORE_STATS && python scripts/score_stats.py [C2]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rescan_b200 import api, pipeline  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "C2"
api.set_device(0)
scene, rotations, translations = pipeline.make_workload(name)
models = pipeline.upload_objects(scene.objects)
g1 = api.HashGrid(scene.scan.pos(1), 0.05, normals=scene.scan.nor(1))
for m in models:
    if not m.is_static:
        print(f"object {m.uidx}: level-4 points {len(m.levels[4])}", file=sys.stderr)
        api.propose_poses(m.levels[4], m.levels[3], m.levels[2], g1, rotations, translations, top_k=64)
