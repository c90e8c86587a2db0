"""dense level-4 launches of a named workload's objects, one at a time (ncu captures / A-B runs of the dense search):
   python scripts/dense_one.py [C2] [object index | all] [reps]   -> ms per launch by phase, digest of the scores' proposals
   RSGPU_DENSE_IMPL=warp selects the first design; RSGPU_DENSE_STATS=1 prints the queue / item census per chunk"""
import hashlib
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rescan_b200 import api, pipeline, posegrid  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "C2"
which = sys.argv[2] if len(sys.argv) > 2 else "all"
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
api.set_device(0)
scene, rotations, translations = pipeline.make_workload(name)
dyn = [o for o in scene.objects if not o.is_static]
if which != "all":
    dyn = [dyn[int(which)]]
g1 = api.HashGrid(scene.scan.pos(1), 0.05, normals=scene.scan.nor(1))
walk = posegrid.spatial_order(translations)
trans = np.ascontiguousarray(translations[walk])
names = ("score_dense", "dense_prefilter", "dense_bin", "dense_search", "dense_reduce", "score")
h = hashlib.sha1()
tot = {n: 0.0 for n in names}
for o in dyn:
    lv = {l: api.PointCloud(o.cloud.pos(l), o.cloud.nor(l)) for l in (4, 3, 2)}
    for rep in range(reps):
        api.profile_reset()
        api.profile_enable(True)
        props, ids = api.propose_poses(lv[4], lv[3], lv[2], g1, rotations, trans, top_k=64, translation_ids=walk)
        api.profile_enable(False)
    prof = {n: round(api.profile_get(n)[0], 3) for n in names}
    for n in names:
        tot[n] += prof[n]
    h.update(props.tobytes()); h.update(ids.tobytes())
    print(f"object {o.uidx}: {len(lv[4])} level-4 points, {len(props)} proposals, ms {prof}")
print(f"total ms {dict((k, round(v, 2)) for k, v in tot.items())} digest {h.hexdigest()[:12]}")
