"""per-step wall times of the pose step with the scan resident in HBM vs uploaded from host buffers, alternated (why do the two
differ?):  python scripts/ab_steps.py [C3] [rounds]"""
import os
import sys
import time

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rescan_b200 import api, pipeline  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "C3"
rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 2
torch.cuda.set_device(0)
api.set_device(0)
scene, rot, trans = pipeline.make_workload(name)
models = pipeline.upload_objects(scene.objects)
host = {k: torch.from_numpy(v).pin_memory().numpy() for k, v in dict(p1=scene.scan.pos(1), n1=scene.scan.nor(1), p2=scene.scan.pos(2), n2=scene.scan.nor(2)).items()}
dev_t = {k: torch.from_numpy(v).cuda() for k, v in host.items()}
dev = {k: v.data_ptr() for k, v in dev_t.items()}
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def step(resident, do_flush=True):
    if do_flush:
        flush.zero_()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    pipeline.run_step((host["p1"], host["n1"]), (host["p2"], host["n2"]), models, rot, trans, top_k=64, nms_dist=0.2,
                      scan_dev=dev if resident else None)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) * 1e3


for _ in range(2):
    step(True)
for r in range(rounds):
    for mode, res, fl in (("resident", True, True), ("host", False, True), ("resident/noflush", True, False), ("host/noflush", False, False)):
        print(f"round {r} {mode:18s} " + " ".join(f"{step(res, fl):8.1f}" for _ in range(3)), flush=True)
