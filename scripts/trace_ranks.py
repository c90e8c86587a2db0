"""host-side stage trace of one pose-sharded step on every rank (torchrun):  torchrun ... scripts/trace_ranks.py [C2]"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rescan_b200 import api, pipeline, posegrid, synth  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "C2"
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
api.set_device(local)
dev = torch.device("cuda", local)
if os.environ.get("TRACE_HP") == "1":
    os.environ["TORCH_NCCL_HIGH_PRIORITY"] = "1"
dist.init_process_group("nccl", device_id=dev)
cfg = synth.CONFIGS[name]
scene = synth.make_scene(**cfg["scene"])
rot = posegrid.rotation_xforms(cfg["n_rot"])
trans = synth.translation_seeds(scene.scan, cfg["n_seeds"] * world)
models = pipeline.upload_objects(scene.objects)
args = ((scene.scan.pos(1), scene.scan.nor(1)), (scene.scan.pos(2), scene.scan.nor(2)), models, rot, trans)
hg = dist.new_group(backend="gloo") if os.environ.get("TRACE_GLOO") == "1" else None
kw = dict(top_k=64, nms_dist=0.2, rank=rank, world=world, dist=dist, device=dev, host_group=hg)
for _ in range(3):
    pipeline.run_step(*args, **kw)
dist.barrier()
torch.cuda.synchronize()
tr = pipeline.run_step(*args, trace=True, **kw).trace
for r in range(world):
    dist.barrier()
    if r == rank:
        print(f"--- rank {rank}", flush=True)
        for k, stage, a, b in tr:
            print(f"  obj {k:2d} {stage:9s} {a * 1e3:7.2f} -> {b * 1e3:7.2f} ms  ({(b - a) * 1e3:6.2f})", flush=True)
dist.destroy_process_group()
