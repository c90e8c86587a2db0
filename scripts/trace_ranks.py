"""host-side stage trace of one pose-sharded step on every rank (torchrun):  torchrun ... scripts/trace_ranks.py [C2|C3]
   TRACE_EXCHANGE=peer (default: one peer-mapped slot per object chain) | gloo | nccl;  TRACE_FULL=1 prints every event"""
import collections
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
dist.init_process_group("nccl", device_id=dev)
cfg = synth.CONFIGS[name]
scene = synth.make_scene(**cfg["scene"])
rot = posegrid.rotation_xforms(cfg["n_rot"])
trans = synth.translation_seeds(scene.scan, cfg["n_seeds"])  # one fixed problem, sharded (strong scaling)
models = pipeline.upload_objects(scene.objects)
args = ((scene.scan.pos(1), scene.scan.nor(1)), (scene.scan.pos(2), scene.scan.nor(2)), models, rot, trans)
mode = os.environ.get("TRACE_EXCHANGE", "peer")
hg = dist.new_group(backend="gloo") if mode == "gloo" else None
peer = None
if mode == "peer":
    from rescan_b200 import peerx
    peer = peerx.PeerExchange(dist, rank, world, local, n_slots=max(64, len(models)))
kw = dict(top_k=64, nms_dist=0.2, rank=rank, world=world, dist=dist, device=dev, host_group=hg, peer=peer)
for _ in range(2):
    pipeline.run_step(*args, **kw)
dist.barrier()
torch.cuda.synchronize()
api.profile_reset()
api.profile_enable(True)
tr = pipeline.run_step(*args, trace=True, **kw).trace
api.profile_enable(False)
prof = {n: round(api.profile_get(n)[0], 1) for n in ("score_dense", "dense_prefilter", "dense_search", "score", "icp", "overlap")}
for r in range(world):
    dist.barrier()
    if r == rank:
        print(f"--- rank {rank}  kernels (overlapping lanes, ms) {prof}", flush=True)
        by = collections.defaultdict(list)
        for k, stage, a, b in tr:
            by[stage].append((a, b))
            if os.environ.get("TRACE_FULL") == "1":
                print(f"  obj {k:2d} {stage:9s} {a * 1e3:7.2f} -> {b * 1e3:7.2f} ms  ({(b - a) * 1e3:6.2f})", flush=True)
        for stage, v in by.items():
            d = [b - a for a, b in v]
            print(f"  {stage:10s} n {len(v):3d}  sum {sum(d) * 1e3:8.1f}  mean {sum(d) / len(d) * 1e3:7.2f}  max {max(d) * 1e3:7.2f}  "
                  f"first start {min(a for a, b in v) * 1e3:7.1f}  last end {max(b for a, b in v) * 1e3:7.1f}", flush=True)
if peer is not None:
    peer.close()
dist.destroy_process_group()
