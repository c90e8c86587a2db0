"""All-gather of the pose-sharded step's small lists over NVLink, one process per GPU (include/rsgpu.h rsgpu_peer_*,
rescan_b200/csrc/peer.cu): every rank's receive area in HBM is mapped by all peers through CUDA IPC, the payloads and round
flags are written by copy engines, one warp waits.  What is exchanged is the survivor list of mgs_propose_poses (reference
apps/pose_proposal/pose_proposal.cpp:348-359) and the refined rows of apps/pose_proposal/main.cpp:195-201.

torch.distributed is used ONCE, at construction, to hand the 64-byte IPC handles round (all_gather_object) and for the
barrier before the areas are freed; the data path never touches it."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import api


class PeerExchange:
    def __init__(self, dist, rank, world, local_rank=None, slot_bytes=4 << 20, timeout_s=30.0):
        self.dist, self.rank, self.world, self.timeout_s = dist, int(rank), int(world), float(timeout_s)
        self.slot_bytes = int(slot_bytes)
        L = api.lib()
        hb = L.rsgpu_peer_handle_bytes()
        mine = (C.c_ubyte * hb)()
        api._check(L.rsgpu_peer_init(self.rank, self.world, self.slot_bytes, C.cast(mine, C.c_void_p)))
        self._live = True
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(mine))  # set-up only
        blob = b"".join(handles)
        assert len(blob) == hb * self.world
        buf = (C.c_ubyte * len(blob)).from_buffer_copy(blob)
        api._check(L.rsgpu_peer_open(C.cast(buf, C.c_void_p)))
        dist.barrier()

    def allgather(self, buf):
        """equally sized byte buffers -> uint8 [world, nbytes]; every rank calls this in the same order"""
        send = np.ascontiguousarray(buf).view(np.uint8).reshape(-1)
        if send.nbytes > self.slot_bytes:
            raise ValueError(f"peer exchange: payload of {send.nbytes} bytes exceeds the slot size {self.slot_bytes}")
        recv = np.empty((self.world, send.nbytes), np.uint8)
        api._check(api.lib().rsgpu_peer_allgather(api._ptr(send), send.nbytes, api._ptr(recv), self.timeout_s))
        return recv

    def close(self):
        if getattr(self, "_live", False):
            self._live = False
            try:
                self.dist.barrier()  # nobody may still be writing into an area that is about to be freed
            finally:
                api.lib().rsgpu_peer_close()
