"""All-gather of the pose-sharded step's small lists over NVLink, one process per GPU (include/rsgpu.h rsgpu_peer_*,
rescan_b200/csrc/peer.cu): every rank's receive area in HBM is mapped by all peers through CUDA IPC, the payloads and
sequence flags are written by copy engines, one warp waits.  What is exchanged is the survivor list of mgs_propose_poses
(reference apps/pose_proposal/pose_proposal.cpp:348-359) and the refined rows of apps/pose_proposal/main.cpp:195-201.

The area is divided into SLOTS, one per object chain: exchanges on different slots share nothing, so every object's chain
(search -> top-k exchange -> NMS -> refinement of this rank's share -> row exchange -> NMS) runs on its own lane thread and
meets its peers whenever they get there - no global order of collectives (rescan_b200/pipeline.py).

torch.distributed is used ONCE, at construction, to hand the 64-byte IPC handles round (all_gather_object) and for the
barrier before the areas are freed; the data path never touches it."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import api


class PeerExchange:
    def __init__(self, dist, rank, world, local_rank=None, n_slots=64, slot_bytes=256 << 10, timeout_s=30.0):
        self.dist, self.rank, self.world, self.timeout_s = dist, int(rank), int(world), float(timeout_s)
        self.n_slots, self.slot_bytes = int(n_slots), int(slot_bytes)
        self._seq = [0] * self.n_slots  # uses of each slot so far; a slot is driven by one thread at a time
        L = api.lib()
        hb = L.rsgpu_peer_handle_bytes()
        mine = (C.c_ubyte * hb)()
        api._check(L.rsgpu_peer_init(self.rank, self.world, self.n_slots, self.slot_bytes, C.cast(mine, C.c_void_p)))
        self._live = True
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(mine))  # set-up only
        blob = b"".join(handles)
        assert len(blob) == hb * self.world
        buf = (C.c_ubyte * len(blob)).from_buffer_copy(blob)
        api._check(L.rsgpu_peer_open(C.cast(buf, C.c_void_p)))
        dist.barrier()

    def allgather(self, buf, slot=0):
        """equally sized byte buffers -> uint8 [world, nbytes] on `slot`; the uses of ONE slot are ordered identically on every
        rank, different slots are independent (any thread, any order).  Payloads above the slot size go in pieces."""
        send = np.ascontiguousarray(buf).view(np.uint8).reshape(-1)
        slot = int(slot) % self.n_slots
        recv = np.empty((self.world, send.nbytes), np.uint8)
        L = api.lib()
        for lo in range(0, max(send.nbytes, 1), self.slot_bytes):
            piece = send[lo: lo + self.slot_bytes]
            got = np.empty((self.world, piece.nbytes), np.uint8)
            self._seq[slot] += 1
            api._check(L.rsgpu_peer_allgather(slot, self._seq[slot], api._ptr(piece), piece.nbytes, api._ptr(got), self.timeout_s))
            recv[:, lo: lo + piece.nbytes] = got
        return recv

    # ---- rooted patterns: the two halves of one use of a slot (gather to an owner, broadcast from it)
    def begin_use(self, slot):
        """number of the next use of `slot`; every rank numbers a slot's uses identically, whatever its role in them"""
        slot = int(slot) % self.n_slots
        self._seq[slot] += 1
        return self._seq[slot]

    @staticmethod
    def _mask(ranks):
        m = 0
        for r in ranks:
            m |= 1 << int(r)
        return m

    def put(self, slot, seq, dst_ranks, buf):
        """this rank's payload of use `seq` into its row of `slot` on the ranks `dst_ranks`; never waits for a peer"""
        send = np.ascontiguousarray(buf).view(np.uint8).reshape(-1)
        if send.nbytes > self.slot_bytes:
            raise ValueError(f"peer exchange: payload of {send.nbytes} bytes exceeds the slot size {self.slot_bytes} (PeerExchange(slot_bytes=...))")
        api._check(api.lib().rsgpu_peer_put(int(slot) % self.n_slots, int(seq), self._mask(dst_ranks), api._ptr(send), send.nbytes))

    def get(self, slot, seq, src_ranks):
        """waits for the payloads the ranks `src_ranks` put for use `seq` of `slot` -> {rank: uint8 array}"""
        rows = np.empty((self.world, self.slot_bytes), np.uint8)
        lens = np.zeros(self.world, np.int64)
        api._check(api.lib().rsgpu_peer_get(int(slot) % self.n_slots, int(seq), self._mask(src_ranks), api._ptr(rows), self.slot_bytes,
                                            api._ptr(lens), self.timeout_s))
        return {int(r): rows[int(r), : int(lens[int(r)])].copy() for r in src_ranks}

    def slot(self, slot):
        """the exchange interface of pipeline._allgather_bytes bound to one slot"""
        return _Slot(self, slot)

    def close(self):
        if getattr(self, "_live", False):
            self._live = False
            try:
                self.dist.barrier()  # nobody may still be writing into an area that is about to be freed
            finally:
                api.lib().rsgpu_peer_close()


class _Slot:
    def __init__(self, peer, slot):
        self.peer, self.index = peer, int(slot)

    def allgather(self, buf):
        return self.peer.allgather(buf, self.index)
