"""NVLink peer-memory communication for the X-slab sharded mode: one IPC-shared pool per rank, bump-allocated in the same
sequence on every rank (so a tensor, a flag pair or a slot area sits at the same offset everywhere), and the two collectives
of the voxel stack -- halo exchange with the two neighbours, all-reduce of the GroupNorm sums -- as stream-ordered kernels of
libstereoscene_b200.so that store straight into the peers' pools (csrc/peer.cu).  No NCCL launch per collective, no host
synchronisation, capturable in a CUDA graph.  torch.distributed is only used once, to exchange the IPC handles."""
from __future__ import annotations

import ctypes as C
import math
from typing import Sequence

import torch
import torch.distributed as dist

from . import cabi


class _RawMem:
    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


class PeerPool:
    CONTROL_BYTES = 1 << 20          # epoch word, flag words, tickets
    ALIGN = 256

    def __init__(self, data_bytes: int, world: int, rank: int, device: torch.device, group=None, slot_bytes: int = 4 << 20):
        self.lib = cabi.load()
        self.world, self.rank, self.device = world, rank, device
        self.slot_bytes = slot_bytes
        self.nbytes = self.CONTROL_BYTES + slot_bytes + data_bytes
        base = C.c_void_p()
        cabi.check(self.lib.ss_peer_pool_alloc(self.nbytes, C.byref(base)), "ss_peer_pool_alloc")
        self.base = int(base.value)
        handle = C.create_string_buffer(64)
        cabi.check(self.lib.ss_peer_ipc_export(self.base, handle), "ss_peer_ipc_export")
        mine = (bytes(handle.raw), int(device.index))
        everyone = [None] * world
        if world > 1:
            dist.all_gather_object(everyone, mine, group=group)
        else:
            everyone = [mine]
        self.peer_base = []
        self._opened = []
        for r, (h, dev_r) in enumerate(everyone):
            if r == rank:
                self.peer_base.append(self.base)
                continue
            p = C.c_void_p()
            hb = C.create_string_buffer(h, 64)
            cabi.check(self.lib.ss_peer_ipc_open(hb, dev_r, C.byref(p)), "ss_peer_ipc_open")
            self.peer_base.append(int(p.value))
            self._opened.append(int(p.value))
        self.mem = torch.as_tensor(_RawMem(self.base, self.nbytes), device=device)        # uint8 view of my pool
        self.epoch_ptr = self.base                                                      # int at offset 0
        self.begin_host()
        if world > 1:
            dist.barrier(group=group)          # every pool is mapped everywhere before anybody pushes

    # ---- per-forward bump allocation (identical sequence on every rank) ------------------------------------------------
    def begin_host(self):
        self.ctl_off = 256
        self.slot_off = self.CONTROL_BYTES
        self.data_off = self.CONTROL_BYTES + self.slot_bytes

    def begin(self):
        """Start of a forward: rewind the bump allocators and advance the device epoch (stream-ordered)."""
        self.begin_host()
        cabi.check(self.lib.ss_peer_epoch_bump(self.epoch_ptr, torch.cuda.current_stream(self.device).cuda_stream), "ss_peer_epoch_bump")

    def _ctl(self, nbytes: int) -> int:
        off = self.ctl_off
        self.ctl_off += (nbytes + 15) // 16 * 16
        if self.ctl_off > self.CONTROL_BYTES:
            raise RuntimeError("peer pool: control region exhausted")
        return off

    def tensor(self, shape: Sequence[int], dtype=torch.float32):
        """(tensor viewing my pool, its pool offset)."""
        n = math.prod(shape) * torch.empty((), dtype=dtype).element_size()
        off = (self.data_off + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        if off + n > self.nbytes:
            raise RuntimeError(f"peer pool exhausted: need {off + n} of {self.nbytes} bytes")
        self.data_off = off + n
        return self.mem[off: off + n].view(dtype).view(*shape), off

    # ---- collectives ----------------------------------------------------------------------------------------------------
    def halo_push(self, buf: torch.Tensor, off: int, n: int, replicate: bool):
        """buf [1, n + 2, ...] at pool offset `off` on every rank."""
        plane = buf[0, 0].numel()
        flags = self._ctl(16)
        ticket = self._ctl(4)
        lo, hi = self.rank - 1, self.rank + 1
        plo = self.peer_base[lo] if lo >= 0 else None
        phi = self.peer_base[hi] if hi < self.world else None
        rc = self.lib.ss_peer_halo_push(
            buf.data_ptr(), (plo + off) if plo else None, (phi + off) if phi else None, plane, n, 1 if replicate else 0,
            self.base + flags, (plo + flags) if plo else None, (phi + flags) if phi else None, self.base + ticket, self.epoch_ptr,
            torch.cuda.current_stream(self.device).cuda_stream)
        cabi.check(rc, "ss_peer_halo_push")

    def allreduce(self, stats: torch.Tensor):
        """double tensor (contiguous) <- sum over ranks, in place."""
        n = stats.numel()
        if stats.dtype != torch.float64 or not stats.is_contiguous():
            raise RuntimeError("peer all-reduce: contiguous float64 sums expected")
        slot = self.slot_off
        self.slot_off += (self.world * n * 8 + 255) // 256 * 256
        if self.slot_off > self.CONTROL_BYTES + self.slot_bytes:
            raise RuntimeError("peer pool: slot region exhausted")
        flags = self._ctl(4 * self.world)
        slots = (C.c_void_p * self.world)(*[b + slot for b in self.peer_base])
        flgs = (C.c_void_p * self.world)(*[b + flags for b in self.peer_base])
        rc = self.lib.ss_peer_stats_allreduce(stats.data_ptr(), n, self.world, self.rank, slots, flgs, self.epoch_ptr,
                                              torch.cuda.current_stream(self.device).cuda_stream)
        cabi.check(rc, "ss_peer_stats_allreduce")

    def close(self):
        for p in self._opened:
            self.lib.ss_peer_ipc_close(p)
        self._opened = []
        if self.base:
            self.mem = None
            self.lib.ss_peer_pool_free(self.base)
            self.base = 0
