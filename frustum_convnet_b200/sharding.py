"""Batch sharding helpers for the one-process-per-GPU deployment.

Frustums are independent in eval mode, so the path shards by contiguous batch slices with
replicated weights and no data-path collective.  ``pack_outputs`` flattens the 6-tuple of
PointNetDet.forward (/root/reference/models/det_base.py:411) into one block so that a single
``all_gather_into_tensor`` (NCCL over NVLink) — or a single D2H copy — moves a rank's result.
"""
from __future__ import annotations

from typing import Sequence

import torch


def shard_slice(B: int, rank: int, world: int) -> slice:
    """Contiguous, balanced partition of B frustums over `world` ranks (first ranks get the remainder)."""
    base, rem = divmod(B, world)
    start = rank * base + min(rank, rem)
    return slice(start, start + base + (1 if rank < rem else 0))


def pack_outputs(outs: Sequence[torch.Tensor]) -> torch.Tensor:
    return torch.cat([o.reshape(-1).float() for o in outs])


def unpack_outputs(flat: torch.Tensor, B: int, T2: int, num_bins: int, num_size: int):
    widths = (2, 3, 1, 3, num_bins, num_size)
    outs, off = [], 0
    for i, w in enumerate(widths):
        n = B * T2 * w
        v = flat[off: off + n]
        outs.append(v.view(B, T2) if i == 2 else v.view(B, T2, w))
        off += n
    return tuple(outs)


class PeerResultExchange:
    """Result exchange of the inference path WITHOUT a collective ("a final logits
    all-gather in inference"): every rank owns one gather buffer of ``slots x world`` result blocks; rank r's
    forward in flight on slot k stores its decoded rows straight into block [k][r] of EVERY rank's buffer (its own
    included) from the heads epilogue of the persistent FCN kernel - plain stores to peer memory mapped over
    NVLink (CUDA IPC), no NCCL kernel competing with the persistent CTAs for SMs, no per-step Python collective.
    A per-(slot, rank) int32 epoch flag is raised in every peer when the forward is complete (system-scope
    release), which is what a downstream consumer polls.

    Set-up (once): ONE torch CUDA tensor per rank holds blocks and flags; it is exported with ``fcn_ipc_export``
    (CUDA-IPC handle of its allocation + offset), the handles travel through
    ``torch.distributed.all_gather_object`` and every peer maps them with ``fcn_ipc_open`` on ITS OWN device
    (lazy peer access).  Raises if peer memory cannot be mapped (the caller may then fall back to
    ``all_gather_into_tensor``)."""

    def __init__(self, slots: int, block_numel: int, device: torch.device, group=None):
        import ctypes as C

        import torch.distributed as dist

        from . import _lib
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.slots, self.n = int(slots), int(block_numel)
        self.device = device
        nblk = self.slots * self.world * self.n
        nflag = self.slots * self.world
        # a dedicated allocation > 1 MB: the caching allocator gives it its own cudaMalloc segment family; the
        # export covers whatever allocation it lives in anyway (handle of the base + offset)
        self.mem = torch.zeros(nblk + ((nflag + 3) // 4) * 4, dtype=torch.float32, device=device)
        self.buf = self.mem[:nblk].view(self.slots, self.world, self.n)
        self.flags = self.mem[nblk:nblk + nflag].view(torch.int32).view(self.slots, self.world)
        torch.cuda.synchronize(device)
        handle = (C.c_ubyte * 64)()
        off = C.c_longlong(0)
        with torch.cuda.device(device):
            _lib.call("fcn_ipc_export", self.mem.data_ptr(), C.addressof(handle), C.byref(off))
        everyone = [None] * self.world
        dist.all_gather_object(everyone, (bytes(handle), int(off.value)), group=group)
        self._opened = []
        self.peer_base = []                                  # device address of every rank's `mem` as seen from here
        for r, (h, o) in enumerate(everyone):
            if r == self.rank:
                self.peer_base.append(self.mem.data_ptr())
                continue
            hb = (C.c_ubyte * 64).from_buffer_copy(h)
            base = C.c_void_p()
            with torch.cuda.device(device):
                _lib.call("fcn_ipc_open", C.addressof(hb), C.byref(base))
            self._opened.append(base.value)
            self.peer_base.append(base.value + o)
        self._flag_off = 4 * nblk
        torch.cuda.synchronize(device)
        dist.barrier(group=group)

    def close(self):
        from . import _lib
        for b in self._opened:
            _lib.call("fcn_ipc_close", b)
        self._opened = []

    def local_block(self, slot: int) -> torch.Tensor:
        """Where THIS rank's results of `slot` live in its own gather buffer (use as the plan's output block)."""
        return self.buf[slot, self.rank]

    def peer_targets(self, slot: int):
        """-> (device addresses of this rank's block in every OTHER rank's buffer, flag addresses in ALL ranks)."""
        blk = lambda r: self.peer_base[r] + 4 * ((slot * self.world + self.rank) * self.n)
        flg = lambda r: self.peer_base[r] + self._flag_off + 4 * (slot * self.world + self.rank)
        return ([blk(r) for r in range(self.world) if r != self.rank], [flg(r) for r in range(self.world)])

    def gathered(self, slot: int) -> torch.Tensor:
        """(world, block) view of everything gathered on this rank for `slot`."""
        return self.buf[slot]
