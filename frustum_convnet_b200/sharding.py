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

    Set-up (once): buffers are plain torch CUDA tensors; their CUDA-IPC handles travel through
    ``torch.distributed.all_gather_object`` (torch.multiprocessing.reductions.reduce_tensor), peers are opened with
    ``rebuild_cuda_tensor`` and peer access is enabled by a first device-to-device copy.  Raises if peer memory
    cannot be mapped (the caller may then fall back to ``all_gather_into_tensor``)."""

    def __init__(self, slots: int, block_numel: int, device: torch.device, group=None):
        import torch.distributed as dist
        from torch.multiprocessing.reductions import reduce_tensor
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.slots, self.n = int(slots), int(block_numel)
        self.device = device
        self.buf = torch.zeros((self.slots, self.world, self.n), dtype=torch.float32, device=device)
        self.flags = torch.zeros((self.slots, self.world), dtype=torch.int32, device=device)
        torch.cuda.synchronize(device)
        mine = (reduce_tensor(self.buf), reduce_tensor(self.flags))
        everyone = [None] * self.world
        dist.all_gather_object(everyone, mine, group=group)
        self.peer_buf, self.peer_flags = [], []
        for r, ((fb, ab), (ff, af)) in enumerate(everyone):
            if r == self.rank:
                self.peer_buf.append(self.buf)
                self.peer_flags.append(self.flags)
                continue
            pb, pf = fb(*ab), ff(*af)                       # tensors living on the exporter's device
            assert tuple(pb.shape) == tuple(self.buf.shape) and pb.dtype == torch.float32
            # enable peer access from OUR device to the peer's (what a kernel here needs to store there)
            if not torch.cuda.can_device_access_peer(device.index, pb.device.index):
                raise RuntimeError("no peer access from cuda:%d to cuda:%d" % (device.index, pb.device.index))
            probe = torch.zeros(1, dtype=torch.int32, device=device)
            pf[0, self.rank:self.rank + 1].copy_(probe)     # device-to-device copy: torch enables P2P on first use
            self.peer_buf.append(pb)
            self.peer_flags.append(pf)
        torch.cuda.synchronize(device)
        dist.barrier(group=group)

    def local_block(self, slot: int) -> torch.Tensor:
        """Where THIS rank's results of `slot` live in its own gather buffer (use as the plan's output block)."""
        return self.buf[slot, self.rank]

    def peer_targets(self, slot: int):
        """-> (list of peer blocks, list of flag addresses incl. the local one) for FrustumEngine plans."""
        blocks = [self.peer_buf[r][slot, self.rank] for r in range(self.world) if r != self.rank]
        flags = [self.peer_flags[r][slot, self.rank].data_ptr() for r in range(self.world)]
        return blocks, flags

    def gathered(self, slot: int) -> torch.Tensor:
        """(world, block) view of everything gathered on this rank for `slot`."""
        return self.buf[slot]
