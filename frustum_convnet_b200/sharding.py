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
