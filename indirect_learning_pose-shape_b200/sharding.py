"""Batch sharding across one process per GPU.

Samples are independent end to end (every op of the path is per-sample; the reference itself only ever replicated
towers over batch slices, train.py:205-210), so the decode path needs NO collective: rank r owns a contiguous slice of
the global batch.  The only collective is an all-gather that assembles outputs for validation, outside any timed
region.  Works with any torch.distributed backend (NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.distributed as dist


def shard_bounds(global_batch: int, rank: int, world_size: int) -> Tuple[int, int]:
    """[lo, hi) of rank's contiguous slice; the first `global_batch % world_size` ranks get one extra sample."""
    if not (0 <= rank < world_size):
        raise ValueError("rank %d outside world of %d" % (rank, world_size))
    base, rem = divmod(int(global_batch), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_slice(global_tensor: torch.Tensor, rank: int, world_size: int) -> torch.Tensor:
    lo, hi = shard_bounds(global_tensor.shape[0], rank, world_size)
    return global_tensor[lo:hi]


def all_gather_outputs(local: torch.Tensor, global_batch: int, group=None) -> torch.Tensor:
    """Assemble the (global_batch, ...) tensor from every rank's slice (validation only; ragged slices allowed)."""
    if not dist.is_available() or not dist.is_initialized():
        return local
    world = dist.get_world_size(group)
    sizes = [shard_bounds(global_batch, r, world)[1] - shard_bounds(global_batch, r, world)[0] for r in range(world)]
    pad = max(sizes)
    buf = local.new_zeros((pad,) + tuple(local.shape[1:]))
    buf[: local.shape[0]] = local
    parts: List[torch.Tensor] = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf.contiguous(), group=group)
    return torch.cat([p[:s] for p, s in zip(parts, sizes)], dim=0)
