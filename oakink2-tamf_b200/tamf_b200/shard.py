"""Batch sharding of independent reverse chains across ranks + the final gather.

Reference: launch/sample.py:198-199 (contiguous index range [len*w/W, len*(w+1)/W) per worker) and :272-289 (one
process per worker, no collective: workers write .npy independently).  Here one process per GPU (torchrun); the only
exchange on the path is one all_gather of the finished samples.  Works on any torch.distributed
backend (nccl on the B200s, gloo in the CPU tests)."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> range:
    """Contiguous shard of `n` sequences owned by `rank` (launch/sample.py:198-199)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world: {rank}/{world}")
    return range(n * rank // world, n * (rank + 1) // world)


def batches(r: range, batch_size: int):
    """Split a shard into per-step batches of at most `batch_size` sequences."""
    if batch_size <= 0:
        raise ValueError("batch_size must be positive")
    for s in range(r.start, r.stop, batch_size):
        yield range(s, min(s + batch_size, r.stop))


def gather_samples(local: torch.Tensor, n_total: int) -> torch.Tensor:
    """All ranks contribute their shard [n_local, ...] (ragged across ranks when world does not divide n_total);
    every rank returns the full [n_total, ...] tensor in sequence order."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        if local.shape[0] != n_total:
            raise ValueError("single process: local shard must be the whole set")
        return local
    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = [len(shard_range(n_total, r, world)) for r in range(world)]
    if local.shape[0] != sizes[rank]:
        raise ValueError(f"rank {rank}: shard has {local.shape[0]} rows, expected {sizes[rank]}")
    mx = max(sizes)
    if all(s == mx for s in sizes):
        out = local.new_empty((n_total,) + tuple(local.shape[1:]))
        dist.all_gather_into_tensor(out, local.contiguous())
        return out
    pad = local.new_zeros((mx,) + tuple(local.shape[1:]))
    pad[: local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad)
    return torch.cat([p[:s] for p, s in zip(parts, sizes)], 0)
