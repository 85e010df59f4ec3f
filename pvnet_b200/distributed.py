"""Batch sharding across the GPUs of one box.

The hot path is embarrassingly parallel over images (the reference loops `for bi in
range(b)` with no cross-image state, ransac_voting_gpu.py:525; DataParallel splits dim 0).
One process per GPU owns a contiguous shard of the batch; the only exchange is an
all_gather of the pose inputs -- keypoints [b/N,K,2] (+ covariances [b/N,K,2,2]) --
about 216 B per image at K=9.  Works over NCCL (GPU) and gloo (CPU tests).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int):
    """Contiguous, balanced [lo, hi) of `n_items` for `rank`; shard sizes differ by <= 1."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def image_seed(base_seed: int, image_index: int) -> int:
    """Per-image RNG seed independent of how the batch is sharded (1/2/4/8 GPUs agree)."""
    return (base_seed * 1000003 + image_index * 7919) & 0x7FFFFFFF


def gather_results(local: torch.Tensor, n_items: int | None = None) -> torch.Tensor:
    """all_gather of per-rank result rows (dim 0 = images) in rank order.  Shards may be
    ragged (see shard_range); rows are padded to the largest shard for the collective."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    if n_items is None:
        sizes = [torch.zeros(1, dtype=torch.int64, device=local.device) for _ in range(world)]
        dist.all_gather(sizes, torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device))
        counts = [int(s.item()) for s in sizes]
    else:
        counts = [shard_range(n_items, r, world)[1] - shard_range(n_items, r, world)[0] for r in range(world)]
    if min(counts) == max(counts) == local.shape[0] and hasattr(dist, "all_gather_into_tensor"):
        # equal shards (the usual case): one collective into one preallocated tensor, no list / cat
        out = torch.empty((world * counts[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous())
        return out
    width = max(counts)
    padded = local
    if local.shape[0] < width:
        pad = torch.zeros((width - local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        padded = torch.cat([local, pad], 0)
    bufs = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(bufs, padded.contiguous())
    return torch.cat([b[:c] for b, c in zip(bufs, counts)], 0)
