"""Data-parallel plumbing for independent (face, shape, color) triples: static round-robin
sharding, one parameter broadcast at init, max-over-ranks timing.  No collective runs inside a step.
Works with NCCL (GPU) and gloo (CPU tests)."""
from __future__ import annotations

from typing import List, Sequence

import torch
import torch.distributed as dist


def shard_indices(n_items: int, rank: int, world: int) -> List[int]:
    """Rank r takes items r, r+world, ... (the reference loops triples serially, main.py:23-44)."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world {world}")
    return list(range(rank, n_items, world))


def broadcast_module_(module: torch.nn.Module, src: int = 0, bucket_bytes: int = 256 << 20) -> int:
    """Replicate parameters and buffers from `src` (weights are replicated, never sharded).  Tensors are packed by dtype
    into flat buckets of <= `bucket_bytes`, so a network is a handful of collectives instead of one per tensor (round 1
    issued thousands of tiny `dist.broadcast` calls for 2.45 GB).  Returns the number of bytes broadcast."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return 0
    tensors = [t.data for t in list(module.parameters()) + list(module.buffers()) if t.numel() > 0]
    nbytes = 0
    by_type = {}
    for t in tensors:
        by_type.setdefault((t.dtype, t.device), []).append(t)
    for (_dtype, _dev), group in by_type.items():
        bucket, size = [], 0
        for t in group + [None]:
            if t is not None and (not bucket or size + t.numel() * t.element_size() <= bucket_bytes):
                bucket.append(t)
                size += t.numel() * t.element_size()
                continue
            flat = torch.cat([b.reshape(-1) for b in bucket])
            dist.broadcast(flat, src=src)
            off = 0
            for b in bucket:
                b.copy_(flat[off:off + b.numel()].view_as(b))
                off += b.numel()
            nbytes += size
            bucket, size = ([t], t.numel() * t.element_size()) if t is not None else ([], 0)
    return nbytes


def max_over_ranks(value: float, device=None) -> float:
    """Multi-GPU numbers are reported as the max over ranks of the device-timed duration."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_counts(count: int, device=None) -> Sequence[int]:
    """Host-side gather of per-rank processed-unit counters (after the timed region)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [count]
    t = torch.tensor([count], dtype=torch.int64, device=device)
    out = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return [int(o.item()) for o in out]
