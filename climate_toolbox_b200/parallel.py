"""Multi-GPU: the path shards along time (every day is independent given the plan).

One process per GPU; each rank builds the (deterministic, ~5 MB) plan itself and
aggregates a contiguous block of days.  There is NO data-path collective; the only
exchange is an optional final gather of the region x time outputs (NCCL all_gather over
NVLink on GPUs, gloo in the CPU tests).  SURVEY.md section 8(e).
"""
from __future__ import annotations

import torch
import torch.distributed as dist

__all__ = ["shard_range", "shard_sizes", "all_gather_time"]


def shard_sizes(T, world_size, align=32):
    """Days per rank: contiguous blocks, multiples of ``align`` (the kernel's time tile)
    except the last non-empty one; earlier ranks take the extra tiles."""
    tiles = (T + align - 1) // align
    base, extra = divmod(tiles, world_size)
    sizes, left = [], T
    for r in range(world_size):
        n = min(left, (base + (1 if r < extra else 0)) * align)
        sizes.append(n)
        left -= n
    return sizes


def shard_range(T, world_size, rank, align=32):
    sizes = shard_sizes(T, world_size, align)
    t0 = sum(sizes[:rank])
    return t0, t0 + sizes[rank]


def all_gather_time(local, T, group=None, align=32):
    """Gather per-rank ``[n_out, R, T_local]`` blocks into ``[n_out, R, T]`` on every rank.
    Ranks may hold different T_local (ragged last shard): blocks are padded to the largest
    shard for the collective and trimmed after."""
    world = dist.get_world_size(group)
    sizes = shard_sizes(T, world, align)
    tmax = max(sizes)
    n_out, R = local.shape[0], local.shape[1]
    pad = torch.zeros((n_out, R, tmax), dtype=local.dtype, device=local.device)
    pad[:, :, : local.shape[2]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([b[:, :, :n] for b, n in zip(bufs, sizes)], dim=2)
