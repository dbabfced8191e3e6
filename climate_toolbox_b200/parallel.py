"""Multi-GPU: the path shards along time (every day is independent given the plan).

One process per GPU; each rank builds the (deterministic, ~5 MB) plan itself and
aggregates a contiguous block of days.  There is NO data-path collective; the only
exchange is an optional final gather of the region x time outputs (NCCL all_gather over
NVLink on GPUs, gloo in the CPU tests).  SURVEY.md section 8(e).
"""
from __future__ import annotations

import torch
import torch.distributed as dist

__all__ = ["shard_range", "shard_sizes", "all_gather_time", "aggregate_time_sharded"]


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


def aggregate_time_sharded(ds, variable, aggwt, agglev, weights, backup_aggwt="areawt", gather=True,
                           group=None, time_dim="time", aggregate_fn=None, **engine_opts):
    """``weighted_aggregate_grid_to_regions`` over a process group: every rank aggregates its own
    contiguous block of days of ``ds`` (each rank holds, or lazily views, the same Dataset) on its
    own GPU with the replicated plan; with ``gather=True`` the region x time blocks are exchanged
    once at the end (one all_gather) and every rank returns the full result, otherwise each rank
    returns its block (what a job that writes per-rank files wants: 285 MB per rank and year-block
    would otherwise cross NVLink for nothing).

    ``aggregate_fn`` (tests) replaces the single-GPU aggregation; it must have the signature of
    ``weighted_aggregate_grid_to_regions``.
    """
    import numpy as np

    from ._xr import Dataset, Variable, from_any
    if aggregate_fn is None:
        from .aggregations.aggregations import weighted_aggregate_grid_to_regions as aggregate_fn
    ds = from_any(ds)
    if not (dist.is_available() and dist.is_initialized()):
        return aggregate_fn(ds, variable, aggwt, agglev, weights=weights, backup_aggwt=backup_aggwt,
                            **engine_opts)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    T = ds.dims[time_dim]
    t0, t1 = shard_range(T, world, rank)
    local = aggregate_fn(ds.isel(**{time_dim: np.arange(t0, t1)}), variable, aggwt, agglev,
                         weights=weights, backup_aggwt=backup_aggwt, **engine_opts)
    local = from_any(local)
    if not gather:
        return local
    names = [variable] if isinstance(variable, str) else list(variable)
    backend = dist.get_backend(group)
    out = Dataset()
    for name in names:
        v = local._vars[name]
        ax_t, ax_r = v.dims.index(time_dim), v.dims.index(agglev)
        a = v.physical if isinstance(v.physical, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(v.values))
        if backend == "nccl" and not a.is_cuda:
            a = a.cuda()
        blk = a.permute(ax_r, ax_t).contiguous()[None]                    # [1, R, T_local]
        full = all_gather_time(blk, T, group)[0]                          # [R, T]
        full = full if (ax_r, ax_t) == (0, 1) else full.t()
        out[name] = Variable(v.dims, full if full.is_cuda else full.numpy(), v.attrs)
    out._coords[agglev] = local._coords[agglev]
    for d, c in ds._coords.items():
        if d not in ("lat", "lon") and "lat" not in c.dims and "lon" not in c.dims:
            out._coords[d] = Variable(c.dims, c.values, c.attrs)
    return out
