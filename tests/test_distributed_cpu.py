"""world_size-2 gloo test of the N>1 host logic (time sharding + output gather).
The per-shard compute is stood in by the oracle: what is under test is the partition
and the reassembly, which are identical on NCCL."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from climate_toolbox_b200.parallel import shard_range, shard_sizes


def test_shard_sizes_cover_and_align():
    for T in (0, 1, 31, 32, 33, 365, 1460, 1461):
        for w in (1, 2, 3, 4, 8):
            s = shard_sizes(T, w)
            assert sum(s) == T and len(s) == w
            assert all(n % 32 == 0 for n in s[:-1] if n and sum(s[: s.index(n) + 1]) < T)
            ranges = [shard_range(T, w, r) for r in range(w)]
            assert ranges[0][0] == 0 and ranges[-1][1] == T
            assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
            if T >= 32 * w:
                assert max(s) - min(s) <= 32 + 31


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, T, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import oracle
        from climate_toolbox_b200 import synthetic
        from climate_toolbox_b200.parallel import all_gather_time, shard_range

        lat, lon = synthetic.grid_labels(4.0)
        df = synthetic.weights_table(4.0, 60, seed=2)
        tas, _, _ = synthetic.tas_field(T, len(lat), len(lon), seed=1, dtype=np.float64)
        t0, t1 = shard_range(T, world, rank)
        loc = oracle.weighted_aggregate_grid_to_regions(
            tas[t0:t1], ("time", "lat", "lon"), lat, lon, df, "popwt", "hierid")[0]   # (t, R)
        local = torch.from_numpy(np.ascontiguousarray(loc.T))[None]                 # [1, R, t]
        full = all_gather_time(local, T)
        ref = oracle.weighted_aggregate_grid_to_regions(
            tas, ("time", "lat", "lon"), lat, lon, df, "popwt", "hierid")[0]
        ok = np.array_equal(full[0].numpy().T, ref, equal_nan=True)
        q.put((rank, bool(ok), tuple(full.shape)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("T", [70, 33])
def test_time_sharded_gather_world2(T):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, T, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert sorted(r[0] for r in res) == [0, 1]
    assert all(r[1] for r in res), res


def _oracle_aggregate(ds, variable, aggwt, agglev, weights=None, backup_aggwt="areawt", **kw):
    """Stand-in for the single-GPU aggregation (CPU tests have no GPU): the oracle behind the
    public signature, returning this package's Dataset."""
    import oracle
    from climate_toolbox_b200 import Dataset
    x = ds[variable].values
    out, rd, labels = oracle.weighted_aggregate_grid_to_regions(
        x, ds[variable].dims, ds["lat"].values, ds["lon"].values, weights, aggwt, agglev, backup_aggwt)
    return Dataset({variable: (rd, out)}, coords={"time": ds["time"].values, agglev: labels})


def _worker_api(rank, world, port, T, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import oracle
        from climate_toolbox_b200 import Dataset, synthetic
        from climate_toolbox_b200.parallel import aggregate_time_sharded
        lat, lon = synthetic.grid_labels(4.0)
        df = synthetic.weights_table(4.0, 60, seed=2)
        tas, _, _ = synthetic.tas_field(T, len(lat), len(lon), seed=1, dtype=np.float64)
        ds = Dataset({"tas": (("time", "lat", "lon"), tas)}, coords={"time": np.arange(T), "lat": lat, "lon": lon})
        full = aggregate_time_sharded(ds, "tas", "popwt", "hierid", df, aggregate_fn=_oracle_aggregate)
        part = aggregate_time_sharded(ds, "tas", "popwt", "hierid", df, gather=False, aggregate_fn=_oracle_aggregate)
        ref = oracle.weighted_aggregate_grid_to_regions(tas, ("time", "lat", "lon"), lat, lon, df, "popwt", "hierid")[0]
        ok = np.array_equal(full["tas"].values, ref, equal_nan=True) and full["tas"].dims == ("time", "hierid") \
            and np.array_equal(full["time"].values, np.arange(T)) and part["tas"].shape[0] < T
        q.put((rank, bool(ok), tuple(full["tas"].shape)))
    finally:
        dist.destroy_process_group()


def test_aggregate_time_sharded_api_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_api, args=(r, 2, port, 70, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert all(r[1] for r in res), res


def _oracle_aggregate_nd(ds, variable, aggwt, agglev, weights=None, backup_aggwt="areawt", **kw):
    import oracle
    from climate_toolbox_b200 import Dataset
    x = ds[variable].values
    out, rd, labels = oracle.weighted_aggregate_grid_to_regions(
        x, ds[variable].dims, ds["lat"].values, ds["lon"].values, weights, aggwt, agglev, backup_aggwt)
    return Dataset({variable: (rd, out)}, coords={"time": ds["time"].values, "model": ds["model"].values,
                                                  agglev: labels})


def _worker_leading_dims(rank, world, port, T, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import oracle
        from climate_toolbox_b200 import Dataset, synthetic
        from climate_toolbox_b200.parallel import aggregate_time_sharded
        lat, lon = synthetic.grid_labels(4.0)
        df = synthetic.weights_table(4.0, 60, seed=2)
        tas, _, _ = synthetic.tas_field(3 * T, len(lat), len(lon), seed=1, dtype=np.float64)
        x = tas.reshape(3, T, len(lat), len(lon))
        dims = ("model", "time", "lat", "lon")
        ds = Dataset({"tas": (dims, x)}, coords={"model": ["a", "b", "c"], "time": np.arange(T), "lat": lat, "lon": lon})
        full = aggregate_time_sharded(ds, "tas", "popwt", "hierid", df, aggregate_fn=_oracle_aggregate_nd)
        ref, rd, _ = oracle.weighted_aggregate_grid_to_regions(x, dims, lat, lon, df, "popwt", "hierid")
        ok = full["tas"].dims == rd == ("model", "time", "hierid") and \
            np.array_equal(full["tas"].values, ref, equal_nan=True)
        q.put((rank, bool(ok), tuple(full["tas"].shape)))
    finally:
        dist.destroy_process_group()


def test_aggregate_time_sharded_leading_dims_world2():
    """A variable with a leading (model) dim: everything but the time dim rides along in the gather."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_leading_dims, args=(r, 2, port, 70, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert all(r[1] for r in res), res
