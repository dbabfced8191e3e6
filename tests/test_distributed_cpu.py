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
