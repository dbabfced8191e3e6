"""Strong scaling of one 1460-day batch over the ranks (torchrun): compute only, compute + gather,
overlapped; and the raw all_gather.  torchrun --nproc-per-node N bench_micro/strong_scaling.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from climate_toolbox_b200 import _engine as E, synthetic  # noqa: E402
from climate_toolbox_b200.parallel import PeerOutput, aggregate_shard_overlapped, aggregate_shard_p2p  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
T = 1460
lat, lon = synthetic.grid_labels(0.25)
df = synthetic.weights_table(0.25, 24378)
g = torch.Generator(device=dev).manual_seed(7)
x = 288.0 + 10.0 * torch.randn((T, len(lat) * len(lon)), generator=g, device=dev, dtype=torch.float32)
plan = E.get_plan(E.GridSpec(lat, lon), df, "popwt", "hierid", device=dev)


def timed(fn, n=10):
    for _ in range(3):
        fn()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    dist.barrier(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / n], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


res = {}
for pieces in (1, 2, 4, 8):
    res["overlap%d" % pieces] = timed(lambda: aggregate_shard_overlapped(plan, x, None, x.shape[1], T, pieces=pieces))
res["compute_only"] = timed(lambda: aggregate_shard_overlapped(plan, x, None, x.shape[1], T, pieces=1, gather=False))
M = plan.R
tl = -(-T // world)
loc = torch.randn((M, tl), device=dev, dtype=torch.float64)
recv = torch.empty((world * M, tl), device=dev, dtype=torch.float64)
res["raw_all_gather"] = timed(lambda: dist.all_gather_into_tensor(recv, loc))
full = torch.empty((M, world * tl), device=dev, dtype=torch.float64)
res["strided_copy"] = timed(lambda: full.view(M, world, tl).copy_(recv.view(world, M, tl).permute(1, 0, 2)))
ref, _ = aggregate_shard_overlapped(plan, x, None, x.shape[1], T, pieces=1)
ref = ref.clone()
try:
    po = PeerOutput(plan, T)
    res["p2p_fused"] = timed(lambda: aggregate_shard_p2p(plan, x, None, x.shape[1], T, po))
    torch.cuda.synchronize()
    ok = torch.equal(torch.nan_to_num(po.gathered()), torch.nan_to_num(ref))
    res["p2p_equal_to_nccl_gather"] = float(ok)
    po.close()
except Exception as ex:  # noqa: BLE001
    res["p2p_error"] = -1.0
    if rank == 0:
        print("p2p failed:", repr(ex)[:300])
if rank == 0:
    print("world", world, {k: round(v, 3) for k, v in res.items()}, "ms; bytes received per rank",
          8 * M * (T - tl))
dist.destroy_process_group()
