"""NVLink push of one rank's column block of the [R][T] result to every peer: ctb_push_rows (SM kernel)
against 2-D DMA copies and a contiguous DMA copy of the same bytes (the ceiling).
torchrun --nproc-per-node N bench_micro/push_bw.py"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from climate_toolbox_b200 import _engine as E, _native as N, synthetic  # noqa: E402
from climate_toolbox_b200.parallel import PeerOutput, shard_range  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
T = 1460
lat, lon = synthetic.grid_labels(0.25)
df = synthetic.weights_table(0.25, 24378)
plan = E.get_plan(E.GridSpec(lat, lon), df, "popwt", "hierid", device=dev)
po = PeerOutput(plan, T)
po.raw.normal_()
t0, t1 = shard_range(T, world, rank)
n = t1 - t0
rt = C.CDLL("libcudart.so.12")
rt.cudaMemcpy2DAsync.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, C.c_void_p]
rt.cudaMemcpyAsync.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
streams = [torch.cuda.Stream(dev) for _ in range(world)]
nbytes = 8 * n * plan.R


def timed(fn, label):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / 10], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        print("{:44s} {:7.3f} ms  {:6.0f} GB/s out per GPU".format(label, ms.item(), nbytes * (world - 1) / ms.item() / 1e6),
              flush=True)


def push_kernel():
    po.push(t0, n)


def dma_2d(parallel):
    def f():
        cur = torch.cuda.current_stream(dev)
        ev = torch.cuda.Event()
        ev.record(cur)
        k = 0
        for r in range(world):
            if r == rank:
                continue
            st = streams[k] if parallel else cur
            k += 1
            if parallel:
                st.wait_event(ev)
            rc = rt.cudaMemcpy2DAsync(po.ptrs[r] + 8 * t0, 8 * T, po._own + 8 * t0, 8 * T, 8 * n, plan.R, 4, st.cuda_stream)
            assert rc == 0, rc
            if parallel:
                e2 = torch.cuda.Event()
                e2.record(st)
                cur.wait_event(e2)
    return f


def dma_flat():
    cur = torch.cuda.current_stream(dev)
    for r in range(world):
        if r != rank:
            rc = rt.cudaMemcpyAsync(po.ptrs[r] + rank * nbytes, po._own, nbytes, 4, cur.cuda_stream)
            assert rc == 0, rc


if rank == 0:
    print("world", world, "bytes per peer", nbytes)
timed(push_kernel, "ctb_push_rows (SM kernel)")
timed(dma_2d(False), "2-D DMA, one stream")
timed(dma_2d(True), "2-D DMA, one stream per peer")
timed(dma_flat, "contiguous DMA, one stream (ceiling)")
po.close()
dist.destroy_process_group()
