"""Loads-only replay: effect of the plane stride (DRAM bank camping across day-planes)
and of the L2 fetch granularity limit."""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as G
G.build()
from climate_toolbox_b200 import synthetic, _engine as E, _native as N

T = 736
lat, lon = synthetic.grid_labels(0.25)
df = synthetic.weights_table(0.25, 24378)
dev = torch.device("cuda", 0)
ncell = len(lat) * len(lon)
sink = torch.zeros(1, dtype=torch.float32, device=dev)
plan = E.get_plan(E.GridSpec(lat, lon), df, "popwt", "hierid", cache=False)
nbytes = plan.info["n_pieces"] * 16 * T
def run(stride, lanes_p=8, unr=8, warps=16, cps=2):
    x = torch.zeros((T, stride), dtype=torch.float32, device=dev)
    ms = []
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        N.check(N.lib().ctb_debug_stage_bw(plan._h, C.c_void_p(x.data_ptr()), stride, T, lanes_p, unr, warps, cps,
                                           C.c_void_p(sink.data_ptr()), None))
        e1.record(); torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    del x
    return nbytes / min(ms) / 1e6
for gran in (None, "32", "64", "128"):
    if gran: os.environ["CTB_L2_FETCH"] = gran
    for pad in (0, 64, 256, 1024, 2048 + 64, 4096 + 256, 16384 + 1024 + 64):
        r = [run(ncell + pad, 8, 8, 16, 2), run(ncell + pad, 8, 8, 16, 1), run(ncell + pad, 32, 8, 16, 2)]
        print("gran", gran, "pad elems", pad, "GB/s (16w x2, 16w x1, lanes32 16w x2):", " ".join("%6.0f" % v for v in r), flush=True)
