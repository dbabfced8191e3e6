"""Loads-only replay with an L2 prefetch of the same footprint N time blocks ahead.
NOTE: in this microkernel consecutive time blocks of a bundle are n_bundles items apart,
i.e. ~n_bundles/296 waves later."""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as G
G.build()
from climate_toolbox_b200 import synthetic, _engine as E, _native as N
lat, lon = synthetic.grid_labels(0.25)
df = synthetic.weights_table(0.25, 24378)
dev = torch.device("cuda", 0)
ncell = len(lat) * len(lon)
T = 1460
sink = torch.zeros(1, dtype=torch.float32, device=dev)
plan = E.get_plan(E.GridSpec(lat, lon), df, "popwt", "hierid", cache=False)
x = 288 + 10 * torch.randn((T, ncell), dtype=torch.float32, device=dev)
def run(lanes_p=8, unr=8, warps=16, cps=2):
    nbytes = plan.info["n_pieces"] * 16 * T
    ms = []
    for rep in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        N.check(N.lib().ctb_debug_stage_bw(plan._h, C.c_void_p(x.data_ptr()), ncell, T, lanes_p, unr, warps, cps,
                                           C.c_void_p(sink.data_ptr()), None))
        e1.record(); torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    return nbytes / min(ms) / 1e6, min(ms)
print("bundles", plan.info["n_bundles"])
os.environ["CTB_DBG_ORDER"] = "2"
os.environ["CTB_DBG_SMEM"] = "76000"
for sync in (0, 1):
    os.environ["CTB_DBG_SYNC"] = str(sync)
    for unr, warps, cps in ((8, 16, 2), (4, 16, 2), (8, 8, 2)):
        print("barrier per tile", sync, "unr", unr, "warps", warps, "cps", cps, "GB/s %.0f  ms %.3f" % run(8, unr, warps, cps), flush=True)
