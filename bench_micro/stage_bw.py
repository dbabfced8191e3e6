"""Loads-only replay of the staging traffic on the real plan footprint (config 2)."""
import ctypes as C, sys, os, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as G
G.build()
from climate_toolbox_b200 import synthetic, _engine as E, _native as N

T = int(os.environ.get("T", 736))
lat, lon = synthetic.grid_labels(0.25)
df = synthetic.weights_table(0.25, 24378)
dev = torch.device("cuda", 0)
x = torch.zeros((T, len(lat) * len(lon)), dtype=torch.float32, device=dev)
sink = torch.zeros(1, dtype=torch.float32, device=dev)
for budget in (0, 40 * 1024):
    plan = E.get_plan(E.GridSpec(lat, lon), df, "popwt", "hierid", smem_budget=budget, cache=False)
    i = plan.info
    nbytes = i["n_pieces"] * 16 * T
    print("budget", budget, "bundles", i["n_bundles"], "staged pieces", i["n_pieces"], "distinct", i["n_pieces_distinct"],
          "staged GB", nbytes / 1e9, flush=True)
    for lanes_p, unr, warps, cps in itertools.product((8, 16, 32), (4, 8, 16), (16, 32), (1, 2)):
        if warps * cps > 64 or (unr == 16 and warps > 16) or (unr == 16 and cps > 1):
            continue
        ms = []
        for rep in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            N.check(N.lib().ctb_debug_stage_bw(plan._h, C.c_void_p(x.data_ptr()), x.shape[1], T, lanes_p, unr, warps,
                                               cps, C.c_void_p(sink.data_ptr()), None))
            e1.record(); torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        print("  lanes_p %2d unr %2d warps %2d cps %d : %6.0f GB/s" % (lanes_p, unr, warps, cps, nbytes / min(ms) / 1e6), flush=True)
    plan.close()
