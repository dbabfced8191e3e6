"""Per-phase cycle breakdown of the fused kernel (CTB_DEBUG=16 [+ other bits])."""
import ctypes as C, sys, os, subprocess, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as G
G.build()
from climate_toolbox_b200 import synthetic, _engine as E, _native as N
T = 1460
lat, lon = synthetic.grid_labels(0.25)
df = synthetic.weights_table(0.25, 24378)
dev = torch.device("cuda", 0)
x = 288 + 10 * torch.randn((T, len(lat) * len(lon)), dtype=torch.float32, device=dev)
plan = E.get_plan(E.GridSpec(lat, lon), df, "popwt", "hierid")
out = torch.empty((1, plan.R, T), dtype=torch.float64, device=dev)
names = ["metadata", "wait-gather", "stage", "wait-stage", "gather"]
for dbg in (16, 16 + 8, 16 + 4, 16 + 1):
    os.environ["CTB_DEBUG"] = str(dbg)
    for _ in range(3):
        E.aggregate_device(plan, x, None, N.LAYOUT_TIME_MAJOR, x.shape[1], None, T, out=out)
    torch.cuda.synchronize()
    buf = (C.c_uint64 * 8)()
    N.check(N.lib().ctb_debug_timers(buf, 1))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    E.aggregate_device(plan, x, None, N.LAYOUT_TIME_MAJOR, x.shape[1], None, T, out=out)
    e1.record(); torch.cuda.synchronize()
    N.check(N.lib().ctb_debug_timers(buf, 1))
    n = buf[5]
    tot = sum(buf[i] for i in range(5))
    print("dbg", dbg, "ms %.3f" % e0.elapsed_time(e1), "ctas", n, "cycles/CTA %.0f" % (tot / max(n, 1)),
          " ".join("%s %.1f%%" % (names[i], 100.0 * buf[i] / tot) for i in range(5)), flush=True)
