"""Sweep tile size (plan smem budget) and chunk length of the fused kernel.

The round-1 sweep (profiles/micro/r1_sweep1.log) also compared CTA shapes: 256- vs 512-thread
per-unit CTAs ("variant 3/4") and the warp-specialised persistent kernel ("variant 1" then);
the 512-thread per-unit CTA won and is what ctb_aggregate runs now."""
import ctypes as C, sys, os
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
ref = None
for budget, threads in ((67584, 512), (50688, 512), (33792, 512)):   # r1_sweep2.log also had 1024-thread CTAs
    plan = E.get_plan(E.GridSpec(lat, lon), df, "popwt", "hierid", smem_budget=budget, cache=False)
    out = torch.empty((1, plan.R, T), dtype=torch.float64, device=dev)
    i = plan.info
    for variant in (1,):
        for chunk in (1, 2, 4):
            os.environ["CTB_CHUNK_TB"] = str(chunk)
            try:
                for _ in range(3):
                    E.aggregate_device(plan, x, None, N.LAYOUT_TIME_MAJOR, x.shape[1], None, T, out=out, variant=variant)
                torch.cuda.synchronize()
                ms = []
                for rep in range(5):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    E.aggregate_device(plan, x, None, N.LAYOUT_TIME_MAJOR, x.shape[1], None, T, out=out, variant=variant)
                    e1.record(); torch.cuda.synchronize()
                    ms.append(e0.elapsed_time(e1))
                if ref is None:
                    ref = out.clone()
                ok = torch.allclose(out, ref, rtol=1e-12, atol=0, equal_nan=True)
                print("threads %4d tile %6d B cells %4d bundles %4d staged %6d | variant %d chunk %2d : %.3f ms  ok=%s" %
                      (threads, budget, i["max_bundle_cells"], i["n_bundles"], i["n_pieces"], variant, chunk, min(ms), ok), flush=True)
            except Exception as ex:
                print("budget", budget, "variant", variant, "failed:", str(ex)[:100], flush=True)
    plan.close()
