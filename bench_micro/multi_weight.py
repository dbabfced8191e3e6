"""several weight columns in one pass (virtual regions) vs one pass per column.
Config 2 shape, device-resident input, results kept on the device; CUDA-event timing."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as G
G.build()
from climate_toolbox_b200 import Dataset, synthetic, _engine as E
from climate_toolbox_b200.aggregations.aggregations import (
    weighted_aggregate_grid_to_regions, weighted_aggregate_grid_to_regions_multi)
T = 1460
lat, lon = synthetic.grid_labels(0.25)
df = synthetic.weights_table(0.25, 24378)
x = 288.0 + 10.0 * torch.randn((T, len(lat), len(lon)), device="cuda", dtype=torch.float32)
ds = Dataset({"tas": (("time", "lat", "lon"), x)}, coords={"time": np.arange(T), "lat": lat, "lon": lon})


def timed(f, n=10):
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        f()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


from climate_toolbox_b200 import _native as N
from climate_toolbox_b200.aggregations.aggregations import _STACKED
ncell = len(lat) * len(lon)
x2 = x.view(T, ncell)


def kernel_ms(plan):
    out = torch.empty((1, plan.R, T), dtype=torch.float64, device="cuda")
    ws = torch.empty(max(1, N.lib().ctb_aggregate_workspace_bytes(plan._h, T, 1) // 8), dtype=torch.float64, device="cuda")
    return timed(lambda: E.aggregate_device(plan, x2, None, N.LAYOUT_TIME_MAJOR, ncell, None, T, "identity", (), 1,
                                            0, out=out, workspace=ws), 20)


grid = E.GridSpec(lat, lon)
for cols in (["popwt"], ["popwt", "areawt"], ["popwt", "areawt", "cropwt"]):
    sep = sum(kernel_ms(E.get_plan(grid, df, c, "hierid")) for c in cols)
    weighted_aggregate_grid_to_regions_multi(ds, "tas", cols, "hierid", df, keep_on_device=True)   # builds the stacked frame
    st = list(_STACKED.values())[-1][1]
    plan = E.get_plan(grid, st, "_w", "_lev", "_bk")
    one = kernel_ms(plan)
    api = timed(lambda: weighted_aggregate_grid_to_regions_multi(ds, "tas", cols, "hierid", df, keep_on_device=True))
    print("%d column(s): kernels of separate passes %.3f ms, one fused pass %.3f ms (through the API incl. host "
          "overhead %.3f ms); bundles %d, staged pieces %d" % (len(cols), sep, one, api, plan.info["n_bundles"],
                                                               plan.info["n_pieces"]), flush=True)
