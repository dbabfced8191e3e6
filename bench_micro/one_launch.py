"""A few launches of the config-2 kernel (for ncu): python bench_micro/one_launch.py [identity|poly|edd|packed] [n]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from climate_toolbox_b200 import _engine as E, _native as N, synthetic  # noqa: E402

kind = sys.argv[1] if len(sys.argv) > 1 else "identity"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 5
T = int(os.environ.get("SWEEP_T", "1460"))
lat, lon = synthetic.grid_labels(0.25)
df = synthetic.weights_table(0.25, 24378)
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(7)
x = 288.0 + 10.0 * torch.randn((T, len(lat) * len(lon)), generator=g, device=dev, dtype=torch.float32)
x1 = None
aggwt, params, n_out, sb = "popwt", (), 1, 4
if kind == "poly":
    params, n_out = (273.15, 1, 2, 3, 4), 4
elif kind == "edd":
    x1 = x + (3.0 * torch.randn(x.shape, generator=g, device=dev, dtype=torch.float32)).abs()
    x = x - (3.0 * torch.randn(x.shape, generator=g, device=dev, dtype=torch.float32)).abs()
    params, n_out, aggwt, sb = (283.15, 303.15), 2, "cropwt", 8
packed = kind == "packed"
if packed:
    kind = "identity"
plan = E.get_plan(E.GridSpec(lat, lon), df, aggwt, "hierid", device=dev, stage_bytes=sb, compact=packed)
if packed:   # compact the archive on the device once, then aggregate the packed planes
    import ctypes as C
    xp = torch.empty((T, plan.info["n_packed_cells"]), dtype=torch.float32, device=dev)
    N.check(N.lib().ctb_pull_pack(plan._h, C.c_void_p(x.data_ptr()), N.F32, x.shape[1], None, 0, T,
                                  C.c_void_p(xp.data_ptr()), E._stream_ptr(dev)))
    x = xp
out = torch.empty((n_out, plan.R, T), dtype=torch.float64, device=dev)
for _ in range(n):
    E.aggregate_device(plan, x, x1, N.LAYOUT_TIME_MAJOR, x.shape[1], None, T, kind, params, n_out, out=out)
torch.cuda.synchronize()
print("ok", float(out.nansum()))
