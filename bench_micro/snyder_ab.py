"""Config 4 (Snyder EDD x 2, cropwt, 730 days): streaming kernel (variant 1) against the direct
kernel (variant 2); outputs compared.  (Round 2 used this script with variant 3 = the round-1
transposed-tile kernel, since removed: profiles/micro/r2_snyder.md.)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from climate_toolbox_b200 import _engine as E, _native as N, synthetic  # noqa: E402

T = int(os.environ.get("SWEEP_T", "730"))
lat, lon = synthetic.grid_labels(0.25)
df = synthetic.weights_table(0.25, 24378)
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(7)
tas = 288.0 + 10.0 * torch.randn((T, len(lat) * len(lon)), generator=g, device=dev, dtype=torch.float32)
hi = tas + (3.0 * torch.randn(tas.shape, generator=g, device=dev, dtype=torch.float32)).abs()
lo = tas - (3.0 * torch.randn(tas.shape, generator=g, device=dev, dtype=torch.float32)).abs()
del tas
plan = E.get_plan(E.GridSpec(lat, lon), df, "cropwt", "hierid", device=dev, stage_bytes=8)
print({k: plan.info[k] for k in ("n_bundles", "n_pieces", "nnz", "n_quads")})
outs = {}
for variant in (1, 2):
    out = torch.empty((2, plan.R, T), dtype=torch.float64, device=dev)
    f = lambda: E.aggregate_device(plan, lo, hi, N.LAYOUT_TIME_MAJOR, lo.shape[1], None, T, "edd", (283.15, 303.15), 2,
                                   variant=variant, out=out)
    for _ in range(2):
        f()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
    ev[0].record()
    for i in range(5):
        f()
        ev[i + 1].record()
    torch.cuda.synchronize()
    ms = np.array([ev[i].elapsed_time(ev[i + 1]) for i in range(5)])
    print("variant", variant, "%.3f ms (min %.3f)" % (ms.mean(), ms.min()), flush=True)
    outs[variant] = out
a, b = outs[1], outs[2]
ok = torch.isfinite(b)
print("max rel diff", float(((a - b).abs()[ok] / b.abs()[ok].clamp_min(1e-3)).max()), "nan pattern equal", bool((torch.isnan(a) == torch.isnan(b)).all()))
