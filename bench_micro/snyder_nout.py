"""Streaming kernel, Snyder forms with 1..4 outputs (EDD) and 1..2 (GDD), 730 days of config 4's inputs."""
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
thr = (283.15, 303.15, 288.15, 298.15, 293.15, 300.15, 281.15, 305.15)
print("library:", os.environ.get("CTB_LIBRARY", "libctb.so"))
for kind, n_out in (("edd", 1), ("edd", 2), ("edd", 3), ("edd", 4), ("gdd", 1), ("gdd", 2)):
    params = thr[:n_out] if kind == "edd" else thr[:2 * n_out]
    out = torch.empty((n_out, plan.R, T), dtype=torch.float64, device=dev)
    f = lambda: E.aggregate_device(plan, lo, hi, N.LAYOUT_TIME_MAJOR, lo.shape[1], None, T, kind, params, n_out, out=out)
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
    print("{} n_out={}: {:.3f} ms (min {:.3f})  checksum {:.6e}".format(kind, n_out, ms.mean(), ms.min(),
                                                                       float(torch.nansum(out))), flush=True)
