"""Config-2 kernel with and without a time index (leap days removed / a ring of year buffers)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from climate_toolbox_b200 import _engine as E, _native as N, synthetic  # noqa: E402

T = 1460
lat, lon = synthetic.grid_labels(0.25)
df = synthetic.weights_table(0.25, 24378)
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(7)
x = 288.0 + 10.0 * torch.randn((T + 1, len(lat) * len(lon)), generator=g, device=dev, dtype=torch.float32)
plan = E.get_plan(E.GridSpec(lat, lon), df, "popwt", "hierid", device=dev)
out = torch.empty((1, plan.R, T), dtype=torch.float64, device=dev)
cases = {"no time index": None, "identity index": np.arange(T), "one day skipped (leap day)": np.delete(np.arange(T + 1), 59),
         "ring of 4 x 365 days, years 0..3": np.concatenate([(y % 4) * 365 + np.arange(365) for y in range(4)])}
for label, tix in cases.items():
    f = lambda: E.aggregate_device(plan, x, None, N.LAYOUT_TIME_MAJOR, x.shape[1], tix, T, out=out)
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(11)]
    ev[0].record()
    for i in range(10):
        f()
        ev[i + 1].record()
    torch.cuda.synchronize()
    ms = np.array([ev[i].elapsed_time(ev[i + 1]) for i in range(10)])
    print("{:36s} {:7.3f} ms (min {:.3f})  checksum {:.6f}".format(label, ms.mean(), ms.min(), float(out.sum())), flush=True)
