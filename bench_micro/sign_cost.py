"""Config-2 kernel against the sign of the data: the integer widening (bits * 2^29 + bias) is exact for
positive normal floats only; a region-tile that holds anything else is reduced again by the exact loop."""
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
z = torch.randn((T, len(lat) * len(lon)), generator=g, device=dev, dtype=torch.float32)
plan = E.get_plan(E.GridSpec(lat, lon), df, "popwt", "hierid", device=dev)
out = torch.empty((1, plan.R, T), dtype=torch.float64, device=dev)
for label, mean, sd in (("kelvin 288 +- 10 (all positive)", 288.0, 10.0), ("celsius 15 +- 10 (7 % negative)", 15.0, 10.0),
                        ("celsius 30 +- 10 (0.1 % negative)", 30.0, 10.0), ("anomaly 0 +- 10 (half negative)", 0.0, 10.0)):
    x = z * sd + mean
    f = lambda: E.aggregate_device(plan, x, None, N.LAYOUT_TIME_MAJOR, x.shape[1], None, T, out=out)
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
    print("{:38s} {:7.3f} ms (min {:.3f})  negative {:.2%}".format(label, ms.mean(), ms.min(), float((x < 0).float().mean())), flush=True)
    del x
