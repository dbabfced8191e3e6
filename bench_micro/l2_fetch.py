"""Config-2 kernel time against the L2 fetch granularity hint (cudaLimitMaxL2FetchGranularity)."""
import ctypes
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
x = 288.0 + 10.0 * torch.randn((T, len(lat) * len(lon)), generator=g, device=dev, dtype=torch.float32)
plan = E.get_plan(E.GridSpec(lat, lon), df, "popwt", "hierid", device=dev)
out = torch.empty((1, plan.R, T), dtype=torch.float64, device=dev)
balg = plan.algorithmic_bytes(T, 1, 4, 1)
rt = ctypes.CDLL("libcudart.so.12")
LIMIT = 0x05


def get():
    v = ctypes.c_size_t(0)
    rc = rt.cudaDeviceGetLimit(ctypes.byref(v), LIMIT)
    return rc, v.value


def run(label):
    for _ in range(3):
        E.aggregate_device(plan, x, None, N.LAYOUT_TIME_MAJOR, x.shape[1], None, T, out=out)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(11)]
    ev[0].record()
    for i in range(10):
        E.aggregate_device(plan, x, None, N.LAYOUT_TIME_MAJOR, x.shape[1], None, T, out=out)
        ev[i + 1].record()
    torch.cuda.synchronize()
    ms = np.array([ev[i].elapsed_time(ev[i + 1]) for i in range(10)])
    print("{:34s} {:7.3f} ms (min {:.3f})  frac {:.3f}  checksum {:.6f}".format(
        label, ms.mean(), ms.min(), balg / ms.mean() / 1e6 / 6537, float(out.sum())), flush=True)


print("default limit (rc, bytes):", get())
run("default")
for v in (32, 64, 128, 32):
    rc = rt.cudaDeviceSetLimit(LIMIT, ctypes.c_size_t(v))
    print("set", v, "rc", rc, "now", get())
    run("granularity %d" % v)
