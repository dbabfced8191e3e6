"""End to end through the public API from PINNED host memory: host packing (ctb_host_pack + chunked
H2D) against the GPU pull (ctb_pull_pack).  python bench_micro/e2e_ingest.py [T]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from climate_toolbox_b200 import Dataset, _engine as E, synthetic  # noqa: E402
from climate_toolbox_b200.aggregations.aggregations import weighted_aggregate_grid_to_regions  # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 1460
lat, lon = synthetic.grid_labels(0.25)
df = synthetic.weights_table(0.25, 24378)
host = torch.empty((T, len(lat), len(lon)), dtype=torch.float32, pin_memory=True)
g = torch.Generator(device="cuda").manual_seed(7)
for a in range(0, T, 365):
    b = min(T, a + 365)
    host[a:b].copy_(288.0 + 10.0 * torch.randn((b - a, len(lat), len(lon)), generator=g, device="cuda"))
torch.cuda.synchronize()
ds = Dataset({"tas": (("time", "lat", "lon"), host.numpy())}, coords={"time": np.arange(T), "lat": lat, "lon": lon})
res = {}
for mode in ("pack", "pull", "pack", "pull"):
    ts = []
    for i in range(8):
        t0 = time.perf_counter()
        r = weighted_aggregate_grid_to_regions(ds, "tas", "popwt", "hierid", weights=df, ingest=mode)
        ts.append((time.perf_counter() - t0) * 1e3)
    res.setdefault(mode, []).append(float(np.nansum(r["tas"].values)))
    print(mode, "ms per call:", " ".join("%.1f" % t for t in ts), "| median of last 6: %.1f" % np.median(ts[2:]), flush=True)
    ys = []
    for i in range(4):
        t0 = time.perf_counter()
        weighted_aggregate_grid_to_regions(ds, "tas", "popwt", "hierid", weights=df, ingest=mode, time_groups=365)
        ys.append((time.perf_counter() - t0) * 1e3)
    print(mode, "annual sums ms:", " ".join("%.1f" % t for t in ys), flush=True)
print("checksums", res, "threads", E.pack_threads())
