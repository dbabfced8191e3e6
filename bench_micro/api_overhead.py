"""Host-side cost of one public-API call on device-resident data (config 2; the kernel takes 0.6 ms)."""
import cProfile
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from climate_toolbox_b200 import Dataset, synthetic  # noqa: E402
from climate_toolbox_b200.aggregations.aggregations import weighted_aggregate_grid_to_regions  # noqa: E402

T = 1460
lat, lon = synthetic.grid_labels(0.25)
df = synthetic.weights_table(0.25, 24378)
x = 288.0 + 10.0 * torch.randn((T, len(lat), len(lon)), device="cuda", dtype=torch.float32)
ds = Dataset({"tas": (("time", "lat", "lon"), x)}, coords={"time": np.arange(T), "lat": lat, "lon": lon})
f = lambda: weighted_aggregate_grid_to_regions(ds, "tas", "popwt", "hierid", weights=df, keep_on_device=True)
for _ in range(5):
    f()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(50):
    f()
torch.cuda.synchronize()
print("per call: %.3f ms" % ((time.perf_counter() - t0) / 50 * 1e3))
pr = cProfile.Profile()
pr.enable()
for _ in range(50):
    f()
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(18)
