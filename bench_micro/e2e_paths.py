"""e2e (host buffers) of config 2 through the public API: host packing vs whole-chunk copies."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as G
G.build()
from climate_toolbox_b200 import Dataset, synthetic
from climate_toolbox_b200.aggregations.aggregations import weighted_aggregate_grid_to_regions as agg
T = 1460
lat, lon = synthetic.grid_labels(0.25)
df = synthetic.weights_table(0.25, 24378)
host = torch.empty((T, len(lat), len(lon)), dtype=torch.float32, pin_memory=True)
host.normal_(288, 10)
ds = Dataset({"tas": (("time", "lat", "lon"), host.numpy())}, coords={"time": np.arange(T), "lat": lat, "lon": lon})
for name, kw in (("packed", {}), ("whole chunks", {"pack_host": False}), ("packed", {}), ("whole chunks", {"pack_host": False})):
    agg(ds, "tas", "popwt", "hierid", weights=df, **kw)
    ts = []
    for _ in range(4):
        t = time.perf_counter(); r = agg(ds, "tas", "popwt", "hierid", weights=df, **kw); ts.append(time.perf_counter() - t)
    print("%-13s ms per call: %s" % (name, " ".join("%.1f" % (x * 1e3) for x in ts)), flush=True)
