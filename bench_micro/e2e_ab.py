"""A/B of the end-to-end (host buffers) path on config 2: CTB_PACK_NT=1 (streaming whole-line stores in
ctb_host_pack) vs 0 (memcpy), alternating rounds so that the shared host's drift hits both arms."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as G
G.build()
from climate_toolbox_b200 import Dataset, synthetic
from climate_toolbox_b200.aggregations.aggregations import weighted_aggregate_grid_to_regions
T = 1460
lat, lon = synthetic.grid_labels(0.25)
df = synthetic.weights_table(0.25, 24378)
host = torch.empty((T, len(lat), len(lon)), dtype=torch.float32, pin_memory=True)
host.normal_(288.0, 10.0)
ds = Dataset({"tas": (("time", "lat", "lon"), host.numpy())}, coords={"time": np.arange(T), "lat": lat, "lon": lon})
for _ in range(3):
    weighted_aggregate_grid_to_regions(ds, "tas", "popwt", "hierid", weights=df)
res = {"1": [], "0": []}
for rnd in range(int(sys.argv[1]) if len(sys.argv) > 1 else 5):
    for nt in ("1", "0"):
        os.environ["CTB_PACK_NT"] = nt
        ts = []
        for _ in range(4):
            t = time.perf_counter()
            weighted_aggregate_grid_to_regions(ds, "tas", "popwt", "hierid", weights=df)
            torch.cuda.synchronize()
            ts.append((time.perf_counter() - t) * 1e3)
        res[nt] += ts
        print("round %d nt=%s: %s" % (rnd, nt, " ".join("%.1f" % x for x in ts)), flush=True)
for nt in ("1", "0"):
    a = np.array(res[nt])
    print("nt=%s  min %.1f  median %.1f  mean %.1f ms" % (nt, a.min(), np.median(a), a.mean()))
