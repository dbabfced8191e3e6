import os, sys, time, cProfile, pstats
sys.path.insert(0, os.getcwd())
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
pr = cProfile.Profile()
pr.enable()
for _ in range(5):
    r = weighted_aggregate_grid_to_regions(ds, "tas", "popwt", "hierid", weights=df)
torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr); st.sort_stats("cumulative").print_stats(28)
