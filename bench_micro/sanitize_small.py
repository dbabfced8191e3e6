"""Small end-to-end exercise of every kernel variant, meant to run under
`compute-sanitizer --tool memcheck` (one tool per gpurun call)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, pandas as pd, torch
import __graft_entry__ as G
G.build()
from climate_toolbox_b200 import Dataset, DataArray, synthetic, _native as N
from climate_toolbox_b200.aggregations.aggregations import weighted_aggregate_grid_to_regions as agg
from climate_toolbox_b200.transformations.transformations import tas_poly, snyder_edd, snyder_gdd
lat, lon = synthetic.grid_labels(2.0)
df = synthetic.weights_table(2.0, 120, seed=4)
tas, tmin, tmax = synthetic.tas_field(37, len(lat), len(lon), seed=1, nan_frac=0.01)
t = pd.date_range("2000-02-01", periods=37)
c = {"time": t, "lat": lat, "lon": lon}
for data in (tas, tas.astype(np.float64), torch.from_numpy(tas).cuda()):
    ds = Dataset({"tas": (("time", "lat", "lon"), data)}, coords=c)
    agg(ds, "tas", "popwt", "hierid", weights=df)
    agg(ds, "tas", "popwt", "ISO", weights=df, smem_budget=8 * 1024)       # split regions + fix-up
    agg(ds, "tas", "popwt", "hierid", weights=df, variant=N.VARIANT_DIRECT)
    n = ["a", "b", "c", "d"]
    agg(tas_poly(ds, [1, 2, 3, 4], n), n, "popwt", "hierid", weights=df)
tn = DataArray(torch.from_numpy(tmin).cuda(), dims=("time", "lat", "lon"), coords=c, attrs={"units": "K"})
tx = DataArray(torch.from_numpy(tmax).cuda(), dims=("time", "lat", "lon"), coords=c, attrs={"units": "K"})
ds = Dataset(coords=c)
ds["e1"] = snyder_edd(tn, tx, 283.15); ds["e2"] = snyder_edd(tn, tx, 300.0); ds["g"] = snyder_gdd(tn, tx, 283.15, 300.0)
agg(ds, ["e1", "e2", "g"], "cropwt", "hierid", weights=df)
ds["e1"].values
# odd grid: scalar staging loads
la, lo = np.arange(7) * 1.0, np.arange(9) * 1.0
x = np.random.default_rng(0).standard_normal((5, 7, 9)).astype(np.float32)
d2 = pd.DataFrame({"lat": la[[0, 1, 2, 3, 6]], "lon": lo[[0, 8, 4, 4, 8]], "hierid": list("aabbc"),
                   "popwt": [1.0, np.nan, 2.0, 0.0, 1.0], "areawt": [1.0] * 5})
agg(Dataset({"v": (("time", "lat", "lon"), x)}, coords={"time": np.arange(5), "lat": la, "lon": lo}),
    "v", "popwt", "hierid", weights=d2)
agg(Dataset({"v": (("lat", "lon", "time"), np.ascontiguousarray(x.transpose(1, 2, 0)))},
            coords={"time": np.arange(5), "lat": la, "lon": lo}), "v", "popwt", "hierid", weights=d2)
torch.cuda.synchronize()
print("sanitize_small: done")
