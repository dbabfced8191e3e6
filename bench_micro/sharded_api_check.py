"""torchrun check of parallel.aggregate_time_sharded on real GPUs (NCCL): every rank aggregates its
block of days with the CUDA path; the gathered result must match the oracle on every rank.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 bench_micro/sharded_api_check.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import oracle
from climate_toolbox_b200 import Dataset, synthetic
from climate_toolbox_b200.parallel import aggregate_time_sharded
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
T = 100
lat, lon = synthetic.grid_labels(1.0)
df = synthetic.weights_table(1.0, 3000)
tas, _, _ = synthetic.tas_field(T, len(lat), len(lon), seed=7, nan_frac=0.001, dtype=np.float32)
for where in ("host", "device"):
    data = tas if where == "host" else torch.from_numpy(tas).cuda()
    ds = Dataset({"tas": (("time", "lat", "lon"), data)}, coords={"time": np.arange(T), "lat": lat, "lon": lon})
    full = aggregate_time_sharded(ds, "tas", "popwt", "hierid", df, keep_on_device=(where == "device"))
    got = full["tas"].values
    ref = oracle.weighted_aggregate_grid_to_regions(tas, ("time", "lat", "lon"), lat, lon, df, "popwt", "hierid")[0]
    err = np.nanmax(np.abs(got - ref) / np.maximum(np.abs(ref), 1e-300))
    print("rank %d %s: shape %s max rel err %.2e %s" % (dist.get_rank(), where, got.shape, err,
                                                        "OK" if err < 1e-9 else "FAIL"), flush=True)
dist.destroy_process_group()
