"""Config-2 kernel time of the streaming kernel under experiment knobs (needs libctb_exp.so):
CTB_KNOBS 1 = skip the copies (reduction only), 4 = skip the reduction (copies only);
CTB_CHUNK_TB time blocks per work unit."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("CTB_LIBRARY", os.path.join(ROOT, "bench_micro", "libctb_exp.so"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from climate_toolbox_b200 import _engine as E, _native as N, synthetic  # noqa: E402

T = int(os.environ.get("SWEEP_T", "1460"))
lat, lon = synthetic.grid_labels(0.25)
df = synthetic.weights_table(0.25, 24378)
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(7)
x = 288.0 + 10.0 * torch.randn((T, len(lat) * len(lon)), generator=g, device=dev, dtype=torch.float32)
plan = E.get_plan(E.GridSpec(lat, lon), df, "popwt", "hierid", device=dev)
print({k: plan.info[k] for k in ("n_bundles", "n_pieces", "n_pieces_distinct", "n_quads", "n_quads_conflict", "nnz")})
out = torch.empty((1, plan.R, T), dtype=torch.float64, device=dev)
balg = plan.algorithmic_bytes(T, 1, 4, 1)


def run(label, **env):
    for k in ("CTB_KNOBS", "CTB_CHUNK_TB"):
        os.environ.pop(k, None)
    for k, v in env.items():
        os.environ[k] = str(v)
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
    print("{:34s} {:7.3f} ms (min {:.3f})  {:6.0f} GB/s algorithmic  frac {:.3f}".format(
        label, ms.mean(), ms.min(), balg / ms.mean() / 1e6, balg / ms.mean() / 1e6 / 6537), flush=True)


print("library:", os.environ["CTB_LIBRARY"])
run("default")
run("reduce only (no copies)", CTB_KNOBS=1)
run("copies only (no reduce)", CTB_KNOBS=4)
if os.environ.get("SWEEP_CHUNKS"):
    for c in (4, 12, 16):
        run("chunk_tb={}".format(c), CTB_CHUNK_TB=c)
