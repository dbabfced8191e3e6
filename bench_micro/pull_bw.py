"""ctb_pull_pack alone: the GPU reads the referenced pieces of T pinned host planes over PCIe.
CTB_LIBRARY selects a build (e.g. one with -DCTB_PULL_L2_HINT='".L2::64B"')."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from climate_toolbox_b200 import _engine as E, _native as N, synthetic  # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 365
lat, lon = synthetic.grid_labels(0.25)
df = synthetic.weights_table(0.25, 24378)
dev = torch.device("cuda", 0)
host = torch.empty((T, len(lat) * len(lon)), dtype=torch.float32, pin_memory=True)
host.copy_(288.0 + 10.0 * torch.randn(host.shape, device=dev))
plan = E.get_plan(E.GridSpec(lat, lon), df, "popwt", "hierid", device=dev, compact=True)
width = plan.info["n_packed_cells"]
dst = torch.empty((T, width), dtype=torch.float32, device=dev)
useful = 16 * plan.info["n_pieces_distinct"] * T if "n_pieces_distinct" in plan.info else 4 * width * T


def pull():
    N.check(N.lib().ctb_pull_pack(plan._h, C.c_void_p(host.data_ptr()), N.F32, host.shape[1], None, 0, T,
                                  C.c_void_p(dst.data_ptr()), E._stream_ptr(dev)))


for _ in range(2):
    pull()
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
ev[0].record()
for i in range(5):
    pull()
    ev[i + 1].record()
torch.cuda.synchronize()
ms = np.array([ev[i].elapsed_time(ev[i + 1]) for i in range(5)])
print("{}: pull of {} days: {:.2f} ms (min {:.2f}), {:.1f} GB/s of packed bytes, checksum {:.6e}".format(
    os.environ.get("CTB_LIBRARY", "libctb.so"), T, ms.mean(), ms.min(), 4 * width * T / ms.mean() / 1e6, float(dst.sum())), flush=True)
