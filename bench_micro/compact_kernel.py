"""Fused kernel on a PACKED (compact-plan) device-resident input vs the full grid."""
import ctypes as C, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as G
G.build()
from climate_toolbox_b200 import synthetic, _engine as E, _native as N
T = 1460
lat, lon = synthetic.grid_labels(0.25)
df = synthetic.weights_table(0.25, 24378)
dev = torch.device("cuda", 0)
ncell = len(lat) * len(lon)
x = 288 + 10 * torch.randn((T, ncell), dtype=torch.float32, device=dev)
def timeit(plan, xin, stride):
    out = torch.empty((1, plan.R, T), dtype=torch.float64, device=dev)
    for _ in range(3): E.aggregate_device(plan, xin, None, N.LAYOUT_TIME_MAJOR, stride, None, T, out=out)
    ms = []
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); E.aggregate_device(plan, xin, None, N.LAYOUT_TIME_MAJOR, stride, None, T, out=out); e1.record()
        torch.cuda.synchronize(); ms.append(e0.elapsed_time(e1))
    return min(ms), out
full = E.get_plan(E.GridSpec(lat, lon), df, "popwt", "hierid")
t_full, o_full = timeit(full, x, ncell)
comp = E.get_plan(E.GridSpec(lat, lon), df, "popwt", "hierid", compact=True)
w = comp.info["n_packed_cells"]
# pack on the device with torch (index_select of the packed cells) -- experiment only
xh = x[:8].cpu().numpy()
dst = np.empty((8, w), dtype=np.float32)
N.check(N.lib().ctb_host_pack(comp._h, C.c_void_p(xh.ctypes.data), N.F32, ncell, None, 0, 8, C.c_void_p(dst.ctypes.data), 0))
# recover the packed->physical cell map from a packed arange
idx = np.arange(ncell, dtype=np.float32)[None, :].repeat(1, 0)
pm = np.empty((1, w), dtype=np.float32)
N.check(N.lib().ctb_host_pack(comp._h, C.c_void_p(idx.ctypes.data), N.F32, ncell, None, 0, 1, C.c_void_p(pm.ctypes.data), 0))
cells = torch.from_numpy(pm[0].astype(np.int64)).to(dev)
xc = x.index_select(1, cells).contiguous()
t_comp, o_comp = timeit(comp, xc, w)
print("full grid: %.3f ms   packed input: %.3f ms   equal: %s" % (t_full, t_comp, torch.allclose(o_full, o_comp, rtol=1e-12, equal_nan=True)))
