"""Throughput of ctb_host_pack (host-side packing of the referenced gridcells) vs threads."""
import ctypes as C, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as G
G.build()
from climate_toolbox_b200 import synthetic, _engine as E, _native as N
T = 730
lat, lon = synthetic.grid_labels(0.25)
df = synthetic.weights_table(0.25, 24378)
plan = E.get_plan(E.GridSpec(lat, lon), df, "popwt", "hierid", compact=True)
w = plan.info["n_packed_cells"]
src = torch.empty((T, len(lat) * len(lon)), dtype=torch.float32, pin_memory=True)
src.normal_()
dst = torch.empty((T, w), dtype=torch.float32, pin_memory=True)
print("packed cells", w, "of", src.shape[1], "cpus", os.cpu_count(), "runs per plane ~", "n/a")
for nt, th in [(n, t) for t in (1, 4, 8, 16, 0) for n in ("1", "0")] * 2:
    os.environ["CTB_PACK_NT"] = nt     # 1: streaming whole-line stores (default), 0: memcpy
    best = 1e9
    for rep in range(3):
        t = time.perf_counter()
        N.check(N.lib().ctb_host_pack(plan._h, C.c_void_p(src.data_ptr()), N.F32, src.shape[1], None, 0, T,
                                      C.c_void_p(dst.data_ptr()), th))
        best = min(best, time.perf_counter() - t)
    print("nt=%s threads %2d: %.1f ms  %.1f GB/s packed" % (nt, th, best * 1e3, T * w * 4 / best / 1e9), flush=True)
# plain big memcpy for reference
a = np.empty(T * w, dtype=np.float32); b = np.empty_like(a)
t = time.perf_counter(); b[:] = a; dt = time.perf_counter() - t
print("single-thread numpy copy of the same bytes: %.1f ms (%.1f GB/s)" % (dt * 1e3, a.nbytes / dt / 1e9))
