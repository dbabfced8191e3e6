"""Can the copy engine gather the referenced row runs straight out of the pinned source array?
One cudaMemcpy2DAsync per run of consecutive referenced pieces: width = run bytes, height = all
days, source pitch = one day plane, destination = the packed plane layout.  No host packing, no
staging buffer: host memory is read once."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from climate_toolbox_b200 import synthetic
T = int(sys.argv[1]) if len(sys.argv) > 1 else 730
lat, lon = synthetic.grid_labels(0.25)
df = synthetic.weights_table(0.25, 24378)
cell = np.searchsorted(lat, df.lat.values) * len(lon) + np.searchsorted(lon, df.lon.values)
pieces = np.unique(cell // 4)
brk = np.flatnonzero(np.diff(pieces) != 1)
starts = np.concatenate([[0], brk + 1]); ends = np.concatenate([brk + 1, [len(pieces)]])
runs = [(int(pieces[s]), int(e - s)) for s, e in zip(starts, ends)]
ncell = len(lat) * len(lon)
src = torch.empty((T, ncell), dtype=torch.float32, pin_memory=True); src.normal_()
packed = 0; offs = []
for p, n in runs:
    packed = (packed + 3) & ~3; offs.append(packed); packed += n
packed = (packed + 3) & ~3
dst = torch.empty((T, packed * 4), dtype=torch.float32, device="cuda")
rt = C.CDLL("libcudart.so.12")
rt.cudaMemcpy2DAsync.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, C.c_void_p]
st = torch.cuda.current_stream().cuda_stream
print("runs", len(runs), "packed MB/day", packed * 16 / 1e6, "T", T)
for merge_gap in (0, 4, 16):
    # optionally merge runs separated by <= merge_gap pieces (fewer, longer rows; a little ocean crosses PCIe)
    m = []
    for (p, n), o in zip(runs, offs):
        if m and p - (m[-1][0] + m[-1][1]) <= merge_gap and merge_gap:
            m[-1][1] = p + n - m[-1][0]
        else:
            m.append([p, n, o])
    byts = sum(n for _, n, _ in m) * 16 * T
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        o2 = 0
        for p, n, o in m:
            rc = rt.cudaMemcpy2DAsync(C.c_void_p(dst.data_ptr() + o2 * 16), packed * 16, C.c_void_p(src.data_ptr() + p * 16),
                                      ncell * 4, n * 16, T, 1, C.c_void_p(st))
            assert rc == 0, rc
            o2 = (o2 + n + 3) & ~3 if merge_gap else o + n
            if o2 > packed - 4096: o2 = 0
        t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
        print("gap %2d: %5d copies, issue %.1f ms, total %.1f ms, %.1f GB/s (%.2f GB)" % (merge_gap, len(m), (t1 - t0) * 1e3, (t2 - t0) * 1e3, byts / (t2 - t0) / 1e9, byts / 1e9), flush=True)
