"""cp.async staging replay on the real plan footprint: width x loader warps x buffers in flight x tile size."""
import ctypes as C, sys, os, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as G
G.build()
from climate_toolbox_b200 import synthetic, _engine as E, _native as N

T = 736
lat, lon = synthetic.grid_labels(0.25)
df = synthetic.weights_table(0.25, 24378)
dev = torch.device("cuda", 0)
x = torch.randn((T, len(lat) * len(lon)), dtype=torch.float32, device=dev)
for budget in (101376, 66 * 1024, 50 * 1024, 33 * 1024):
    plan = E.get_plan(E.GridSpec(lat, lon), df, "popwt", "hierid", smem_budget=budget, cache=False)
    i = plan.info
    nbytes = i["n_pieces"] * 16 * T
    tile = i["max_bundle_cells"] * 33 * 4
    print("budget", budget, "bundles", i["n_bundles"], "staged pieces", i["n_pieces"], "max cells", i["max_bundle_cells"],
          "tile bytes", tile, "staged GB %.3f" % (nbytes / 1e9), flush=True)
    for width, lw, nbuf in itertools.product((16, 8, 4), (4, 8, 16), (2, 3, 4, 6)):
        if ((tile + 127) // 128 * 128) * nbuf > 227 * 1024:
            continue
        ms = []
        for rep in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            N.check(N.lib().ctb_debug_cpasync_bw(plan._h, C.c_void_p(x.data_ptr()), x.shape[1], T, width, lw, nbuf, None))
            e1.record(); torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        print("  width %2d loaders %2d nbuf %d (%.0f KB in flight): %6.0f GB/s" % (width, lw, nbuf, tile * (nbuf - 1) / 1024, nbytes / min(ms) / 1e6), flush=True)
    plan.close()
