"""Opcode histograms of the shipped kernels (cuobjdump -sass on libctb.so; no GPU needed) ->
profiles/r2_sass_histograms.md"""
import os
import re
import subprocess
import sys
from collections import Counter

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "climate_toolbox_b200", "libctb.so")
WANT = [("agg_stream_kernel<float, IDENTITY, 1 output, 768 threads, 3 stages, no gate, no time index>", "agg_stream_kernelIfLi0ELi1ELi768ELi3ELb0ELb0E"),
        ("agg_stream_kernel<float, IDENTITY, ..., with a time index>", "agg_stream_kernelIfLi0ELi1ELi768ELi3ELb0ELb1E"),
        ("agg_stream_kernel<float, POLY_SEQ (orders 1..4), 4 outputs, 640 threads>", "agg_stream_kernelIfLi16ELi4ELi640ELi3ELb0ELb0E"),
        ("agg_stream_kernel<float, EDD, 2 thresholds, 384 threads>", "agg_stream_kernelIfLi2ELi2ELi384ELi3ELb0ELb0E"),
        ("agg_stream_kernel<float, EDD, 2 thresholds, growing-season gate>", "agg_stream_kernelIfLi2ELi2ELi384ELi3ELb1ELb0E"),
        ("push_rows_kernel", "push_rows_kernel"),
        ("pull_pack_kernel", "pull_pack_kernel"), ("k0_match_kernel", "k0_match_kernel")]
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
funcs = {}
cur = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        funcs[cur] = []
    elif cur and re.match(r"\s+/\*[0-9a-f]{4}\*/", line):
        toks = [t for t in line.split("*/", 1)[1].split() if not t.startswith("@")]
        if toks:
            funcs[cur].append(toks[0].rstrip(";"))
out = ["# SASS of the shipped kernels (`cuobjdump -sass climate_toolbox_b200/libctb.so`, sm_100a)", "",
       "Blackwell/Hopper-era mnemonics to look for: `UBLKCP` = `cp.async.bulk` (TMA 1-D), `SYNCS.*` = mbarrier,",
       "`LDGSTS` = `cp.async`, `IMAD.WIDE.U32` + `VIMNMX3` = the integer float→double widening and its range check,",
       "`MUFU.RCP64H` / `MUFU.RSQ64H` = the fp64 hardware seeds. No tensor-core instructions: the path is a sparse,",
       "bandwidth-bound SpMM (BASELINE.json north_star).", ""]
for title, key in WANT:
    name = next((f for f in funcs if key in f), None)
    if not name:
        out.append("## {}: not found".format(title))
        continue
    ops = funcs[name]
    full = Counter(ops)
    base = Counter(o.split(".")[0] for o in ops)
    out.append("## {}".format(title))
    out.append("`{}` — {} instructions".format(name[:110], len(ops)))
    out.append("")
    out.append("| opcode | count | | opcode | count |")
    out.append("|---|---|---|---|---|")
    items = base.most_common(24)
    for i in range(0, len(items), 2):
        a = items[i]
        b = items[i + 1] if i + 1 < len(items) else ("", "")
        out.append("| {} | {} | | {} | {} |".format(a[0], a[1], b[0], b[1]))
    marks = {k: v for k, v in full.items() if any(s in k for s in ("LDGSTS", "UBLKCP", "SYNCS", "IMAD.WIDE.U32", "VIMNMX3", "MUFU", "F2F", "ARRIVES", "ATOMS", "SHFL.BFLY", "STG.E.EF"))}
    out.append("")
    out.append("marked: " + ", ".join("`{}` × {}".format(k, v) for k, v in sorted(marks.items())))
    out.append("")
open(os.path.join(ROOT, "profiles", "r2_sass_histograms.md"), "w").write("\n".join(out) + "\n")
print("\n".join(out[:40]))
