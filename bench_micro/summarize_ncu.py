"""ncu report -> the JSON summary committed under profiles/ (run where ncu is installed; no GPU needed).
python bench_micro/summarize_ncu.py gpurun_out/X.ncu-rep profiles/Y.json"""
import csv
import io
import json
import subprocess
import sys
from collections import Counter

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors_srcunit_tex_op_read.sum",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.sum", "smsp__inst_executed_pipe_fp64.sum", "sm__cycles_elapsed.max",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
    "sm__sass_inst_executed_op_ldgsts_cache_bypass.sum",
    "smsp__sass_l1tex_m_xbar2l1tex_read_sectors_mem_global_op_ldgsts_cache_bypass.sum",
]


def run(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main(rep, out):
    rows = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = dict(zip(hdr, vals))
    u = dict(zip(hdr, units))
    res = {"kernel": d.get("Kernel Name"), "report": rep}
    for k in KEEP:
        if k in d and d[k] != "":
            try:
                res[k] = float(d[k].replace(",", ""))
            except ValueError:
                res[k] = d[k]
            if u.get(k):
                res[k + " [unit]"] = u[k]
    src = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "source", "--csv"]))))
    h = src[1]
    ix = {c: i for i, c in enumerate(h)}
    ops, stalls, total = Counter(), Counter(), 0
    wait_count = None
    for r in src[2:]:
        if len(r) < len(h):
            continue
        ie = int(r[ix["Instructions Executed"]] or 0)
        toks = [t for t in r[ix["Source"]].split() if not t.startswith("@")]
        if not toks:
            continue
        if "SYNCS" in toks[0] or "NANOSLEEP" in toks[0]:
            wait_count = ie     # the barrier-wait loops carry replay counts, not issued instructions
            continue
        if toks[0].startswith("BRA") and wait_count is not None and ie == wait_count:
            continue            # ... and so does the branch that closes such a loop
        ops[toks[0].split(".")[0]] += ie
        total += ie
        for c in h:
            if c.startswith("stall_") and "Not Issued" not in c:
                stalls[c] += int(r[ix[c]] or 0)
    res["executed_warp_instructions_by_opcode"] = dict(ops.most_common(25))
    res["executed_warp_instructions_total_without_barrier_waits"] = total
    res["stall_samples"] = dict(stalls.most_common(8))
    json.dump(res, open(out, "w"), indent=1)
    print(json.dumps({k: res[k] for k in list(res)[:12]}, indent=1))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
