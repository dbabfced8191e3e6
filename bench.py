#!/usr/bin/env python
"""Benchmark of the grid->region aggregation hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload config2|config3|config4|config1|config5]

One "step" = one pass of the hot path over one synthetic batch of the workload
(config2: 4 x 365 days of 0.25-degree tas [1460][720][1440] f32 -> 24,378 regions, popwt).
N > 1 (torchrun, one rank per GPU): every rank aggregates its OWN batch (another
GCM / scenario / block of years) with the plan replicated -- weak scaling, no data-path
collective.  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (deg, R, T, aggwt, kind, n_in, n_out, description)
    "config1": (1.0, 3000, 365, "areawt", "identity", 1, 1,
                "synthetic 1deg daily tas 365x180x360 f32 -> 3,000 regions, areawt"),
    "config2": (0.25, 24378, 1460, "popwt", "identity", 1, 1,
                "BCSD-shaped 0.25deg daily tas 4x365x720x1440 f32 -> 24,378 hierid regions, popwt"),
    "config3": (0.25, 24378, 1460, "popwt", "poly", 1, 4,
                "fused tas_poly orders 1-4 + aggregation, 0.25deg 1460 days -> 24,378 regions, popwt"),
    "config4": (0.25, 24378, 730, "cropwt", "edd", 2, 2,
                "fused Snyder EDD (2 thresholds) from tasmin/tasmax, 0.25deg 730 days -> 24,378 regions, cropwt"),
    # CMIP5-ensemble scale (21 GCMs x 2 RCPs x 95 years = 3,990 model-years): one step streams a
    # slice of YEARS_PER_STEP model-years through a pool of POOL resident year buffers
    "config5": (0.25, 24378, 365, "popwt", "identity", 1, 1,
                "ensemble streaming: model-years of 365x720x1440 f32 -> 24,378 regions, popwt; "
                "8 model-years per step from a pool of 4 resident buffers (full job: 3,990 model-years)"),
}
YEARS_PER_STEP, POOL, ENSEMBLE_YEARS = 8, 4, 21 * 2 * 95
METRIC = "region-days/sec"


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.sm, self.reasons, self.sm_max = index, False, [], set(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
            nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
            nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake",
        }
        while not self.stop_flag:
            try:
                self.sm.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, nm in names.items():
                    if mask & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.004)

    def result(self):
        self.stop_flag = True
        self.join(1.0)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None,
                "sm_max_mhz": self.sm_max, "reasons": sorted(self.reasons), "samples": len(self.sm)}


# --------------------------------------------------------------------------- #
# reference arm: the restated reference (numpy/pandas oracle) on the host cores
# --------------------------------------------------------------------------- #
_REF = {}


def _ref_worker(args):
    t0, t1 = args
    import oracle
    x, lat, lon, df, aggwt = _REF["x"], _REF["lat"], _REF["lon"], _REF["df"], _REF["aggwt"]
    out = oracle.weighted_aggregate_grid_to_regions(
        x[t0:t1], ("time", "lat", "lon"), lat, lon, df, aggwt, "hierid")[0]
    return float(np.nansum(out))


def _oracle_transform(kind, xs):
    import oracle
    if kind == "identity":
        return [xs[0]]
    if kind == "poly":
        return [(xs[0].astype(np.float64) - 273.15) ** p for p in (1, 2, 3, 4)]
    return [oracle.snyder_edd(xs[0], xs[1], e) for e in (283.15, 303.15)]


def cpu_sample(workload, days, procs):
    """Time the oracle on `days` days of the workload with `procs` forked workers."""
    import multiprocessing as mp
    from climate_toolbox_b200 import synthetic
    deg, R, T, aggwt, kind, n_in, n_out, _ = WORKLOADS[workload]
    lat, lon = synthetic.grid_labels(deg)
    df = synthetic.weights_table(deg, R)
    rng = np.random.default_rng(7)
    x = (288.0 + 10.0 * rng.standard_normal((days, len(lat), len(lon)), dtype=np.float32))
    _REF.update(x=x, lat=lat, lon=lon, df=df, aggwt=aggwt)
    if kind != "identity":
        raise SystemExit("reference arm runs the headline identity workloads (config1/config2)")
    per = max(1, days // procs)
    chunks = [(a, min(days, a + per)) for a in range(0, days, per)]

    def one_pass():
        t = time.perf_counter()
        if procs == 1:
            for c in chunks:
                _ref_worker(c)
        else:
            with mp.get_context("fork").Pool(procs) as pool:
                pool.map(_ref_worker, chunks)
        return time.perf_counter() - t

    return one_pass, R, days


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    procs = max(1, min(cores, 32))
    deg, R, T, aggwt, kind, n_in, n_out, desc = WORKLOADS[args.workload]
    days = procs * (8 if deg < 1 else 64)
    one_pass, R, days = cpu_sample(args.workload, days, procs)
    for _ in range(args.warmup):
        one_pass()
    ts = [one_pass() for _ in range(args.steps)]
    sec = float(np.mean(ts))
    v = R * days / sec
    sample = "{} days of the workload per step ({} forked workers x {} days, time-sharded)".format(
        days, procs, max(1, days // procs))
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "region-days/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": desc, "impl_note": "restated reference (oracle/oracle.py, numpy/pandas): "
                   "xarray is not installable in this image, so the unmodified reference cannot run"},
        "cpu_baseline": {"value": v, "unit": "region-days/s", "cores": procs, "kind": "port",
                         "sample": sample},
        "e2e": {"value": v, "unit": "region-days/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------- #
# our arm
# --------------------------------------------------------------------------- #
def run_ours(args):
    import torch
    import torch.distributed as dist

    import __graft_entry__ as G
    G.build()
    from climate_toolbox_b200 import Dataset, DataArray, synthetic
    from climate_toolbox_b200 import _engine as E
    from climate_toolbox_b200 import _native as N
    from climate_toolbox_b200.aggregations.aggregations import weighted_aggregate_grid_to_regions
    from climate_toolbox_b200.transformations.transformations import snyder_edd

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus {} needs torchrun with {} ranks".format(args.gpus, args.gpus))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:   # ranks share the host cores for the e2e packing
        os.environ.setdefault("CTB_PACK_THREADS", str(max(1, (3 * (os.cpu_count() or 1)) // (4 * world))))
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    deg, R, T, aggwt, kind, n_in, n_out, desc = WORKLOADS[args.workload]
    if args.days:
        T = args.days
    lat, lon = synthetic.grid_labels(deg)
    df = synthetic.weights_table(deg, R)
    g = torch.Generator(device=dev).manual_seed(7 + rank)
    shape = (T, len(lat), len(lon))
    tas = 288.0 + 10.0 * torch.randn(shape, generator=g, device=dev, dtype=torch.float32)
    xs = [tas]
    if n_in == 2:
        lo = tas - (3.0 * torch.randn(shape, generator=g, device=dev, dtype=torch.float32)).abs()
        hi = tas + (3.0 * torch.randn(shape, generator=g, device=dev, dtype=torch.float32)).abs()
        xs = [lo, hi]
        del tas
    params = {"identity": (), "poly": (273.15, 1, 2, 3, 4), "edd": (283.15, 303.15)}[kind]

    grid = E.GridSpec(lat, lon)
    t_plan = time.perf_counter()
    plan = E.get_plan(grid, df, aggwt, "hierid", stage_bytes=4 * n_in, device=dev,
                      smem_budget=args.smem_budget)
    torch.cuda.synchronize()
    plan_ms = (time.perf_counter() - t_plan) * 1e3
    info = plan.info
    ncell = len(lat) * len(lon)
    x2 = [x.view(T, ncell) for x in xs]
    out = torch.empty((n_out, plan.R, T), dtype=torch.float64, device=dev)
    ws_bytes = N.lib().ctb_aggregate_workspace_bytes(plan._h, T, n_out)
    ws = torch.empty(max(1, ws_bytes // 8), dtype=torch.float64, device=dev)

    streaming = args.workload == "config5"
    years = YEARS_PER_STEP if streaming else 1
    if streaming:   # more resident model-years (the first is `tas`), outputs checksummed and discarded
        pool = [x2[0]] + [(288.0 + 10.0 * torch.randn(shape, generator=g, device=dev, dtype=torch.float32)
                           ).view(T, ncell) for _ in range(POOL - 1)]
        chk = torch.zeros((), dtype=torch.float64, device=dev)

    def step():
        if streaming:
            for y in range(YEARS_PER_STEP):
                E.aggregate_device(plan, pool[y % POOL], None, N.LAYOUT_TIME_MAJOR, ncell, None, T, kind,
                                   params, n_out, args.variant, out=out, workspace=ws)
                chk.add_(out.sum())
            return
        E.aggregate_device(plan, x2[0], x2[1] if n_in == 2 else None, N.LAYOUT_TIME_MAJOR, ncell,
                           None, T, kind, params, n_out, args.variant, out=out, workspace=ws)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    n0 = E.launch_count()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
           for _ in range(args.steps)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for a, b in evs:
        a.record()
        step()
        b.record()
    e1.record()
    barrier()
    launches = E.launch_count() - n0
    total_ms = e0.elapsed_time(e1)
    kern_ms = float(np.mean([a.elapsed_time(b) for a, b in evs]))
    clocks = sampler.result()
    tmax = torch.tensor([total_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    total_ms = float(tmax.item())
    ms_per_step = total_ms / args.steps
    value = world * plan.R * T * years / (ms_per_step * 1e-3)
    checksum = float(torch.nansum(out).item())

    # ---- roofline of the dominant kernel (the fused staged kernel = the whole step) ----
    balg = plan.algorithmic_bytes(T, n_in, 4, n_out)   # per launch (config5: one model-year)
    if streaming:   # per launch, including the checksum reductions between launches
        kern_ms = total_ms / args.steps / years
    peak, peak_src = _peaks()
    achieved = balg / (kern_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get(args.workload)
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "kernel": "agg_fused_kernel", "algorithmic_bytes_per_launch": balg,
                "bytes_per_region_day": balg / (plan.R * T), "launch_ms": kern_ms,
                "frac_of_8TBs_nominal": achieved / 8000.0}

    # ---- end to end through the public API with HOST (pinned) buffers ----
    e2e = None
    if not args.no_e2e and not streaming:
        # every rank pins its own host copy of the batch: keep the node's total under 40 % of the
        # free host memory (8 ranks x 6 GB would not fit a small host) by shortening the e2e batch
        T_full = T
        try:
            import psutil
            avail = psutil.virtual_memory().available
        except Exception:
            avail = 32 << 30
        need = world * len(xs) * T * ncell * 4
        if need > 0.4 * avail:
            T = max(64, int(T * 0.4 * avail / need) // 32 * 32)
        if world > 1:
            tmin = torch.tensor([T], device=dev, dtype=torch.int64)
            dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
            T = int(tmin.item())
        shape = (T, len(lat), len(lon))
        host = [torch.empty(shape, dtype=torch.float32, pin_memory=True) for _ in xs]
        for h, x in zip(host, xs):
            h.copy_(x[:T])
        torch.cuda.synchronize()
        coords = {"time": np.arange(T), "lat": lat, "lon": lon}
        dims = ("time", "lat", "lon")
        if kind == "identity":
            ds = Dataset({"tas": (dims, host[0].numpy())}, coords=coords)
            names = "tas"
        elif kind == "poly":
            from climate_toolbox_b200._xr import Deferred, Variable
            src = Variable(dims, host[0].numpy())
            ds = Dataset(coords=coords)
            names = ["p1", "p2", "p3", "p4"]
            for p, nme in zip((1, 2, 3, 4), names):
                ds._vars[nme] = Variable(dims, None, None, None, Deferred("poly", (273.15, float(p)), (src,)))
        else:
            tn = DataArray(host[0].numpy(), dims=dims, coords=coords, attrs={"units": "K"})
            tx = DataArray(host[1].numpy(), dims=dims, coords=coords, attrs={"units": "K"})
            ds = Dataset(coords=coords)
            ds["edd10"] = snyder_edd(tn, tx, 283.15, check=False)
            ds["edd30"] = snyder_edd(tn, tx, 303.15, check=False)
            names = ["edd10", "edd30"]

        def e2e_step():
            r = weighted_aggregate_grid_to_regions(ds, names, aggwt, "hierid", weights=df,
                                                   smem_budget=args.smem_budget)
            return r

        e2e_steps = max(2, min(args.steps, 10))   # ~0.1 s each; the shared host's spikes average out a little
        # warm-up: pinned staging/result blocks come from pools, and the host needs a moment after the
        # 6 GB pinned source was allocated and filled (the first second of calls runs 3-5x slower) --
        # at least 3 calls and 2.5 s, whichever is longer
        t_w, n_w = time.perf_counter(), 0
        while n_w < 3 or (time.perf_counter() - t_w < 2.5 and n_w < 30):
            r = e2e_step()
            n_w += 1
        barrier()
        E.TRANSFER_BYTES.update(h2d=0, d2h=0)
        t0 = time.perf_counter()
        step_ms = []
        for _ in range(e2e_steps):
            ts = time.perf_counter()
            r = e2e_step()          # returns host arrays: the call is synchronous
            step_ms.append(round((time.perf_counter() - ts) * 1e3, 1))
        torch.cuda.synchronize()
        barrier()
        dt = (time.perf_counter() - t0) / e2e_steps
        h2d_step = E.TRANSFER_BYTES["h2d"] // e2e_steps
        d2h_step = E.TRANSFER_BYTES["d2h"] // e2e_steps
        tt = torch.tensor([dt], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        first = names if isinstance(names, str) else names[0]
        e2e_check = float(np.nansum(r[first].values))
        e2e = {"value": world * plan.R * T / dt, "unit": "region-days/s",
               "h2d_bytes_per_step": int(h2d_step), "d2h_bytes_per_step": int(d2h_step),
               "host_input_bytes_per_step": int(sum(h.numel() * 4 for h in host)),
               "ms_per_step": dt * 1e3, "days_per_step": T, "step_ms": step_ms, "warmup_calls": n_w,
               "pinned_result_blocks_allocated": int(sum(E._RESULT_OUT.values())),
               "steps": e2e_steps, "api": "weighted_aggregate_grid_to_regions(ds[numpy over pinned host "
               "memory], ...) -> Dataset[numpy]: host packing of the referenced gridcells + pinned chunked "
               "H2D + fused kernel + pinned D2H",
               "checksum": e2e_check}
        del host
        T = T_full

    # ---- optional: the final gather of region x time outputs (north_star) ----
    gather = None
    if world > 1:
        from climate_toolbox_b200.parallel import all_gather_time  # noqa: F401
        bufs = [torch.empty_like(out) for _ in range(world)]
        dist.all_gather(bufs, out)
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        dist.all_gather(bufs, out)
        g1.record()
        barrier()
        gather = {"ms": g0.elapsed_time(g1), "bytes_per_rank": int(out.numel() * 8),
                  "note": "NCCL all_gather of every rank's [n_out][R][T] f64 block; not in `value`"}

    # ---- CPU baseline: the oracle on a bounded sample, rank 0, N = 1 only ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu and kind == "identity":
        days = 365 if deg < 1 else 365
        one_pass, R_, days = cpu_sample(args.workload, days, 1)
        sec = one_pass()
        cpu = {"value": plan.R * days / sec, "unit": "region-days/s", "cores": 1, "kind": "port",
               "sample": "{} days of the workload, single process (the reference is single-threaded "
                         "eager numpy), {:.1f} s; host has {} cores".format(days, sec, os.cpu_count())}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "region-days/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": desc, "T": T, "grid": [len(lat), len(lon)], "regions": plan.R,
                       "input_dtype": "f32", "accumulate": "f64", "n_in": n_in, "n_out": n_out,
                       "l2": "inputs larger than L2 ({:.2f} GB per step)".format(n_in * T * ncell * 4 / 1e9),
                       "U": info["n_cells_distinct"], "nnz": info["nnz"], "bundles": info["n_bundles"],
                       "staged_pieces": info["n_pieces"], "distinct_pieces": info["n_pieces_distinct"],
                       "plan_build_ms": plan_ms, "variant": args.variant,
                       "per_rank": "own batch, plan replicated, outputs stay sharded"},
            **({"ensemble": {"model_years_per_step": years, "pool_buffers": POOL,
                             "full_job_model_years": ENSEMBLE_YEARS,
                             "full_job_seconds_extrapolated": ENSEMBLE_YEARS / (years * world) * ms_per_step * 1e-3,
                             "note": "outputs are summed into a checksum and discarded; e2e is not measured "
                                     "for this workload (6 TB of input do not exist on the host)"}}
               if streaming else {}),
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
            "clocks": clocks, "checksum": checksum,
        }
        if gather:
            line["gather"] = gather
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="config2", choices=sorted(WORKLOADS))
    ap.add_argument("--days", type=int, default=0, help="override T (debug)")
    ap.add_argument("--variant", type=int, default=0)
    ap.add_argument("--smem-budget", type=int, default=0)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
