#!/usr/bin/env python
"""Benchmark of the elastic-distance hot path (BASELINE.json metric: pairwise elastic-distance
GCUPS = DP cells / s / 1e9).

Workload (config.workload): BASELINE configs[2] -- pairwise_distance(metric="dtw", r=0.1) on
10 000 x 512 vs 10 000 x 512 float64 random walks (seeds 1 / 2), the x rows sharded over the
ranks in contiguous blocks (no collective on the data path; y is replicated).  One "step" = one
full pass over the rank's row block.

    python bench.py [--gpus N] [--steps K] [--warmup W]          # our CUDA path
    python bench.py --impl reference ...                          # reference CPU arm (bounded sample)
    torchrun --nproc-per-node N ... bench.py --gpus N ...         # one rank per GPU

JSON line keys follow the driver's contract; additions: `roofline` (FP64-ALU bound, measured
peak from the in-run DADD issue microbenchmark), `cpu_baseline` (the reference's Cython build
from oracle/_ref on the host cores, or the C oracle port when that build is absent).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = {
    "cfg3": dict(metric="dtw", r=0.1, nx=10000, ny=10000, T=512,
                 desc="pairwise_distance dtw r=0.1, 10000x512 vs 10000x512 float64 random walks (BASELINE configs[2])"),
    "cfg1": dict(metric="dtw", r=0.1, nx=200, ny=200, T=150, desc="pairwise dtw r=0.1 200x150 vs 200x150 (configs[0])"),
    "cfg2": dict(metric="dtw", r=1.0, nx=5000, ny=5000, T=140, desc="pairwise dtw r=1.0 5000x140 vs copy (configs[1], one metric)"),
}
# per-metric variants of the cfg2 / cfg5 shapes (profiling and per-config evidence; not the driver's bench line)
for _m in ("wdtw", "ddtw", "adtw", "msm", "twe", "erp", "lcss", "edr"):
    WORKLOAD[f"cfg2_{_m}"] = dict(metric=_m, r=1.0, nx=5000, ny=5000, T=140,
                                  desc=f"pairwise {_m} r=1.0 5000x140 vs copy (configs[1], one metric)")
for _m in ("msm", "twe", "dtw"):
    WORKLOAD[f"cfg5_{_m}"] = dict(metric=_m, r=0.05, nx=2000, ny=2000, T=4096,
                                  desc=f"pairwise {_m} r=0.05 2000x4096 vs 2000x4096 (configs[4])")
FP64_OPS_PER_CELL = {"dtw": 5, "ddtw": 5, "wdtw": 6, "adtw": 7, "lcss": 4, "erp": 6, "edr": 7, "msm": 8, "twe": 10}


def random_walks(n, T, seed):
    return np.cumsum(np.random.default_rng(seed).standard_normal((n, T)), axis=1)


def cells_per_pair(T, r, metric="dtw"):
    """Reference cell count of one equal-length pair; ddtw runs on T-2 points with R from T."""
    R = max(int(np.floor(T * r)), 1)
    if metric in ("ddtw", "wddtw"):
        T = T - 2
    R = min(R, T)
    return T * (2 * R - 1) - R * (R - 1)


from wildboar_b200.sharding import aggregate_throughput, max_over_ranks, row_block  # noqa: E402


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln)

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [t.strip() for t in ln.split(",")]
            if len(f) < 9 or f[0] != str(self.idx):
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "power_w_max": max(pw), "samples": len(sm),
                "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------
# reference CPU arm / cpu_baseline
# ---------------------------------------------------------------------------------------------
def cpu_reference_callable():
    """(kind, fn(x, y, r, n_jobs) -> matrix).  'reference' = wildboar's own Cython build from
    oracle/_ref; 'port' = the C oracle (oracle/elastic_oracle.c)."""
    from oracle import ref
    wd = ref.load()
    if wd is not None:
        def fn(x, y, r, n_jobs, metric="dtw"):
            return wd.pairwise_distance(x, y, metric=metric, metric_params={"r": r}, n_jobs=n_jobs)
        return "reference", fn
    from oracle import oracle as O

    def fn(x, y, r, n_jobs, metric="dtw"):
        return O.pairwise(metric, x, y, r=r, n_jobs=n_jobs if n_jobs > 0 else 0)
    return "port", fn


def time_cpu_sample(wl, target_s=12.0, steps=1):
    """Time the CPU reference on a bounded row sample of the workload; returns dict."""
    kind, fn = cpu_reference_callable()
    cores = os.cpu_count() or 1
    x = random_walks(wl["nx"], wl["T"], 1)
    y = random_walks(wl["ny"], wl["T"], 2)
    cpp = cells_per_pair(wl["T"], wl["r"], wl["metric"])
    ny_s = min(wl["ny"], 1024)
    # warm-up / calibration (joblib thread start-up, page-in)
    nx_p = min(wl["nx"], max(cores, 8))
    t0 = time.perf_counter(); fn(x[:nx_p], y[:ny_s], wl["r"], cores, wl["metric"]); fn(x[:nx_p], y[:ny_s], wl["r"], cores, wl["metric"])
    dt = (time.perf_counter() - t0) / 2
    rate = nx_p * ny_s * cpp / max(dt, 1e-6)
    nx_s = int(min(wl["nx"], max(cores, target_s * rate / (ny_s * cpp))))
    nx_s = max(cores, (nx_s // cores) * cores)
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        fn(x[:nx_s], y[:ny_s], wl["r"], cores, wl["metric"])
        times.append(time.perf_counter() - t0)
    cells = nx_s * ny_s * cpp
    return {"kind": kind, "cores": cores, "times": times, "cells_per_step": cells,
            "value": cells / min(times) / 1e9, "unit": "GCUPS",
            "sample": f"first {nx_s} x rows vs first {ny_s} y rows of the workload, n_jobs={cores}, best of {steps}"}


def run_reference_arm(args, wl, wl_name):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(args.steps, 1)
    # warm-up steps are folded into time_cpu_sample's calibration calls
    res = time_cpu_sample(wl, target_s=max(3.0, min(20.0, 60.0 / steps)), steps=steps)
    tot = sum(res["times"])
    value = res["cells_per_step"] * steps / tot / 1e9
    line = {
        "impl": "reference", "metric": "pairwise_elastic_distance_gcups", "value": value, "unit": "GCUPS",
        "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup, "ms_per_step": tot / steps * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["desc"], "workload_id": wl_name, "sample": res["sample"]},
        "cpu_baseline": {"value": value, "unit": "GCUPS", "cores": res["cores"], "kind": res["kind"], "sample": res["sample"]},
        "e2e": {"value": value, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=sorted(WORKLOAD))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--precision", default="fp64", choices=["fp64", "fp32"],
                    help="fp64 = bit-exact mode (the driver's bench line); fp32 = the optional fp32 mode (extra evidence only)")
    ap.add_argument("--e2e-steps", type=int, default=None)
    ap.add_argument("--profile-rows", type=int, default=0,
                    help="PROFILING ONLY (ncu): shrink the x row count so ~40 kernel replays stay short; "
                         "the JSON line is then marked and is not a bench value")
    args = ap.parse_args()
    wl = dict(WORKLOAD[args.workload])
    if args.profile_rows > 0:
        wl["nx"] = min(wl["nx"], args.profile_rows)
        wl["desc"] += f" [PROFILING RUN: first {wl['nx']} x rows only]"
    if args.impl == "reference":
        run_reference_arm(args, wl, args.workload)
        return
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    from wildboar_b200 import _build, _shim
    import wildboar_b200 as wb

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world == 1 and args.gpus > 1:
        raise SystemExit("launch with torchrun --nproc-per-node N for --gpus N > 1 (one rank per GPU)")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the CUDA path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    if rank == 0:
        _build.build()
    if world > 1:
        dist.barrier()
    _shim.lib()

    metric, r, T = wl["metric"], wl["r"], wl["T"]
    mid = _shim.METRIC_IDS[metric]
    params = wb.check_metric(metric)(r=r)._params()
    params.precision = 1 if args.precision == "fp32" else 0
    wb.set_precision(args.precision)
    lo, hi = row_block(wl["nx"], world, rank)
    x_h = random_walks(wl["nx"], T, 1)[lo:hi].copy()
    y_h = random_walks(wl["ny"], T, 2)
    nx, ny = hi - lo, wl["ny"]
    cpp = cells_per_pair(T, r, metric)
    cells_rank = nx * ny * cpp
    cells_total = wl["nx"] * ny * cpp

    # FP64 issue-rate microbenchmark (roofline denominator), this device, before the timed region
    peak_inst, sm_mhz_est = _shim.fp64_peak(0)
    peak_mix, _ = _shim.fp64_peak(1)

    x_d = torch.from_numpy(x_h).to(dev)
    y_d = torch.from_numpy(y_h).to(dev)
    out_d = torch.empty((nx, ny), dtype=torch.float64, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    stream = torch.cuda.current_stream()

    def step():
        flush.fill_(1)
        return _shim.pairwise_dev(mid, params, x_d.data_ptr(), nx, T, y_d.data_ptr(), ny, T, out_d.data_ptr(),
                                  stream.cuda_stream, want_stats=False)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    barrier()
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1)
    # kernel-only duration (CUDA events inside the library, same stream), one extra untimed pass
    st = _shim.pairwise_dev(mid, params, x_d.data_ptr(), nx, T, y_d.data_ptr(), ny, T, out_d.data_ptr(), stream.cuda_stream)
    kernel_ms = st["kernel_ms"]
    checksum = float(out_d[:: max(nx // 64, 1), ::97].sum().item())

    # end to end through the public API: host numpy in, host numpy out, every step
    e2e_steps = args.e2e_steps if args.e2e_steps is not None else args.steps
    wb.set_devices([local])
    wb.pairwise_distance(x_h[: max(nx // 8, 1)], y_h, metric=metric, metric_params={"r": r})  # warm-up
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        res = wb.pairwise_distance(x_h, y_h, metric=metric, metric_params={"r": r})
    barrier()
    e2e_s = time.perf_counter() - t0
    e2e_stats = wb.last_stats()
    assert res.shape == (nx, ny)

    ms, e2e_ms, kernel_ms = max_over_ranks([ms, e2e_s * 1e3, kernel_ms], device=dev)

    if rank == 0:
        value = aggregate_throughput(cells_total, args.steps, ms) / 1e9
        e2e_value = aggregate_throughput(cells_total, e2e_steps, e2e_ms) / 1e9
        ops = FP64_OPS_PER_CELL[metric]
        achieved = cells_rank * ops / (kernel_ms * 1e-3) / 1e9  # G FP64-pipe lane-instructions / s, this GPU
        nominal = 148 * 64 * 1.965  # G lane-inst/s at max boost
        if args.precision == "fp32":
            # optional fp32 mode: 3 instructions per dtw cell (FADD, FMNMX3, FFMA) against the nominal
            # FP32 issue rate 148 SMs x 128 lanes x 1.965 GHz (no in-run microbenchmark for this mode)
            ops = 3 if metric in ("dtw", "ddtw") else ops
            achieved = cells_rank * ops / (kernel_ms * 1e-3) / 1e9
            nominal = 148 * 128 * 1.965
            peak_inst = nominal * 1e9
        line = {
            "metric": "pairwise_elastic_distance_gcups", "value": value, "unit": "GCUPS", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64" if args.precision == "fp64" else "f32", "data": "synthetic",
            "config": {"workload": wl["desc"], "workload_id": args.workload, "cells_per_pair": cpp,
                       "pairs": wl["nx"] * ny, "sharding": f"x rows in {world} contiguous blocks, y replicated, no collective",
                       "l2": "256 MB buffer written between timed iterations (L2 flush)", "mode": "fp64 bit-exact (-fmad=false)" if args.precision == "fp64" else "optional fp32 mode (<= 1e-4 relative)"},
            "e2e": {"value": e2e_value, "unit": "GCUPS", "h2d_bytes_per_step": int((nx + ny) * T * 8),
                    "d2h_bytes_per_step": int(nx * ny * 8), "steps": e2e_steps, "ms_per_step": e2e_ms / e2e_steps,
                    "api": "wildboar_b200.pairwise_distance(numpy, numpy) -> numpy", "device_ms_last_call": e2e_stats["total_ms"]},
            "gpu_launches": int(args.steps * st["launches"]),
            "clocks": clocks,
            "roofline": {
                "bound": "fp64_alu" if args.precision == "fp64" else "fp32_alu", "kernel": _kernel_label(metric, st), "achieved": achieved, "peak": peak_inst / 1e9,
                "unit": "G FP64-pipe lane-inst/s", "frac": achieved / (peak_inst / 1e9),
                "peak_source": ("measured in this run: wb_cuda_fp64_peak(mix=0), DADD issue rate, all SMs" if args.precision == "fp64"
                                else "nominal FP32 issue rate 148 x 128 lanes x 1.965 GHz"),
                "ops_per_cell": ops, "kernel_ms": kernel_ms, "kernel_gcups": cells_rank / (kernel_ms * 1e-3) / 1e9,
                "nominal_peak": nominal, "frac_of_nominal": achieved / nominal,
                "dtw_mix_peak": peak_mix / 1e9, "frac_of_dtw_mix_peak": achieved / (peak_mix / 1e9),
                "dtw_mix_peak_note": "register-only loop of the same instruction mix (3 FP64 arith + 2x(DSETP+2 FSEL)): practical issue ceiling, profiles/r01_issue_model.md",
                # dram__bytes_read.sum + dram__bytes_write.sum of ONE full-size launch of this kernel, from the ncu capture of
                # this very command line (profiles/r01e_traffic_full_cfg3.csv, profiles/r01d_l2_persist.md); ncu cannot run inside
                # the timed bench, so the number is recorded here for the workload / device count it was measured on
                "traffic": (16.68e9 if (args.workload == "cfg3" and world == 1 and args.precision == "fp64" and args.profile_rows == 0) else None),
                "traffic_unit": "bytes per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum)",
                "hbm": {"algorithmic_bytes": int((nx + ny) * T * 8 + nx * ny * 8),
                        "achieved_gbs": ((nx + ny) * T * 8 + nx * ny * 8) / (kernel_ms * 1e-3) / 1e9,
                        "peak_gbs": _measured_hbm()},
            },
            "checksum": checksum,
        }
        if world == 1 and not args.no_cpu_baseline:
            try:
                cb = time_cpu_sample(wl, target_s=12.0, steps=1)
                line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
            except Exception as e:  # the baseline must never take the GPU number down with it
                line["cpu_baseline"] = {"value": None, "unit": "GCUPS", "cores": os.cpu_count(), "kind": "unavailable", "sample": repr(e)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def _kernel_label(metric, st):
    if st.get("engine") != 2:
        return f"k_rowscan<{metric}>"
    where = "L2-resident global boundary buffers" if st.get("strip_gring") else "shared-memory boundary buffers"
    return (f"k_strip<{metric}, W={st.get('strip_w')}, NR={st.get('strip_nr')}, {st.get('strip_warps')} warps/CTA> "
            f"(thread per pair, {where})")


def _measured_hbm():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        return 6650.0  # fallback stated in B200_PROFILING.md


if __name__ == "__main__":
    main()
