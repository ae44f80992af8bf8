#!/usr/bin/env python
"""All BASELINE.json configs on ONE GPU through the public API (host buffers in, host buffers out).

Not the driver's bench line (that is bench.py); this is the per-config evidence table that goes
to profiles/.  GCUPS = reference cell count / time; `kernel` = device time of the DP kernels only
(wb_stats.kernel_ms), `e2e` = wall time of the API call.  % peak = GCUPS * (FP64 ops per cell) /
measured FP64 issue peak.
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import wildboar_b200 as wb  # noqa: E402
from wildboar_b200 import _shim  # noqa: E402

OPS = {"dtw": 5, "ddtw": 5, "wdtw": 6, "adtw": 7, "lcss": 4, "erp": 6, "edr": 7, "msm": 8, "twe": 10}
NINE = ["dtw", "wdtw", "ddtw", "adtw", "msm", "twe", "erp", "lcss", "edr"]


def rw(n, T, seed):
    return np.cumsum(np.random.default_rng(seed).standard_normal((n, T)), axis=1)


def run(label, fn, peak, metric):
    fn()  # warm-up (allocator pools, module load)
    t0 = time.perf_counter()
    fn()
    dt = time.perf_counter() - t0
    st = wb.last_stats()
    g_k = st["cells"] / (st["kernel_ms"] * 1e-3) / 1e9 if st["kernel_ms"] > 0 else float("nan")
    g_e = st["cells"] / dt / 1e9
    row = dict(config=label, metric=metric, pairs=st["pairs"], cells=st["cells"], engine=st["engine"], launches=st["launches"],
               kernel_ms=round(st["kernel_ms"], 2), e2e_ms=round(dt * 1e3, 2), kernel_gcups=round(g_k, 1), e2e_gcups=round(g_e, 1),
               pct_fp64_peak=round(100 * g_k * OPS[metric] / peak, 1))
    print(json.dumps(row), flush=True)
    return row


def main():
    quick = "--quick" in sys.argv
    wb.set_devices([0])
    peak = _shim.fp64_peak(0)[0] / 1e9
    print(json.dumps({"fp64_peak_g_lane_inst_per_s": peak}), flush=True)
    rows = []
    # cfg1: 200x150 vs 200x150 dtw r=0.1
    x, y = rw(200, 150, 1), rw(200, 150, 2)
    rows.append(run("cfg1 200x150 r=0.1 (full)", lambda: wb.pairwise_distance(x, y, metric="dtw", metric_params={"r": 0.1}), peak, "dtw"))
    # cfg2: 5000x140, all nine metrics, default params: singleton and two-array forms
    X = rw(5000, 140, 1)
    Xc = X.copy()
    for m in NINE:
        rows.append(run("cfg2 5000x140 singleton", lambda m=m: wb.pairwise_distance(X, metric=m), peak, m))
        rows.append(run("cfg2 5000x140 vs copy", lambda m=m: wb.pairwise_distance(X, Xc, metric=m), peak, m))
    # cfg3: 10k x 512, dtw r=0.1 (bench.py is the authoritative number; here a 2500-row block = one GPU of 4)
    x, y = rw(10000, 512, 1), rw(10000, 512, 2)
    nrow = 1250 if quick else 2500
    rows.append(run(f"cfg3 first {nrow} of 10000 x rows vs 10000x512 r=0.1", lambda: wb.pairwise_distance(x[:nrow], y, metric="dtw", metric_params={"r": 0.1}), peak, "dtw"))
    # cfg5: 2000x4096 msm / twe r=0.05 (row block of 250 = one GPU of 8)
    x, y = rw(2000, 4096, 1), rw(2000, 4096, 2)
    nrow = 64 if quick else 250
    for m in ("msm", "twe"):
        rows.append(run(f"cfg5 first {nrow} of 2000 x rows vs 2000x4096 r=0.05", lambda m=m: wb.pairwise_distance(x[:nrow], y, metric=m, metric_params={"r": 0.05}), peak, m))
    # cfg4: argmin k=1 dtw r=0.05, 20k queries x 200k refs x 256: one GPU's share of 8 = 2500 queries
    nq, nr = (256, 50000) if quick else (2500, 200000)
    q, refs = rw(20000, 256, 3)[:nq], rw(200000, 256, 4)[:nr]
    for lb in (False, True):
        def f(lb=lb):
            return wb.argmin_distance(q, refs, k=1, metric="dtw", metric_params={"r": 0.05}, return_distance=True, device_lower_bound=lb)
        f()
        t0 = time.perf_counter(); idx, dist = f(); dt = time.perf_counter() - t0
        st = wb.last_stats()
        nominal = nq * nr * 5756
        row = dict(config=f"cfg4 argmin k=1 {nq} q x {nr} refs x 256 r=0.05 device_lb={lb}", metric="dtw", pairs=nq * nr,
                   cells_evaluated_full_pairs=st["cells"], kernel_ms=round(st["kernel_ms"], 2), e2e_ms=round(dt * 1e3, 2),
                   nominal_gcups=round(nominal / dt / 1e9, 1), launches=st["launches"], checksum=int(idx.sum()))
        print(json.dumps(row), flush=True)
        rows.append(row)
    json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "bench_configs.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
