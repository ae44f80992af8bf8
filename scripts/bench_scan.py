#!/usr/bin/env python
"""Measurement of SURVEY 8f-4 (second half) on ONE GPU: subsequence search with the metrics whose scan is replayed
(lcss / erp / edr / msm / twe and the generic scaled_<metric> wraps), the UNMODIFIED reference (oracle/_ref) timed beside
it on the box's host cores on a stated, bounded sample of the same workload.  One JSON line per metric.

  python scripts/bench_scan.py [--quick] [--no-ref]
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import wildboar_b200 as wb  # noqa: E402

QUICK = "--quick" in sys.argv
NO_REF = "--no-ref" in sys.argv
NCPU = os.cpu_count() or 1


def timed(fn, reps=2):
    fn()
    best = 1e30
    for _ in range(reps):
        t0 = time.perf_counter()
        out = fn()
        best = min(best, time.perf_counter() - t0)
    return best, out


def main():
    wb.set_devices([0])
    wd = None
    if not NO_REF:
        from oracle import ref
        wd = ref.load()
    n, T, ns, m = (300, 256, 8, 48) if QUICK else (2000, 512, 64, 64)
    Xs = np.cumsum(np.random.default_rng(8).standard_normal((n, T)), axis=-1)
    rng = np.random.default_rng(9)
    shp = [Xs[rng.integers(0, n), o:o + m].copy() for o in rng.integers(0, T - m, ns)]
    cases = [("msm", {"r": 0.1}), ("twe", {"r": 0.1}), ("erp", {"r": 0.1}), ("lcss", {"r": 0.1}), ("edr", {"r": 0.1}),
             ("scaled_msm", {"r": 0.1}), ("scaled_twe", {"r": 0.1}), ("scaled_adtw", {"r": 0.1}), ("scaled_ddtw", {"r": 0.1})]
    for metric, mp in cases:
        t_ss, (d, i) = timed(lambda: wb.pairwise_subsequence_distance(shp, Xs, metric=metric, metric_params=mp, return_index=True))
        st = wb.last_stats()
        row = dict(row="8f-4 pairwise_subsequence_distance (replayed scan)", metric=metric,
                   shape=f"{ns} subsequences x {m} vs {n} samples x {T}, r={mp['r']}", windows=n * (T - m + 1) * ns,
                   cells=st["cells"], e2e_ms=round(t_ss * 1e3, 1), kernel_ms=round(st["kernel_ms"], 1),
                   kernel_gcups=round(st["cells"] / (st["kernel_ms"] * 1e-3) / 1e9, 1), launches=st["launches"], engine=st["engine"],
                   windows_per_s=round(n * (T - m + 1) * ns / t_ss))
        if wd is not None:
            nss, nxs = (4, 48) if QUICK else (8, 128)
            t_ref, (rd_, ri_) = timed(lambda: wd.pairwise_subsequence_distance(shp[:nss], Xs[:nxs], metric=metric, metric_params=mp,
                                                                             return_index=True, n_jobs=NCPU), reps=1)
            row.update(ref_sample=f"{nss} subsequences vs {nxs} samples, n_jobs={NCPU} (early abandoning)", ref_ms=round(t_ref * 1e3, 1),
                       ref_windows_per_s=round(nss * nxs * (T - m + 1) / t_ref), ref_cores=NCPU,
                       ref_bit_equal=bool(np.array_equal(rd_, d[:nxs, :nss]) and np.array_equal(ri_, i[:nxs, :nss])))
        print(json.dumps(row), flush=True)
    # ---- distance_profile / subsequence_match: every window's distance, one list-mode launch per pass ----
    Y = np.stack([Xs[(q + 1) % n, (q * 7) % (T - m):(q * 7) % (T - m) + m] for q in range(n)])
    for metric, mp in (("dtw", {"r": 0.1}), ("scaled_dtw", {"r": 0.1}), ("msm", {"r": 0.1}), ("scaled_twe", {"r": 0.1})):
        t_dp, dp = timed(lambda: wb.distance_profile(Y, Xs, metric=metric, metric_params=mp))
        st = wb.last_stats()
        row = dict(row="8f-4 distance_profile", metric=metric, shape=f"{n} subsequences x {m} paired with {n} samples x {T}, r={mp['r']}",
                   windows=n * (T - m + 1), cells=st["cells"], e2e_ms=round(t_dp * 1e3, 1), kernel_ms=round(st["kernel_ms"], 1),
                   kernel_gcups=round(st["cells"] / (st["kernel_ms"] * 1e-3) / 1e9, 1), launches=st["launches"], engine=st["engine"],
                   windows_per_s=round(n * (T - m + 1) / t_dp))
        if wd is not None:
            nxs = 64 if QUICK else 256
            t_ref, rdp = timed(lambda: wd.distance_profile(Y[:nxs], Xs[:nxs], metric=metric, metric_params=mp, n_jobs=NCPU), reps=1)
            row.update(ref_sample=f"{nxs} pairs, n_jobs={NCPU}", ref_ms=round(t_ref * 1e3, 1), ref_windows_per_s=round(nxs * (T - m + 1) / t_ref),
                       ref_cores=NCPU, ref_bit_equal=bool(np.array_equal(rdp, dp[:nxs])))
        print(json.dumps(row), flush=True)
    for metric, mp, scale in (("dtw", {"r": 0.1}, False), ("msm", {"r": 0.1}, True)):
        t_as, (ai, ad) = timed(lambda: wb.argmin_subsequence_distance(Y, Xs, k=5, metric=metric, metric_params=mp, scale=scale, return_distance=True))
        st = wb.last_stats()
        row = dict(row="8f-4 argmin_subsequence_distance k=5", metric=("scaled_" if scale else "") + metric,
                   shape=f"{n} subsequences x {m} paired with {n} samples x {T}, r={mp['r']}", e2e_ms=round(t_as * 1e3, 1),
                   kernel_ms=round(st["kernel_ms"], 1), launches=st["launches"], windows_per_s=round(n * (T - m + 1) / t_as))
        if wd is not None:
            nxs = 64 if QUICK else 256
            t_ref, (ri, rd) = timed(lambda: wd.argmin_subsequence_distance(Y[:nxs], Xs[:nxs], k=5, metric=metric, metric_params=mp, scale=scale,
                                                                         return_distance=True, n_jobs=NCPU), reps=1)
            row.update(ref_sample=f"{nxs} pairs, n_jobs={NCPU}", ref_ms=round(t_ref * 1e3, 1), ref_windows_per_s=round(nxs * (T - m + 1) / t_ref),
                       ref_cores=NCPU, ref_bit_equal=bool(np.array_equal(ri, ai[:nxs]) and np.array_equal(rd, ad[:nxs])))
        print(json.dumps(row), flush=True)
    thr = float(np.quantile(dp, 0.01))
    t_sm, (mi, md) = timed(lambda: wb.subsequence_match(shp[0], Xs, threshold=thr, metric="scaled_twe", metric_params={"r": 0.1}, return_distance=True))
    row = dict(row="8f-4 subsequence_match", metric="scaled_twe", shape=f"1 subsequence x {m} vs {n} samples x {T}, threshold = 1 % quantile",
               e2e_ms=round(t_sm * 1e3, 1), kernel_ms=round(wb.last_stats()["kernel_ms"], 1), matches=int(sum(0 if a is None else len(a) for a in mi)))
    if wd is not None:
        nxs = 64 if QUICK else 256
        t_ref, (ri, rd) = timed(lambda: wd.subsequence_match(shp[0], Xs[:nxs], threshold=thr, metric="scaled_twe", metric_params={"r": 0.1},
                                                           return_distance=True), reps=1)
        same = all((a is None and b is None) or (a is not None and b is not None and np.array_equal(a, b)) for a, b in zip(ri, mi[:nxs])) and \
            all((a is None and b is None) or (a is not None and b is not None and np.array_equal(a, b)) for a, b in zip(rd, md[:nxs]))
        row.update(ref_sample=f"{nxs} samples (1 core: the reference's match loop is serial)", ref_ms=round(t_ref * 1e3, 1), ref_bit_equal=bool(same))
    print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
