"""Golden vectors for SURVEY 8f-4, second half, generated from the UNMODIFIED reference (oracle/_ref).

    python tests/golden/make_golden_scan.py      # needs oracle/_ref (oracle/build_ref.sh)

Writes tests/golden/scan_golden.npz:
  sc|...   pairwise / paired_subsequence_distance for lcss / erp / edr / msm / twe (Lcss ... Erp SubsequenceMetric,
           _elastic.pyx:2616-3124) -- the metrics whose early abandoning decides which window wins
  sw|...   the same calls for the generic scaled metrics scaled_<metric> = ScaledSubsequenceMetricWrap(Metric)
           (_cdistance.pyx:470-551) over adtw, wdtw, ddtw, wddtw, lcss, erp, edr, msm, twe
  sm|...   subsequence_match / paired_subsequence_match (_distance.py:732-1080) and distance_profile (:1477-1600) for
           every elastic subsequence metric, unscaled and scaled; jagged results padded with -1 / NaN
  dd|...   distance_profile with dilation / padding (_dilated_distance_profile, _cdistance.pyx:804-935, 1728-1862)
  as|...   argmin_subsequence_distance (_distance.py:1636-1790), k in {1, 4}, scale in {False, True}
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import ref  # noqa: E402

SC_CASES = [("lcss", {}), ("lcss", {"r": 0.2, "epsilon": 0.5}), ("erp", {}), ("erp", {"r": 0.1, "g": 0.4}),
            ("edr", {}), ("edr", {"r": 0.25}), ("edr", {"r": 0.3, "epsilon": 0.8}), ("msm", {}), ("msm", {"r": 0.15, "c": 0.3}),
            ("twe", {}), ("twe", {"r": 0.2, "penalty": 0.5, "stiffness": 0.05})]
SW_CASES = [("adtw", {"r": 0.2, "p": 0.5}), ("wdtw", {"r": 0.3, "g": 0.1}), ("ddtw", {"r": 0.2}), ("wddtw", {"r": 0.5, "g": 0.2}),
            ("lcss", {"r": 0.2, "epsilon": 0.5}), ("erp", {"r": 0.1, "g": 0.4}), ("edr", {"r": 0.25}),
            ("edr", {"r": 0.3, "epsilon": 0.8}), ("msm", {"r": 0.15, "c": 0.3}), ("twe", {"r": 0.2, "penalty": 0.5, "stiffness": 0.05}),
            ("msm", {}), ("lcss", {})]
LENGTHS = (12, 30, 5, 3, 72, 1, 2, 41)


SM_CASES = [("dtw", {"r": 0.1}), ("wdtw", {"r": 0.3, "g": 0.1}), ("adtw", {"r": 0.2, "p": 0.5}), ("ddtw", {"r": 0.2}),
            ("wddtw", {"r": 0.5, "g": 0.2}), ("lcss", {"r": 0.2, "epsilon": 0.5}), ("erp", {"r": 0.1, "g": 0.4}), ("edr", {"r": 0.25}),
            ("edr", {"r": 0.3, "epsilon": 0.8}), ("msm", {"r": 0.15, "c": 0.3}), ("twe", {"r": 0.2, "penalty": 0.5, "stiffness": 0.05})]
SM_SUBS = (0, 1, 3)   # lengths 12, 30, 3
AS_CASES = SM_CASES


DD_CASES = SM_CASES
DD_GEOMETRY = ((2, 0), (1, 3), (2, "same"), (3, 5), (5, 0))


def pad(lst, fill, dtype):
    """Jagged per-sample results (None = no match) -> (n_samples, K) padded."""
    lst = [np.array([], dtype=dtype) if a is None else np.atleast_1d(a) for a in lst]
    K = max(1, max(len(a) for a in lst))
    out = np.full((len(lst), K), fill, dtype=dtype)
    for i, a in enumerate(lst):
        out[i, :len(a)] = a
    return out


def matches(wd, out, X, subs):
    n = X.shape[0]
    for prefix in ("", "scaled_"):
        for ci, (metric, mp) in enumerate(SM_CASES):
            name = prefix + metric
            for k in SM_SUBS:
                s = subs[k]
                key = f"sm|{name}|{ci}|{k}"
                full_i, full_d = wd.subsequence_match(s, X, max_matches=10**9, metric=name, metric_params=mp, return_distance=True)
                dists = np.concatenate([d for d in full_d if d is not None])
                for ti, thr in enumerate((float(np.median(dists)), float(np.quantile(dists, 0.1)))):
                    i_, d_ = wd.subsequence_match(s, X, threshold=thr, metric=name, metric_params=mp, return_distance=True)
                    out[f"{key}|thr{ti}"] = np.array(thr)
                    out[f"{key}|thr{ti}|idx"], out[f"{key}|thr{ti}|dist"] = pad(i_, -1, np.int64), pad(d_, np.nan, float)
                # API-level forms: the 10 best, "auto", exclusion zone + max_matches, per-sample thresholds
                i_, d_ = wd.subsequence_match(s, X, metric=name, metric_params=mp, return_distance=True)
                out[f"{key}|top|idx"], out[f"{key}|top|dist"] = pad(i_, -1, np.int64), pad(d_, np.nan, float)
                i_, d_ = wd.subsequence_match(s, X, threshold="auto", metric=name, metric_params=mp, return_distance=True)
                out[f"{key}|auto|idx"], out[f"{key}|auto|dist"] = pad(i_, -1, np.int64), pad(d_, np.nan, float)
                i_, d_ = wd.subsequence_match(s, X, threshold=float(np.median(dists)), exclude=0.5, max_matches=4, metric=name,
                                              metric_params=mp, return_distance=True)
                out[f"{key}|excl|idx"], out[f"{key}|excl|dist"] = pad(i_, -1, np.int64), pad(d_, np.nan, float)
            # paired: subsequence q (mixed lengths) against sample q
            paired = [subs[SM_SUBS[q % len(SM_SUBS)]] for q in range(n)]
            i_, d_ = wd.paired_subsequence_match(paired, X, metric=name, metric_params=mp, return_distance=True, max_matches=5)
            out[f"sm|{name}|{ci}|paired|idx"], out[f"sm|{name}|{ci}|paired|dist"] = pad(i_, -1, np.int64), pad(d_, np.nan, float)
            # distance profile: subsequence q = a window of sample (q + 1) % n, length 15
            Y = np.stack([X[(q + 1) % n, 7 + q:22 + q] for q in range(n)])
            out[f"sm|{name}|{ci}|profile"] = wd.distance_profile(Y, X, metric=name, metric_params=mp)


def main():
    wd = ref.load()
    if wd is None:
        raise SystemExit("oracle/_ref is not built; run oracle/build_ref.sh first")
    out = {}
    rng = np.random.default_rng(20261101)
    X = np.cumsum(rng.standard_normal((10, 72)), axis=1)
    X[6, 20:40] = X[6, 20]                      # a constant stretch (zero-variance windows, ties)
    X[8, 40:52] = X[8, 5:17]                    # a repeated motif: the first window must win
    subs = [np.cumsum(rng.standard_normal(m)) for m in LENGTHS]
    subs[0] = X[8, 5:17].copy()
    out["X"] = X
    for k, s in enumerate(subs):
        out[f"s{k}"] = s
    for tag, cases, prefix in (("sc", SC_CASES, ""), ("sw", SW_CASES, "scaled_")):
        for ci, (metric, mp) in enumerate(cases):
            keep = [k for k, s in enumerate(subs) if not (metric in ("ddtw", "wddtw") and len(s) < 3)]  # index unspecified there
            ss = [subs[k] for k in keep]
            d, i = wd.pairwise_subsequence_distance(ss, X, metric=prefix + metric, metric_params=mp, return_index=True)
            out[f"{tag}|{ci}|keep"], out[f"{tag}|{ci}|dist"], out[f"{tag}|{ci}|idx"] = np.array(keep), d, i.astype(np.int64)
            paired = [ss[q % len(ss)] for q in range(X.shape[0])]
            d, i = wd.paired_subsequence_distance(paired, X, metric=prefix + metric, metric_params=mp, return_index=True)
            out[f"{tag}|{ci}|paired_dist"], out[f"{tag}|{ci}|paired_idx"] = d, i.astype(np.int64)
    matches(wd, out, X, subs)
    n = X.shape[0]
    Y = np.stack([X[(q + 1) % n, 7 + q:22 + q] for q in range(n)])
    ragged = [subs[SM_SUBS[q % len(SM_SUBS)]] for q in range(n)]
    for ci, (metric, mp) in enumerate(AS_CASES):
        for scale in (False, True):
            for k in (1, 4):
                i_, d_ = wd.argmin_subsequence_distance(Y, X, k=k, metric=metric, metric_params=mp, scale=scale, return_distance=True)
                out[f"as|{ci}|{int(scale)}|{k}|idx"], out[f"as|{ci}|{int(scale)}|{k}|dist"] = i_.astype(np.int64), d_
            i_, d_ = wd.argmin_subsequence_distance(ragged, X, k=3, metric=metric, metric_params=mp, scale=scale, return_distance=True)
            out[f"as|{ci}|{int(scale)}|ragged|idx"], out[f"as|{ci}|{int(scale)}|ragged|dist"] = i_.astype(np.int64), d_
    Yd = np.stack([X[(q + 2) % n, 11 + q:18 + q] for q in range(n)])   # 7 points per subsequence (odd: padding="same" works)
    for ci, (metric, mp) in enumerate(DD_CASES):
        for scale in (False, True):
            for di, (dil, pad) in enumerate(DD_GEOMETRY):
                out[f"dd|{ci}|{int(scale)}|{di}"] = wd.distance_profile(Yd, X, metric=metric, metric_params=mp, scale=scale, dilation=dil,
                                                                        padding=pad)
    path = os.path.join(HERE, "scan_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
