"""Golden vectors for SURVEY 8f-4, second half, generated from the UNMODIFIED reference (oracle/_ref).

    python tests/golden/make_golden_scan.py      # needs oracle/_ref (oracle/build_ref.sh)

Writes tests/golden/scan_golden.npz:
  sc|...   pairwise / paired_subsequence_distance for lcss / erp / edr / msm / twe (Lcss ... Erp SubsequenceMetric,
           _elastic.pyx:2616-3124) -- the metrics whose early abandoning decides which window wins
  sw|...   the same calls for the generic scaled metrics scaled_<metric> = ScaledSubsequenceMetricWrap(Metric)
           (_cdistance.pyx:470-551) over adtw, wdtw, ddtw, wddtw, lcss, erp, edr, msm, twe
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
    path = os.path.join(HERE, "scan_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
