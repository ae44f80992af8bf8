"""Golden vectors for the SURVEY 8f "next" rows, generated from the UNMODIFIED reference (oracle/_ref).

    python tests/golden/make_golden_next.py      # needs oracle/_ref (oracle/build_ref.sh)

Writes tests/golden/next_golden.npz:
  nd|...   multivariate pairwise / self / paired with dim="mean" / "full" (_distance.py:1245-1297, 1163-1169)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import ref  # noqa: E402

ND_METRICS = ["dtw", "wdtw", "ddtw", "adtw", "lcss", "erp", "edr", "msm", "twe"]


def multivariate(wd, out):
    rng = np.random.default_rng(20261018)
    x = np.cumsum(rng.standard_normal((7, 3, 40)), axis=2)
    y = np.cumsum(rng.standard_normal((5, 3, 40)), axis=2)
    out["nd|x"], out["nd|y"] = x, y
    for metric in ND_METRICS:
        mp = {"r": 0.25}
        for dim in ("mean", "full"):
            out[f"nd|{metric}|{dim}|pairwise"] = wd.pairwise_distance(x, y, dim=dim, metric=metric, metric_params=mp)
            out[f"nd|{metric}|{dim}|self"] = wd.pairwise_distance(x, dim=dim, metric=metric, metric_params=mp)
            out[f"nd|{metric}|{dim}|paired"] = wd.paired_distance(x[:5], y, dim=dim, metric=metric, metric_params=mp)


def main():
    wd = ref.load()
    if wd is None:
        raise SystemExit("oracle/_ref is not built; run oracle/build_ref.sh first")
    out = {}
    multivariate(wd, out)
    path = os.path.join(HERE, "next_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
