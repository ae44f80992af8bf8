"""Generate golden vectors for `dtw_envelop` / `dtw_lb_keogh` (src/wildboar/distance/dtw.py:155-243) from the UNMODIFIED
reference (wildboar built into oracle/_ref; run in the build container after `oracle/build_ref.sh`):

    python tests/golden/make_golden_dtw_lb.py   ->  tests/golden/dtw_lb_golden.npz
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import ref  # noqa: E402

LENGTHS = [1, 2, 3, 9, 40, 128, 257]
RS = [0.0, 0.05, 0.1, 0.5, 1.0]


def main():
    if ref.load() is None:
        raise SystemExit("oracle/_ref is not built; run oracle/build_ref.sh first")
    from wildboar.distance.dtw import dtw_envelop, dtw_lb_keogh
    rng = np.random.default_rng(20261019)
    out = {}
    for T in LENGTHS:
        x = np.cumsum(rng.standard_normal(T))
        y = np.cumsum(rng.standard_normal(T))
        out[f"x|{T}"], out[f"y|{T}"] = x, y
        for r in RS:
            lo, hi = dtw_envelop(y, r=r)
            out[f"lower|{T}|{r}"], out[f"upper|{T}|{r}"] = lo, hi
            md, cb = dtw_lb_keogh(x, y, r=r)
            out[f"min_dist|{T}|{r}"], out[f"cb|{T}|{r}"] = np.float64(md), cb
            md2, cb2 = dtw_lb_keogh(x, lower=lo, upper=hi)
            assert md2 == md and np.array_equal(cb, cb2)
    np.savez_compressed(os.path.join(HERE, "dtw_lb_golden.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
