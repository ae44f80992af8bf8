"""Generate golden vectors for the DTW lower-bound transformers from the UNMODIFIED reference
(wildboar built into oracle/_ref; run in the build container after `oracle/build_ref.sh`):

    python tests/golden/make_golden_lb.py   ->  tests/golden/lb_golden.npz

Outputs are what `DtwKeoghLowerBound(r, kind).fit(X).transform(Q)` and
`DtwKimLowerBound().fit(X).transform(Q)` (src/wildboar/distance/lb.py:198-432) return for seeded
random walks.  The reference's own golden tables (tests/wildboar/distance/test_lb.py:34-91) use
a downloaded dataset and cannot be replayed offline.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import ref  # noqa: E402

SHAPES = [(5, 7, 40), (3, 4, 9), (4, 3, 3), (2, 3, 2), (3, 2, 1), (6, 5, 128)]  # (n_query, n_fit, T)
RS = [0.0, 0.05, 0.1, 0.5, 1.0]
KINDS = ["both", "left", "right"]


def main():
    if ref.load() is None:
        raise SystemExit("oracle/_ref is not built; run oracle/build_ref.sh first")
    from wildboar.distance.lb import DtwKeoghLowerBound, DtwKimLowerBound
    rng = np.random.default_rng(20261018)
    out = {}
    for c, (nq, nx, T) in enumerate(SHAPES):
        q = np.cumsum(rng.standard_normal((nq, T)), axis=1)
        x = np.cumsum(rng.standard_normal((nx, T)), axis=1)
        out[f"q{c}"], out[f"x{c}"] = q, x
        for r in RS:
            for kind in KINDS:
                out[f"{c}|keogh|{r}|{kind}"] = DtwKeoghLowerBound(r=r, kind=kind).fit(x).transform(q)
        out[f"{c}|kim"] = DtwKimLowerBound().fit(x).transform(q)
    np.savez_compressed(os.path.join(HERE, "lb_golden.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
