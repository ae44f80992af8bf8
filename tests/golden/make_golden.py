"""Generate golden vectors from the UNMODIFIED reference (wildboar built into oracle/_ref).

Run in the build container (needs oracle/_ref, i.e. `oracle/build_ref.sh`):

    python tests/golden/make_golden.py

Writes tests/golden/elastic_golden.npz.  Inputs are seeded random walks; outputs are what
the reference's own `pairwise_distance`, `paired_distance` and `argmin_distance`
(src/wildboar/distance/_distance.py:1178,1082,1320) return for them.  The reference's own
golden tables (tests/wildboar/distance/test_distance.py:26-181) cannot be replayed offline
because their inputs are downloaded datasets, so these vectors pin the same code path on
synthetic inputs instead.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import ref  # noqa: E402

METRICS = ["dtw", "wdtw", "ddtw", "adtw", "lcss", "erp", "edr", "msm", "twe", "wddtw", "wlcss"]
# non-default parameter sets exercised in addition to the defaults
EXTRA = {
    "wdtw": {"g": 0.3}, "adtw": {"p": 0.1}, "lcss": {"epsilon": 0.4}, "erp": {"g": 0.5},
    "edr": {"epsilon": 0.3}, "msm": {"c": 0.37}, "twe": {"penalty": 0.5, "stiffness": 0.05},
    "wlcss": {"epsilon": 0.6, "g": 0.2},
}
SHAPES = [  # (nx, ny, Tx, Ty)
    (4, 5, 24, 24), (3, 4, 17, 29), (4, 3, 29, 17), (3, 3, 3, 3), (2, 3, 1, 1), (2, 2, 2, 6), (3, 3, 64, 64),
]
RS = [0.0, 0.05, 0.1, 0.3, 1.0]


def main():
    wd = ref.load()
    if wd is None:
        raise SystemExit("oracle/_ref is not built; run oracle/build_ref.sh first")
    out = {}
    rng = np.random.default_rng(20261017)
    case = 0
    for (nx, ny, Tx, Ty) in SHAPES:
        x = np.cumsum(rng.standard_normal((nx, Tx)), axis=1)
        y = np.cumsum(rng.standard_normal((ny, Ty)), axis=1)
        out[f"x{case}"] = x
        out[f"y{case}"] = y
        for metric in METRICS:
            if metric == "wddtw" and Tx > Ty:
                continue  # reference overflows a heap buffer there
            for pi, extra in enumerate([{}, EXTRA.get(metric)]):
                if extra is None:
                    continue
                for r in RS:
                    mp = dict(extra, r=r)
                    key = f"{case}|{metric}|{pi}|{r}"
                    out[key + "|pairwise"] = wd.pairwise_distance(x, y, metric=metric, metric_params=mp)
                    if Tx == Ty:
                        out[key + "|self"] = wd.pairwise_distance(x, metric=metric, metric_params=mp)
                        n = min(nx, ny)
                        out[key + "|paired"] = wd.paired_distance(x[:n], y[:n], metric=metric, metric_params=mp)
                    if not (metric in ("ddtw", "wddtw") and min(Tx, Ty) < 3):
                        for k in (1, 2):
                            idx, dist = wd.argmin_distance(x, y, k=k, metric=metric, metric_params=mp,
                                                           return_distance=True)
                            out[key + f"|argmin{k}|idx"] = idx.astype(np.int64)
                            out[key + f"|argmin{k}|dist"] = dist
        case += 1
    out["meta_extra"] = np.array(repr(EXTRA))
    path = os.path.join(HERE, "elastic_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
