"""Golden vectors for MDS (src/wildboar/distance/_manifold.py) from the UNMODIFIED reference (oracle/_ref):

    python tests/golden/make_golden_mds.py   ->  tests/golden/mds_golden.npz

Embeddings and final stress of `MDS(n_components, metric_mds, n_init=2, max_iter=60, random_state, metric, metric_params)
.fit_transform(X)` for seeded random walks; scikit-learn's SMACOF runs on the reference's own dissimilarity matrix.
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import ref  # noqa: E402

CASES = [("dtw", {"r": 0.1}, 2, True, 1), ("msm", {"r": 0.2}, 3, True, 2), ("twe", {"r": 0.15}, 2, False, 3), ("erp", {"r": 0.3}, 2, True, 4)]


def main():
    wd = ref.load()
    if wd is None:
        raise SystemExit("oracle/_ref is not built; run oracle/build_ref.sh first")
    rng = np.random.default_rng(20261021)
    out = {"meta_cases": np.array(repr(CASES))}
    for c, (metric, mp, nc, mm, seed) in enumerate(CASES):
        X = np.cumsum(rng.standard_normal((40, 36)), axis=1)
        out[f"{c}|X"] = X
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            est = wd.MDS(n_components=nc, metric_mds=mm, n_init=2, max_iter=60, random_state=seed, metric=metric, metric_params=mp)
            out[f"{c}|emb"] = est.fit_transform(X)
            out[f"{c}|stress"] = np.float64(est.mds_.stress_)
    np.savez_compressed(os.path.join(HERE, "mds_golden.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
