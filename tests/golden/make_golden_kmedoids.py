"""Golden vectors for KMedoids (src/wildboar/distance/_neighbors.py:615-930) from the UNMODIFIED reference (oracle/_ref):

    python tests/golden/make_golden_kmedoids.py   ->  tests/golden/kmedoids_golden.npz

For every case: the input series, the reference's n x n distance matrix, and the fitted attributes (medoid_indices_,
labels_, inertia_, n_iter_) + transform / predict on held-out series.  The `precomputed` cases pin the clustering logic
alone (no distance kernel involved), so the mirror's bookkeeping can be checked without a GPU.
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import ref  # noqa: E402

CASES = [
    # (metric, metric_params, n_clusters, algorithm, init, n_init, seed)
    ("dtw", {"r": 0.1}, 4, "fast", "random", "auto", 1),
    ("dtw", {"r": 0.1}, 3, "pam", "auto", "auto", 2),
    ("msm", {"r": 0.2}, 5, "fast", "min", "auto", 3),
    ("msm", {"r": 0.2}, 3, "pam", "random", 3, 4),
    ("twe", {"r": 0.15}, 4, "pam", "min", "auto", 5),
    ("erp", {"r": 0.3}, 2, "fast", "auto", "auto", 6),
    ("lcss", {"r": 0.5, "epsilon": 0.7}, 3, "fast", "random", 4, 7),
    ("wdtw", {"r": 0.3, "g": 0.1}, 6, "pam", "auto", "auto", 8),
]


def main():
    wd = ref.load()
    if wd is None:
        raise SystemExit("oracle/_ref is not built; run oracle/build_ref.sh first")
    rng = np.random.default_rng(20261020)
    out = {"meta_cases": np.array(repr(CASES))}
    for c, (metric, mp, k, alg, init, n_init, seed) in enumerate(CASES):
        n, T = (70, 40) if alg == "pam" else (110, 48)
        X = np.cumsum(rng.standard_normal((n, T)), axis=1)
        X[: n // 2] += 6.0  # two loose groups so the clustering is not degenerate
        Q = np.cumsum(rng.standard_normal((9, T)), axis=1)
        out[f"{c}|X"], out[f"{c}|Q"] = X, Q
        out[f"{c}|dist"] = wd.pairwise_distance(X, dim="mean", metric=metric, metric_params=mp)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            km = wd.KMedoids(n_clusters=k, metric=metric, metric_params=mp, algorithm=alg, init=init, n_init=n_init, random_state=seed).fit(X)
            kp = wd.KMedoids(n_clusters=k, metric="precomputed", algorithm=alg, init=init, n_init=n_init, random_state=seed).fit(out[f"{c}|dist"])
        for tag, est in (("fit", km), ("pre", kp)):
            out[f"{c}|{tag}|medoids"] = np.asarray(est.medoid_indices_, dtype=np.int64)
            out[f"{c}|{tag}|labels"] = np.asarray(est.labels_, dtype=np.int64)
            out[f"{c}|{tag}|inertia"] = np.float64(est.inertia_)
            out[f"{c}|{tag}|n_iter"] = np.int64(est.n_iter_)
        out[f"{c}|fit|transform"] = km.transform(Q)
        out[f"{c}|fit|predict"] = np.asarray(km.predict(Q), dtype=np.int64)
        out[f"{c}|fit|centers"] = km.cluster_centers_
    np.savez_compressed(os.path.join(HERE, "kmedoids_golden.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
