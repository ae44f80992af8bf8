"""Golden vectors for the SURVEY 8f "next" rows, generated from the UNMODIFIED reference (oracle/_ref).

    python tests/golden/make_golden_next.py      # needs oracle/_ref (oracle/build_ref.sh)

Writes tests/golden/next_golden.npz:
  nd|...   multivariate pairwise / self / paired with dim="mean" / "full" (_distance.py:1245-1297, 1163-1169)
  al|... dba|... km|...  dtw_alignment / dtw_mapping / dtw_average (distance/dtw.py:246-690), KMeans(metric="dtw")
  ee|...   ElasticEnsembleClassifier (ensemble/_elastic.py): cross-validation scores, chosen parameters, probabilities
  ss|...   pairwise / paired_subsequence_distance, DTW family (_distance.py:543-729)
  nb|...   KNeighborsClassifier.predict_proba / predict and NearestNeighbors.kneighbors (distance/_neighbors.py:19-300)
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


NB_CASES = [("dtw", {"r": 0.1}), ("wdtw", {"r": 0.3, "g": 0.1}), ("msm", {"r": 0.2}), ("erp", {"r": 0.2}),
            ("lcss", {"r": 0.5, "epsilon": 0.7}), ("twe", {"r": 0.15}), ("edr", {"r": 0.3}), ("adtw", {"r": 0.1, "p": 0.5})]


def neighbors(wd, out):
    """KNeighborsClassifier / NearestNeighbors of the reference (distance/_neighbors.py) on seeded random walks."""
    rng = np.random.default_rng(20261019)
    X = np.cumsum(rng.standard_normal((60, 50)), axis=1)
    y = rng.integers(0, 3, 60) * 2 + 1            # labels {1, 3, 5}
    Q = np.cumsum(rng.standard_normal((25, 50)), axis=1)
    X3 = np.cumsum(rng.standard_normal((40, 2, 30)), axis=2)
    y3 = rng.integers(0, 2, 40)
    Q3 = np.cumsum(rng.standard_normal((15, 2, 30)), axis=2)
    for k, v in dict(X=X, y=y, Q=Q, X3=X3, y3=y3, Q3=Q3).items():
        out[f"nb|{k}"] = v
    for metric, mp in NB_CASES:
        for k in (1, 3):
            clf = wd.KNeighborsClassifier(n_neighbors=k, metric=metric, metric_params=mp).fit(X, y)
            out[f"nb|{metric}|knn{k}|proba"] = clf.predict_proba(Q)
            out[f"nb|{metric}|knn{k}|predict"] = clf.predict(Q)
        nn = wd.NearestNeighbors(n_neighbors=4, metric=metric, metric_params=mp).fit(X)
        d, i = nn.kneighbors(Q)
        out[f"nb|{metric}|nn|dist"], out[f"nb|{metric}|nn|ind"] = d, i.astype(np.int64)
        try:
            d, i = nn.kneighbors()
            out[f"nb|{metric}|nnself|dist"], out[f"nb|{metric}|nnself|ind"] = d, i.astype(np.int64)
        except ValueError:  # the reference cannot drop the query itself when it is not among the k+1 (ties / abandoning)
            out[f"nb|{metric}|nnself|raises"] = np.array(1)
    for metric, mp in NB_CASES[:3]:
        clf = wd.KNeighborsClassifier(n_neighbors=3, metric=metric, metric_params=mp).fit(X3, y3)
        out[f"nb3|{metric}|knn3|proba"] = clf.predict_proba(Q3)
        nn = wd.NearestNeighbors(n_neighbors=3, metric=metric, metric_params=mp).fit(X3)
        d, i = nn.kneighbors(Q3)
        out[f"nb3|{metric}|nn|dist"], out[f"nb3|{metric}|nn|ind"] = d, i.astype(np.int64)
        d, i = nn.kneighbors()
        out[f"nb3|{metric}|nnself|dist"], out[f"nb3|{metric}|nnself|ind"] = d, i.astype(np.int64)


def alignments(wd, out):
    """dtw_alignment / wdtw_alignment / dtw_mapping / dtw_average / KMeans(metric="dtw") of the reference."""
    from wildboar.distance import dtw as rd
    from wildboar.distance import KMeans
    rng = np.random.default_rng(20261020)
    shapes = [(30, 30, 0.1), (25, 40, 0.2), (40, 25, 0.3), (16, 16, 1.0), (12, 12, 0.0), (1, 5, 0.5), (5, 1, 0.5), (64, 64, 0.05)]
    out["al|shapes"] = np.array(shapes)
    for k, (tx, ty, r) in enumerate(shapes):
        x = np.cumsum(rng.standard_normal(tx))
        y = np.cumsum(rng.standard_normal(ty))
        out[f"al|{k}|x"], out[f"al|{k}|y"] = x, y
        for name, fn, kw in (("dtw", rd.dtw_alignment, {}), ("wdtw", rd.wdtw_alignment, {"g": 0.1})):
            # the reference leaves the cells outside the band uninitialised: pre-fill them with NaN
            a = fn(x, y, r=r, out=np.full((tx, ty), np.nan), **kw)
            out[f"al|{k}|{name}|matrix"] = a
            out[f"al|{k}|{name}|path"] = rd.dtw_mapping(alignment=a)
    X = np.cumsum(rng.standard_normal((24, 48)), axis=1)
    sw = rng.random(24) + 0.5
    out["dba|X"], out["dba|sw"] = X, sw
    for name, kw in (("mm", {}), ("mm_g", {"g": 0.15}), ("mm_sw", {"sample_weight": sw}), ("mm_r1", {"r": 1.0}),
                     ("ssg", {"method": "ssg", "random_state": 3, "max_epoch": 6}),
                     ("random", {"init": "random", "random_state": 5})):
        args = dict(r=0.2, init=X[7], method="mm", return_cost=True)
        args.update(kw)
        mean, cost = rd.dtw_average(X, **args)
        out[f"dba|{name}|mean"], out[f"dba|{name}|cost"] = mean, np.array(cost)
    Xk = np.concatenate([np.cumsum(rng.standard_normal((20, 40)), axis=1) + off for off in (0.0, 6.0, -6.0)])
    out["km|X"] = Xk
    for name, kw in (("dtw", {}), ("wdtw", {"g": 0.1}), ("k7", {"n_clusters": 7, "n_init": 2, "r": 0.1})):
        args = dict(n_clusters=3, metric="dtw", r=0.2, random_state=11, max_iter=20)
        args.update(kw)
        km = KMeans(**args).fit(Xk)
        out[f"km|{name}|centers"], out[f"km|{name}|labels"] = km.cluster_centers_, km.labels_.astype(np.int64)
        out[f"km|{name}|inertia"], out[f"km|{name}|n_iter"] = np.array(km.inertia_), np.array(km.n_iter_)
        out[f"km|{name}|transform"] = km.transform(Xk[::5])


SS_CASES = [("dtw", {"r": 0.1}), ("dtw", {"r": 1.0}), ("wdtw", {"r": 0.3, "g": 0.1}), ("adtw", {"r": 0.2, "p": 0.5}),
            ("ddtw", {"r": 0.2}), ("wddtw", {"r": 0.5, "g": 0.2})]
SS_LENGTHS = (12, 30, 5, 3, 80, 1, 2)


def subsequences(wd, out):
    """pairwise / paired_subsequence_distance of the reference for the DTW-family subsequence metrics."""
    rng = np.random.default_rng(20261021)
    X = np.cumsum(rng.standard_normal((9, 80)), axis=1)
    subs = [np.cumsum(rng.standard_normal(m)) for m in SS_LENGTHS]
    out["ss|X"] = X
    for k, s in enumerate(subs):
        out[f"ss|s{k}"] = s
    for ci, (metric, mp) in enumerate(SS_CASES):
        keep = [k for k, s in enumerate(subs) if not (metric in ("ddtw", "wddtw") and len(s) < 3)]  # index unspecified there
        ss = [subs[k] for k in keep]
        d, i = wd.pairwise_subsequence_distance(ss, X, metric=metric, metric_params=mp, return_index=True)
        out[f"ss|{ci}|keep"], out[f"ss|{ci}|dist"], out[f"ss|{ci}|idx"] = np.array(keep), d, i.astype(np.int64)
        paired = [subs[keep[q % len(keep)]] for q in range(X.shape[0])]
        d, i = wd.paired_subsequence_distance(paired, X, metric=metric, metric_params=mp, return_index=True)
        out[f"ss|{ci}|paired_dist"], out[f"ss|{ci}|paired_idx"] = d, i.astype(np.int64)
    # scaled_dtw (UCR suite): subsequences of >= 3 samples, none constant (all windows tie at sqrt(m) there and the
    # reference's bounds break the tie at rounding level)
    Xs = np.cumsum(rng.standard_normal((11, 120)), axis=1)
    ssubs = [np.cumsum(rng.standard_normal(m)) for m in (3, 4, 5, 6, 17, 40, 120, 64)]
    ssubs.append(Xs[4, 20:52].copy() * 3.0 + 7.0)          # a scaled, shifted copy of a window: distance 0
    out["ssc|X"] = Xs
    for k, s in enumerate(ssubs):
        out[f"ssc|s{k}"] = s
    out["ssc|n"] = np.array(len(ssubs))
    for r in (0.0, 0.05, 0.1, 0.3, 1.0):
        d, i = wd.pairwise_subsequence_distance(ssubs, Xs, metric="scaled_dtw", metric_params={"r": r}, return_index=True)
        out[f"ssc|{r}|dist"], out[f"ssc|{r}|idx"] = d, i.astype(np.int64)
    d, i = wd.paired_subsequence_distance([ssubs[q % len(ssubs)] for q in range(11)], Xs, metric="dtw", scale=True,
                                          metric_params={"r": 0.1}, return_index=True)
    out["ssc|paired_dist"], out["ssc|paired_idx"] = d, i.astype(np.int64)


def ensembles(wd, out):
    """ElasticEnsembleClassifier of the reference (ensemble/_elastic.py): scores, chosen parameters, probabilities."""
    from wildboar.ensemble import ElasticEnsembleClassifier
    rng = np.random.default_rng(20261022)
    X = np.concatenate([np.cumsum(rng.standard_normal((18, 36)), axis=1) + off for off in (0.0, 1.5, -1.5)])
    y = np.repeat([2, 5, 9], 18)
    perm = rng.permutation(54)
    X, y = X[perm], y[perm]
    Q = np.cumsum(rng.standard_normal((12, 36)), axis=1)
    out["ee|X"], out["ee|y"], out["ee|Q"] = X, y, Q
    for name, kw in (("auto_k1", dict(n_neighbors=1, metric="auto")),
                     ("custom_k3", dict(n_neighbors=3, metric={"dtw": {"min_r": 0.1, "max_r": 0.3, "num_r": 3},
                                                                 "msm": {"min_c": 0.1, "max_c": 10, "num_c": 4},
                                                                 "lcss": {"min_r": 0.0, "max_r": 0.25, "num_r": 2,
                                                                          "min_epsilon": 0.3, "max_epsilon": 1.2, "num_epsilon": 2}}))):
        clf = ElasticEnsembleClassifier(**kw).fit(X, y)
        out[f"ee|{name}|scores"] = np.array([s for _, s in clf.scores_])
        out[f"ee|{name}|metrics"] = np.array([m for m, _ in clf.scores_])
        out[f"ee|{name}|params"] = np.array(repr([{k: float(v) for k, v in e.metric_params.items()} for e in clf.estimators_]))
        out[f"ee|{name}|proba"] = clf.predict_proba(Q)
        out[f"ee|{name}|predict"] = clf.predict(Q)


def main():
    wd = ref.load()
    if wd is None:
        raise SystemExit("oracle/_ref is not built; run oracle/build_ref.sh first")
    out = {}
    multivariate(wd, out)
    neighbors(wd, out)
    alignments(wd, out)
    subsequences(wd, out)
    ensembles(wd, out)
    path = os.path.join(HERE, "next_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
