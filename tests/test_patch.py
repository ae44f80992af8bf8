"""CPU: patch()/unpatch() of an importable wildboar (the oracle/_ref build) routes by metric."""
import numpy as np
import pytest


def test_patch_routes_elastic_metrics_only(wb, monkeypatch):
    from oracle import ref
    wd = ref.load()
    if wd is None:
        pytest.skip("oracle/_ref not built")
    from wildboar_b200 import patch as P, distance as D
    calls = []
    monkeypatch.setattr(D, "pairwise_distance", lambda *a, **k: calls.append(k.get("metric")) or "cuda")
    orig = wd.pairwise_distance
    try:
        assert "wildboar.distance.pairwise_distance" in P.patch()
        x = np.random.default_rng(0).standard_normal((3, 10))
        assert wd.pairwise_distance(x, x.copy(), metric="dtw") == "cuda" and calls == ["dtw"]
        e = wd.pairwise_distance(x, x.copy(), metric="euclidean")  # untouched reference path
        assert isinstance(e, np.ndarray) and e.shape == (3, 3)
    finally:
        P.unpatch()
    assert wd.pairwise_distance is orig


@pytest.mark.gpu
def test_patched_reference_estimators_give_identical_results(wb, oracle):
    """the reference's own KNeighborsClassifier / kneighbors (distance/_neighbors.py:100-283) and the
    silhouette score (metrics/_cluster.py) run unmodified on top of the patched entry points and return
    what they return on the reference's CPU path."""
    from oracle import ref
    wd = ref.load()
    if wd is None:
        pytest.skip("oracle/_ref not built")
    from wildboar.distance import KMedoids, KNeighborsClassifier
    from wildboar.distance._neighbors import NearestNeighbors
    from wildboar_b200 import patch as P
    wb.set_devices([0])
    rng = np.random.default_rng(5)
    proto = np.cumsum(rng.standard_normal((3, 64)), axis=1)
    y = rng.integers(0, 3, 150)
    X = proto[y] + 0.4 * np.cumsum(rng.standard_normal((150, 64)), axis=1)
    Xtr, ytr, Xte = X[:100], y[:100], X[100:]
    results = {}
    for mode in ("cpu", "cuda"):
        if mode == "cuda":
            patched = P.patch()
            assert "wildboar.distance._neighbors.argmin_distance" in patched
        try:
            out = []
            for metric, mp in (("dtw", {"r": 0.1}), ("msm", {"r": 0.2}), ("twe", {"r": 0.2}), ("erp", {"r": 0.1})):
                clf = KNeighborsClassifier(n_neighbors=3, metric=metric, metric_params=mp).fit(Xtr, ytr)
                dist, ind = NearestNeighbors(n_neighbors=3, metric=metric, metric_params=mp).fit(Xtr).kneighbors(Xte, return_distance=True)
                km = KMedoids(n_clusters=3, metric=metric, metric_params=mp, random_state=1).fit(Xtr)
                out.append((clf.predict(Xte), dist, ind, clf.predict_proba(Xte), km.labels_, km.predict(Xte)))
            results[mode] = out
        finally:
            if mode == "cuda":
                P.unpatch()
    for cpu, cuda in zip(results["cpu"], results["cuda"]):
        for a, b in zip(cpu, cuda):
            assert np.array_equal(a, b)


def test_patch_routes_elastic_subsequence_metrics_only(wb, monkeypatch):
    from oracle import ref
    wd = ref.load()
    if wd is None:
        pytest.skip("oracle/_ref not built")
    from wildboar_b200 import patch as P, subsequence as S
    calls = []
    for name in ("pairwise_subsequence_distance", "subsequence_match", "distance_profile", "argmin_subsequence_distance"):
        monkeypatch.setattr(S, name, lambda *a, _n=name, **k: calls.append((_n, k.get("metric"))) or "cuda")
    x = np.cumsum(np.random.default_rng(0).standard_normal((3, 20)), axis=1)
    try:
        done = P.patch()
        assert "wildboar.distance.subsequence_match" in done and "wildboar.distance._distance.distance_profile" in done
        assert wd.pairwise_subsequence_distance(x[0, :5], x, metric="scaled_msm") == "cuda"
        assert wd.subsequence_match(x[0, :5], x, threshold=1.0, metric="twe") == "cuda"
        assert wd.distance_profile(x[:, :5], x, metric="dtw") == "cuda"
        assert wd.argmin_subsequence_distance(x[:, :5], x, k=2, metric="erp", scale=True) == "cuda"
        assert [c[0] for c in calls] == ["pairwise_subsequence_distance", "subsequence_match", "distance_profile", "argmin_subsequence_distance"]
        assert wd.distance_profile(x[:, :5], x, metric="dtw", dilation=2) == "cuda" and len(calls) == 5
        # not elastic: the reference's own code paths
        e = wd.pairwise_subsequence_distance(x[0, :5], x, metric="euclidean")
        assert isinstance(e, np.ndarray) and e.shape == (3,)
        dp = wd.distance_profile(x[:, :5], x, metric="manhattan", dilation=2)
        assert isinstance(dp, np.ndarray) and len(calls) == 5
    finally:
        P.unpatch()
    assert not hasattr(wd.subsequence_match, "__wildboar_b200_original__")


@pytest.mark.gpu
def test_patched_reference_subsequence_callers_give_identical_results(wb):
    """The reference's own motif annotation (annotate/_motifs.py, built on subsequence_match) runs unmodified on the patched
    entry points and returns what it returns on its CPU path."""
    from oracle import ref
    wd = ref.load()
    if wd is None:
        pytest.skip("oracle/_ref not built")
    from wildboar_b200 import patch as P
    wb.set_devices([0])
    rng = np.random.default_rng(8)
    X = np.cumsum(rng.standard_normal((12, 90)), axis=1)
    s = X[2, 30:50].copy()
    results = {}
    for mode in ("cpu", "cuda"):
        if mode == "cuda":
            P.patch()
        try:
            out = []
            for metric, mp in (("dtw", {"r": 0.1}), ("scaled_dtw", {"r": 0.1}), ("msm", {"r": 0.2}), ("scaled_twe", {"r": 0.2})):
                i_, d_ = wd.subsequence_match(s, X, threshold="auto", metric=metric, metric_params=mp, exclude=0.25, return_distance=True)
                out.append((i_, d_))
                out.append(wd.pairwise_subsequence_distance([s, s[:7]], X, metric=metric, metric_params=mp, return_index=True))
                out.append((wd.distance_profile(np.stack([s] * 12), X, metric=metric, metric_params=mp),))
            results[mode] = out
        finally:
            if mode == "cuda":
                P.unpatch()
    def same(a, b):
        if a is None or b is None:
            return a is None and b is None
        if isinstance(a, np.ndarray) and a.dtype == object:
            return len(a) == len(b) and all(same(p, q) for p, q in zip(a, b))
        return np.array_equal(a, b)
    for cpu, cuda in zip(results["cpu"], results["cuda"]):
        for a, b in zip(cpu, cuda):
            assert same(a, b)
