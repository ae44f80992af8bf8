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
