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
