"""SURVEY 8f "next" rows: multivariate dim="mean"/"full" combined on the device.

CPU tests pin the oracle (+ numpy's reduction order) to the reference's golden vectors; `-m gpu` tests
compare the CUDA path (public API -> ctypes -> C ABI) with the oracle and the golden vectors, bit for bit.
"""
import os

import numpy as np
import pytest

from util import random_walks

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
ND_METRICS = ["dtw", "wdtw", "ddtw", "adtw", "lcss", "erp", "edr", "msm", "twe"]


@pytest.fixture(scope="module")
def next_golden():
    with np.load(os.path.join(ROOT, "tests", "golden", "next_golden.npz")) as z:
        return {k: z[k] for k in z.files}


def _oracle_nd(oracle, kind, metric, x, y, dim, **mp):
    """Per-dimension oracle calls combined like the reference (np.mean / np.stack over axis 0)."""
    nd = x.shape[1]
    if kind == "pairwise":
        per = [oracle.pairwise(metric, x[:, d], y[:, d], **mp) for d in range(nd)]
    elif kind == "self":
        per = [oracle.pairwise(metric, x[:, d], None, **mp) for d in range(nd)]
    else:
        per = [oracle.paired(metric, x[:, d], y[:, d], **mp) for d in range(nd)]
    return np.mean(per, axis=0) if dim == "mean" else np.stack(per, axis=0)


# ---------------------------------------------------------------------------------------------
# CPU: the oracle restatement equals the reference on multivariate input
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("metric", ND_METRICS)
def test_oracle_multivariate_matches_reference_golden(oracle, next_golden, metric):
    x, y = next_golden["nd|x"], next_golden["nd|y"]
    for dim in ("mean", "full"):
        assert np.array_equal(_oracle_nd(oracle, "pairwise", metric, x, y, dim, r=0.25), next_golden[f"nd|{metric}|{dim}|pairwise"])
        assert np.array_equal(_oracle_nd(oracle, "self", metric, x, None, dim, r=0.25), next_golden[f"nd|{metric}|{dim}|self"])
        assert np.array_equal(_oracle_nd(oracle, "paired", metric, x[:5], y, dim, r=0.25), next_golden[f"nd|{metric}|{dim}|paired"])


# ---------------------------------------------------------------------------------------------
# GPU
# ---------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def W(wb):
    if wb.device_count() < 1:
        pytest.skip("no CUDA device")
    wb.set_devices([0])
    return wb


@pytest.mark.gpu
@pytest.mark.parametrize("metric", ND_METRICS)
def test_multivariate_golden(W, next_golden, metric):
    x, y = next_golden["nd|x"], next_golden["nd|y"]
    mp = {"r": 0.25}
    for dim in ("mean", "full"):
        assert np.array_equal(W.pairwise_distance(x, y, dim=dim, metric=metric, metric_params=mp), next_golden[f"nd|{metric}|{dim}|pairwise"])
        assert np.array_equal(W.pairwise_distance(x, dim=dim, metric=metric, metric_params=mp), next_golden[f"nd|{metric}|{dim}|self"])
        assert np.array_equal(W.paired_distance(x[:5], y, dim=dim, metric=metric, metric_params=mp), next_golden[f"nd|{metric}|{dim}|paired"])
        assert W.last_stats()["launches"] >= x.shape[1]  # every dimension ran on the device, in one library call


@pytest.mark.gpu
def test_multivariate_large_strided_and_chunked(W, oracle):
    """Several result slabs per matrix (48 MB each), strided 3-D views, unequal lengths, multi-device split."""
    rng = np.random.default_rng(5)
    big = np.cumsum(rng.standard_normal((900, 4, 2, 48)), axis=3)   # view [:, ::2] has a dim stride of 2 * 2 * 48
    x = big[:, ::2, 0, :]                                            # (900, 2, 48), non-contiguous samples and dims
    y = np.cumsum(rng.standard_normal((8000, 2, 40)), axis=2)        # 900 x 8000 doubles = 57.6 MB -> two slabs
    for metric in ("dtw", "erp"):
        for dim in ("mean", "full"):
            got = W.pairwise_distance(x, y, dim=dim, metric=metric, metric_params={"r": 0.1})
            sel = rng.integers(0, 900, 40)
            want = _oracle_nd(oracle, "pairwise", metric, np.ascontiguousarray(x[sel]), y, dim, r=0.1)
            assert np.array_equal(got[..., sel, :], want), (metric, dim)
    xs = np.ascontiguousarray(x[:300])
    got = W.pairwise_distance(xs, dim="mean", metric="msm", metric_params={"r": 0.2})
    assert np.array_equal(got, _oracle_nd(oracle, "self", "msm", xs, None, "mean", r=0.2))
    if W.device_count() >= 2:
        W.set_devices([0, 1])
        try:
            two = W.pairwise_distance(xs, dim="mean", metric="msm", metric_params={"r": 0.2})
        finally:
            W.set_devices([0])
        assert np.array_equal(two, got)


# ---------------------------------------------------------------------------------------------
# nearest-neighbour estimators with the training set resident on the device
# ---------------------------------------------------------------------------------------------
NB_CASES = [("dtw", {"r": 0.1}), ("wdtw", {"r": 0.3, "g": 0.1}), ("msm", {"r": 0.2}), ("erp", {"r": 0.2}),
            ("lcss", {"r": 0.5, "epsilon": 0.7}), ("twe", {"r": 0.15}), ("edr", {"r": 0.3}), ("adtw", {"r": 0.1, "p": 0.5})]


def test_neighbors_host_logic(wb):
    from wildboar_b200.neighbors import KNeighborsClassifier, NearestNeighbors
    clf = KNeighborsClassifier(3, metric="dtw", metric_params={"r": 0.1})
    assert clf.get_params() == {"metric": "dtw", "metric_params": {"r": 0.1}, "n_jobs": None, "n_neighbors": 3}
    with pytest.raises(ValueError, match="n_neighbors"):
        KNeighborsClassifier(0, metric="dtw").fit(np.zeros((3, 8)), [0, 1, 0])
    with pytest.raises(ValueError, match="metric"):
        KNeighborsClassifier(1, metric="euclidean").fit(np.zeros((3, 8)), [0, 1, 0])
    with pytest.raises(ValueError, match="inconsistent numbers of samples"):
        KNeighborsClassifier(1, metric="dtw").fit(np.zeros((3, 8)), [0, 1])
    with pytest.raises(ValueError, match="continuous"):
        KNeighborsClassifier(1, metric="dtw").fit(np.zeros((3, 8)), [0.5, 1.0, 2.0])
    with pytest.raises(Exception, match="not fitted"):
        NearestNeighbors(metric="dtw").kneighbors()
    if wb.device_count() == 0:  # no GPU: fitting must fail loudly, not fall back
        with pytest.raises(RuntimeError, match="no CUDA device|no CPU fallback"):
            NearestNeighbors(metric="dtw").fit(np.zeros((3, 8)))


@pytest.mark.gpu
@pytest.mark.parametrize("metric,mp", NB_CASES)
def test_neighbors_match_reference_golden(W, next_golden, metric, mp):
    from wildboar_b200.neighbors import KNeighborsClassifier, NearestNeighbors
    g = next_golden
    X, y, Q = g["nb|X"], g["nb|y"], g["nb|Q"]
    for k in (1, 3):
        clf = KNeighborsClassifier(n_neighbors=k, metric=metric, metric_params=mp).fit(X, y)
        assert np.array_equal(clf.predict_proba(Q), g[f"nb|{metric}|knn{k}|proba"])
        assert np.array_equal(clf.predict(Q), g[f"nb|{metric}|knn{k}|predict"])
    nn = NearestNeighbors(n_neighbors=4, metric=metric, metric_params=mp).fit(X)
    d, i = nn.kneighbors(Q)
    assert np.array_equal(d, g[f"nb|{metric}|nn|dist"]) and np.array_equal(i, g[f"nb|{metric}|nn|ind"])
    assert np.array_equal(nn.kneighbors(Q, return_distance=False), g[f"nb|{metric}|nn|ind"])
    if f"nb|{metric}|nnself|raises" in g:
        with pytest.raises(ValueError):   # the reference fails the same way (query not among its own k+1 neighbours)
            nn.kneighbors()
    else:
        d, i = nn.kneighbors()
        assert np.array_equal(d, g[f"nb|{metric}|nnself|dist"]) and np.array_equal(i, g[f"nb|{metric}|nnself|ind"])
    # the training set stays on the device between queries; releasing it re-uploads transparently
    nn.release()
    assert np.array_equal(nn.kneighbors(Q[:3])[1], g[f"nb|{metric}|nn|ind"][:3])


@pytest.mark.gpu
@pytest.mark.parametrize("metric,mp", NB_CASES[:3])
def test_neighbors_multivariate_match_reference_golden(W, next_golden, metric, mp):
    from wildboar_b200.neighbors import KNeighborsClassifier, NearestNeighbors
    g = next_golden
    X3, y3, Q3 = g["nb|X3"], g["nb|y3"], g["nb|Q3"]
    clf = KNeighborsClassifier(n_neighbors=3, metric=metric, metric_params=mp).fit(X3, y3)
    assert np.array_equal(clf.predict_proba(Q3), g[f"nb3|{metric}|knn3|proba"])
    nn = NearestNeighbors(n_neighbors=3, metric=metric, metric_params=mp).fit(X3)
    d, i = nn.kneighbors(Q3)
    assert np.array_equal(d, g[f"nb3|{metric}|nn|dist"]) and np.array_equal(i, g[f"nb3|{metric}|nn|ind"])
    d, i = nn.kneighbors()
    assert np.array_equal(d, g[f"nb3|{metric}|nnself|dist"]) and np.array_equal(i, g[f"nb3|{metric}|nnself|ind"])


@pytest.mark.gpu
def test_fitted_set_equals_host_calls(W, oracle):
    """wb_cuda_*_fitted against the plain host-buffer entry points and the oracle (larger, two devices if present)."""
    from wildboar_b200 import _shim
    from wildboar_b200.distance import DtwMetric, MsmMetric
    refs, q = random_walks(3000, 64, 31), random_walks(200, 64, 32)
    devs = [0, 1] if W.device_count() >= 2 else [0]
    fit = _shim.FittedSet(refs.reshape(3000, 1, 64), devices=devs)
    for m in (DtwMetric(r=0.1), MsmMetric(r=0.1)):
        idx, dist = _shim.argmin_fitted(m.metric_id, m._params(), q, fit, 5, use_device_lb=m.name == "dtw")
        oi, od = oracle.argmin(m.name, q, refs, k=5, r=0.1)
        assert np.array_equal(idx, oi) and np.array_equal(dist, od)
        pw = _shim.pairwise_fitted(m.metric_id, m._params(), q.reshape(200, 1, 64), fit)
        assert np.array_equal(pw, oracle.pairwise(m.name, q, refs, r=0.1))
    fit.close()
    with pytest.raises(RuntimeError, match="released"):
        _shim.pairwise_fitted(m.metric_id, m._params(), q.reshape(200, 1, 64), fit)
