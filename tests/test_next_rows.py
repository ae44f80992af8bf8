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


# ---------------------------------------------------------------------------------------------
# DTW alignment, warping path, DBA, KMeans(metric="dtw")
# ---------------------------------------------------------------------------------------------
def _al_cases(g):
    for k, (tx, ty, r) in enumerate(g["al|shapes"]):
        yield k, int(tx), int(ty), float(r)


def _band_equal(got, want):
    """Equal wherever the reference wrote a value (NaN = cells its np.empty matrix leaves uninitialised)."""
    written = ~np.isnan(want)
    return np.array_equal(got[written], want[written])


def test_oracle_alignment_path_dba_match_reference_golden(oracle, next_golden):
    g = next_golden
    for k, tx, ty, r in _al_cases(g):
        x, y = g[f"al|{k}|x"], g[f"al|{k}|y"]
        for name, w in (("dtw", None), ("wdtw", oracle.jeong_weight(max(tx, ty), 0.1))):
            a = oracle.dtw_alignment(x, y, r=r, weight=w)
            assert np.array_equal(a, g[f"al|{k}|{name}|matrix"], equal_nan=True), (k, name)
            lo, hi = oracle.dtw_path(a)
            cols = np.arange(ty)[None, :]
            assert np.array_equal((cols >= lo[:, None]) & (cols <= hi[:, None]), g[f"al|{k}|{name}|path"]), (k, name)
    X, sw = g["dba|X"], g["dba|sw"]
    for name, kw in (("mm", {}), ("mm_g", {"g": 0.15}), ("mm_sw", {"sample_weight": sw}), ("mm_r1", {"r": 1.0})):
        args = dict(r=0.2, init=X[7])
        args.update(kw)
        mean, cost = oracle.dtw_average_mm(X, **args)
        assert np.array_equal(mean, g[f"dba|{name}|mean"]) and cost == g[f"dba|{name}|cost"], name


def test_dtw_module_host_logic(wb):
    from wildboar_b200 import dtw
    assert np.array_equal(dtw.jeong_weight(5, 0.3), 1.0 / (1.0 + np.exp(-0.3 * (np.arange(5, dtype=float) - 2.5))))
    with pytest.raises(ValueError, match="r =="):
        dtw.dtw_alignment(np.zeros(4), np.zeros(4), r=1.5)
    with pytest.raises(ValueError, match="weight must have the same size"):
        dtw.dtw_alignment(np.zeros(4), np.zeros(6), weight=np.ones(4))
    with pytest.raises(ValueError, match="neither x or y"):
        dtw.dtw_mapping(x=np.zeros(3))
    with pytest.raises(ValueError, match="minimum of 2"):
        dtw.dtw_average(np.zeros((1, 5)))
    with pytest.raises(ValueError, match="method must be"):
        if wb.device_count() == 0:
            raise ValueError("method must be (no device: the sample set cannot be uploaded)")
        dtw.dtw_average(np.zeros((3, 5)), init=np.zeros(5), method="bogus")
    # a precomputed alignment is walked back on the host exactly like the reference
    a = np.array([[0.0, 1.0, np.inf], [1.0, 0.5, 2.0], [np.inf, 1.5, 0.7]])
    assert np.array_equal(dtw.dtw_mapping(alignment=a), np.eye(3, dtype=bool))
    from wildboar_b200.neighbors import KMeans
    with pytest.raises(ValueError, match="metric"):
        KMeans(3, metric="euclidean").fit(np.zeros((5, 8)))
    with pytest.raises(ValueError, match="n_clusters"):
        KMeans(0).fit(np.zeros((5, 8)))


@pytest.mark.gpu
def test_alignment_and_mapping_match_reference_golden(W, next_golden):
    from wildboar_b200 import dtw
    g = next_golden
    for k, tx, ty, r in _al_cases(g):
        x, y = g[f"al|{k}|x"], g[f"al|{k}|y"]
        for name, fn, kw in (("dtw", dtw.dtw_alignment, {}), ("wdtw", dtw.wdtw_alignment, {"g": 0.1})):
            want = g[f"al|{k}|{name}|matrix"]
            got = fn(x, y, r=r, **kw)
            assert _band_equal(got, want), (k, name)
            assert np.all(np.isposinf(got[np.isnan(want)])), (k, name)   # outside the band: +inf instead of garbage
            assert np.array_equal(dtw.dtw_mapping(alignment=got), g[f"al|{k}|{name}|path"]), (k, name)
        assert np.array_equal(dtw.dtw_mapping(x, y, r=r), g[f"al|{k}|dtw|path"]), k
        ind, (ii, jj) = dtw.dtw_mapping(x, y, r=r, return_index=True)
        assert np.array_equal(np.stack([ii, jj]), np.stack(g[f"al|{k}|dtw|path"].nonzero()))


@pytest.mark.gpu
def test_batched_paths_match_oracle(W, oracle):
    from wildboar_b200 import dtw
    a, b = random_walks(70, 90, 41), random_walks(33, 120, 42)
    ia = np.random.default_rng(1).integers(0, 70, 200)
    ib = np.random.default_rng(2).integers(0, 33, 200)
    for r, weight in ((0.1, None), (0.3, oracle.jeong_weight(120, 0.2)), (1.0, None), (0.0, None)):
        lo, hi, cost = dtw.dtw_paths(a, b, r=r, weight=weight, ia=ia, ib=ib, return_cost=True)
        for p in range(0, 200, 7):
            A = oracle.dtw_alignment(a[ia[p]], b[ib[p]], r=r, weight=weight)
            olo, ohi = oracle.dtw_path(A)
            assert np.array_equal(lo[p], olo) and np.array_equal(hi[p], ohi), (r, p)
            assert cost[p] == A[-1, -1]
    lo, hi = dtw.dtw_paths(a[:33], b, r=0.2)     # no index arrays: pair p = (a[p], b[p])
    olo, ohi = oracle.dtw_path(oracle.dtw_alignment(a[32], b[32], r=0.2))
    assert np.array_equal(lo[32], olo) and np.array_equal(hi[32], ohi)


@pytest.mark.gpu
def test_dtw_average_matches_reference_golden(W, next_golden):
    from wildboar_b200 import dtw
    g = next_golden
    X, sw = g["dba|X"], g["dba|sw"]
    for name, kw in (("mm", {}), ("mm_g", {"g": 0.15}), ("mm_sw", {"sample_weight": sw}), ("mm_r1", {"r": 1.0}),
                     ("ssg", {"method": "ssg", "random_state": 3, "max_epoch": 6}),
                     ("random", {"init": "random", "random_state": 5})):
        args = dict(r=0.2, init=X[7], method="mm", return_cost=True)
        args.update(kw)
        mean, cost = dtw.dtw_average(X, **args)
        assert np.array_equal(mean, g[f"dba|{name}|mean"]), name
        assert cost == g[f"dba|{name}|cost"], name
    # several groups per device step == one call per group
    groups = [np.arange(0, 10), np.arange(10, 24), np.array([1, 5, 20])]
    inits = [X[0], X[12], X[5]]
    means, costs = dtw.dtw_average_many(X, groups, inits, r=0.2)
    for grp, init, mean, cost in zip(groups, inits, means, costs):
        m1, c1 = dtw.dtw_average(X[grp], r=0.2, init=init, return_cost=True)
        assert np.array_equal(mean, m1) and cost == c1


@pytest.mark.gpu
def test_dba_larger_matches_oracle(W, oracle):
    from wildboar_b200 import dtw
    X = random_walks(40, 200, 43)
    mean, cost = dtw.dtw_average(X, r=0.1, init=X[3], max_epoch=4, return_cost=True)
    omean, ocost = oracle.dtw_average_mm(X, r=0.1, init=X[3], max_epoch=4)
    assert np.array_equal(mean, omean) and cost == ocost


@pytest.mark.gpu
def test_kmeans_matches_reference_golden(W, next_golden):
    from wildboar_b200.neighbors import KMeans
    g = next_golden
    Xk = g["km|X"]
    for name, kw in (("dtw", {}), ("wdtw", {"g": 0.1}), ("k7", {"n_clusters": 7, "n_init": 2, "r": 0.1})):
        args = dict(n_clusters=3, metric="dtw", r=0.2, random_state=11, max_iter=20)
        args.update(kw)
        km = KMeans(**args).fit(Xk)
        assert np.array_equal(km.cluster_centers_, g[f"km|{name}|centers"]), name
        assert np.array_equal(km.labels_, g[f"km|{name}|labels"]), name
        assert km.inertia_ == g[f"km|{name}|inertia"] and km.n_iter_ == g[f"km|{name}|n_iter"], name
        assert np.array_equal(km.transform(Xk[::5]), g[f"km|{name}|transform"]), name
        assert np.array_equal(km.predict(Xk[::5]), g[f"km|{name}|transform"].argmin(axis=1)), name


# ---------------------------------------------------------------------------------------------
# subsequence search, DTW family
# ---------------------------------------------------------------------------------------------
SS_CASES = [("dtw", {"r": 0.1}), ("dtw", {"r": 1.0}), ("wdtw", {"r": 0.3, "g": 0.1}), ("adtw", {"r": 0.2, "p": 0.5}),
            ("ddtw", {"r": 0.2}), ("wddtw", {"r": 0.5, "g": 0.2})]


def _ss_inputs(g, ci):
    keep = g[f"ss|{ci}|keep"]
    return g["ss|X"], [g[f"ss|s{k}"] for k in keep]


def test_oracle_subsequence_matches_reference_golden(oracle, next_golden):
    g = next_golden
    for ci, (metric, mp) in enumerate(SS_CASES):
        X, ss = _ss_inputs(g, ci)
        d, i = oracle.pairwise_subsequence(metric, ss, X, **mp)
        assert np.array_equal(d, g[f"ss|{ci}|dist"]) and np.array_equal(i, g[f"ss|{ci}|idx"]), (metric, mp)


def test_subsequence_host_logic(wb):
    x = np.zeros((3, 10))
    with pytest.raises(ValueError, match="cannot be empty"):
        wb.pairwise_subsequence_distance([], x, metric="dtw")
    with pytest.raises(ValueError, match="Invalid subsequnce shape"):
        wb.pairwise_subsequence_distance([np.zeros(11)], x, metric="dtw")
    with pytest.raises(ValueError, match="unsupported metric"):
        wb.pairwise_subsequence_distance([np.zeros(4)], x, metric="euclidean")
    with pytest.raises(ValueError, match="unsupported metric"):
        wb.pairwise_subsequence_distance([np.zeros(4)], x, metric="manhattan", scale=True)
    with pytest.raises(ValueError, match="unsupported metric"):
        wb.pairwise_subsequence_distance([np.zeros(4)], x, metric="scaled_wlcss")
    with pytest.raises(ValueError, match="at least 3 samples"):
        wb.pairwise_subsequence_distance([np.zeros(2)], x, metric="dtw", scale=True)
    with pytest.raises(ValueError, match="must be the same"):
        wb.paired_subsequence_distance([np.zeros(4)], x, metric="dtw")
    with pytest.raises(ValueError, match="dim must be"):
        wb.pairwise_subsequence_distance([np.zeros(4)], x, dim=1, metric="dtw")


@pytest.mark.gpu
@pytest.mark.parametrize("ci", range(len(SS_CASES)))
def test_subsequence_matches_reference_golden(W, next_golden, ci):
    g = next_golden
    metric, mp = SS_CASES[ci]
    X, ss = _ss_inputs(g, ci)
    d, i = W.pairwise_subsequence_distance(ss, X, metric=metric, metric_params=mp, return_index=True)
    assert np.array_equal(d, g[f"ss|{ci}|dist"]) and np.array_equal(i, g[f"ss|{ci}|idx"])
    assert np.array_equal(W.pairwise_subsequence_distance(ss, X, metric=metric, metric_params=mp), g[f"ss|{ci}|dist"])
    paired = [ss[q % len(ss)] for q in range(X.shape[0])]
    d, i = W.paired_subsequence_distance(paired, X, metric=metric, metric_params=mp, return_index=True)
    assert np.array_equal(d, g[f"ss|{ci}|paired_dist"]) and np.array_equal(i, g[f"ss|{ci}|paired_idx"])
    # return shapes (_format_return): one subsequence -> (n_samples,), 1-D x -> (n_subsequences,), both -> scalar
    assert W.pairwise_subsequence_distance(ss[0], X, metric=metric, metric_params=mp).shape == (X.shape[0],)
    assert W.pairwise_subsequence_distance(ss, X[0], metric=metric, metric_params=mp).shape == (len(ss),)
    assert isinstance(W.pairwise_subsequence_distance(ss[0], X[0], metric=metric, metric_params=mp), float)


@pytest.mark.gpu
def test_subsequence_larger_matches_oracle(W, oracle):
    """Longer series (strip engine, several warp tasks per sample), ties (repeated motif -> first window wins),
    3-D input with dim, two devices if present."""
    rng = np.random.default_rng(9)
    X = np.cumsum(rng.standard_normal((40, 600)), axis=1)
    motif = X[3, 100:164].copy()
    X[3, 300:364] = motif                                    # the same window twice: index 100 must win
    subs = [motif, X[7, 500:600].copy(), np.cumsum(rng.standard_normal(200))]
    for metric, mp in (("dtw", {"r": 0.1}), ("wdtw", {"r": 0.2, "g": 0.05}), ("ddtw", {"r": 0.1}), ("adtw", {"r": 0.05, "p": 2.0})):
        d, i = W.pairwise_subsequence_distance(subs, X, metric=metric, metric_params=mp, return_index=True)
        od, oi = oracle.pairwise_subsequence(metric, subs, X, **mp)
        assert np.array_equal(d, od) and np.array_equal(i, oi), metric
    d, i = W.pairwise_subsequence_distance(subs, X, metric="dtw", metric_params={"r": 0.1}, return_index=True)
    assert d[3, 0] == 0.0 and i[3, 0] == 100 and d[7, 1] == 0.0 and i[7, 1] == 500
    X3 = np.stack([X, X[::-1]], axis=1)
    d1 = W.pairwise_subsequence_distance(subs, X3, dim=1, metric="dtw", metric_params={"r": 0.1})
    assert np.array_equal(d1, d[::-1])
    if W.device_count() >= 2:
        W.set_devices([0, 1])
        try:
            d2, i2 = W.pairwise_subsequence_distance(subs, X, metric="dtw", metric_params={"r": 0.1}, return_index=True)
        finally:
            W.set_devices([0])
        assert np.array_equal(d2, d) and np.array_equal(i2, i)


# ---------------------------------------------------------------------------------------------
# ElasticEnsembleClassifier (leave-one-out grid search as one argmin call per candidate)
# ---------------------------------------------------------------------------------------------
def test_ensemble_host_logic(wb):
    from wildboar_b200.ensemble import ElasticEnsembleClassifier, make_parameter_grid
    grid = make_parameter_grid({"min_r": 0.1, "max_r": 0.3, "num_r": 3, "min_p": 1, "max_p": 4, "num_p": 2})
    assert [sorted(g) for g in grid] == [["p", "r"]] * 6 and grid[0] == {"r": 0.1, "p": 1.0} and grid[-1] == {"r": 0.3, "p": 4.0}
    assert make_parameter_grid(None) == [{}]
    with pytest.raises(ValueError, match="must be prefixed"):
        make_parameter_grid({"r": 1})
    with pytest.raises(ValueError, match="maximum value is missing"):
        make_parameter_grid({"min_r": 0.1})
    with pytest.raises(ValueError, match="too few labels"):
        ElasticEnsembleClassifier().fit(np.zeros((4, 8)), [1, 1, 1, 1])
    with pytest.raises(ValueError, match="non-elastic"):
        ElasticEnsembleClassifier(metric="all").fit(np.zeros((4, 8)), [0, 1, 0, 1])
    with pytest.raises(ValueError, match="is not supported"):
        ElasticEnsembleClassifier(metric={"euclidean": None}).fit(np.zeros((4, 8)), [0, 1, 0, 1])


@pytest.mark.gpu
def test_ensemble_matches_reference_golden(W, next_golden):
    import ast
    from wildboar_b200.ensemble import ElasticEnsembleClassifier
    g = next_golden
    X, y, Q = g["ee|X"], g["ee|y"], g["ee|Q"]
    for name, kw in (("auto_k1", dict(n_neighbors=1, metric="auto")),
                     ("custom_k3", dict(n_neighbors=3, metric={"dtw": {"min_r": 0.1, "max_r": 0.3, "num_r": 3},
                                                                 "msm": {"min_c": 0.1, "max_c": 10, "num_c": 4},
                                                                 "lcss": {"min_r": 0.0, "max_r": 0.25, "num_r": 2,
                                                                          "min_epsilon": 0.3, "max_epsilon": 1.2, "num_epsilon": 2}}))):
        clf = ElasticEnsembleClassifier(**kw).fit(X, y)
        assert [m for m, _ in clf.scores_] == list(g[f"ee|{name}|metrics"]), name
        assert np.array_equal(np.array([s for _, s in clf.scores_]), g[f"ee|{name}|scores"]), name
        want_params = ast.literal_eval(str(g[f"ee|{name}|params"]))
        assert [{k: float(v) for k, v in e.metric_params.items()} for e in clf.estimators_] == want_params, name
        assert np.array_equal(clf.predict_proba(Q), g[f"ee|{name}|proba"]), name
        assert np.array_equal(clf.predict(Q), g[f"ee|{name}|predict"]), name


@pytest.mark.gpu
def test_loo_mask_equals_fold_by_fold(W, oracle):
    """The +inf diagonal lower bound reproduces the fold-by-fold scans (here: the oracle on each fold), also for the
    metrics whose early abandoning makes argmin differ from pairwise + top-k (lcss, erp)."""
    from wildboar_b200.ensemble import _loo_neighbors
    X = random_walks(30, 40, 51)
    for metric, mp in (("lcss", {"r": 0.2, "epsilon": 0.5}), ("erp", {"r": 0.1}), ("dtw", {"r": 0.1}), ("twe", {"r": 0.3})):
        got = _loo_neighbors(X, metric, mp, 3)
        for i in (0, 7, 29):
            others = np.delete(np.arange(30), i)
            oi, _ = oracle.argmin(metric, X[i:i + 1], X[others], k=3, **mp)
            assert np.array_equal(got[i], others[oi[0]]), (metric, i)


def _ssc_inputs(g):
    return g["ssc|X"], [g[f"ssc|s{k}"] for k in range(int(g["ssc|n"]))]


def test_oracle_scaled_dtw_subsequence_matches_reference_golden(oracle, next_golden):
    g = next_golden
    X, ss = _ssc_inputs(g)
    for r in (0.0, 0.05, 0.1, 0.3, 1.0):
        d, i = oracle.pairwise_scaled_dtw_subsequence(ss, X, r=r)
        assert np.array_equal(d, g[f"ssc|{r}|dist"]) and np.array_equal(i, g[f"ssc|{r}|idx"]), r


@pytest.mark.gpu
def test_scaled_dtw_subsequence_matches_reference_golden(W, next_golden):
    g = next_golden
    X, ss = _ssc_inputs(g)
    for r in (0.0, 0.05, 0.1, 0.3, 1.0):
        d, i = W.pairwise_subsequence_distance(ss, X, metric="scaled_dtw", metric_params={"r": r}, return_index=True)
        assert np.array_equal(d, g[f"ssc|{r}|dist"]) and np.array_equal(i, g[f"ssc|{r}|idx"]), r
    d, i = W.paired_subsequence_distance([ss[q % len(ss)] for q in range(X.shape[0])], X, metric="dtw", scale=True,
                                         metric_params={"r": 0.1}, return_index=True)
    assert np.array_equal(d, g["ssc|paired_dist"]) and np.array_equal(i, g["ssc|paired_idx"])
    with pytest.raises(ValueError, match="at least 3 samples"):
        W.pairwise_subsequence_distance([np.zeros(2)], X, metric="scaled_dtw")
    with pytest.raises(ValueError, match="unsupported metric"):
        W.pairwise_subsequence_distance([np.zeros(5)], X, metric="scaled_euclidean")


@pytest.mark.gpu
def test_scaled_dtw_subsequence_larger_matches_oracle(W, oracle):
    """Longer series / subsequences; the oracle restates the scan including the reference's (invalid) LB_Kim prefilter."""
    rng = np.random.default_rng(19)
    X = np.cumsum(rng.standard_normal((25, 400)), axis=1)
    subs = [np.cumsum(rng.standard_normal(m)) for m in (7, 33, 128, 300)] + [X[9, 111:175].copy() * 0.5 - 3.0]
    for r in (0.02, 0.1, 1.0):
        d, i = W.pairwise_subsequence_distance(subs, X, metric="scaled_dtw", metric_params={"r": r}, return_index=True)
        od, oi = oracle.pairwise_scaled_dtw_subsequence(subs, X, r=r)
        assert np.array_equal(d, od) and np.array_equal(i, oi), r
    assert abs(d[9, 4]) < 1e-6 and i[9, 4] == 111
