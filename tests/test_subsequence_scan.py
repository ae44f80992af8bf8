"""SURVEY 8f-4, second half: subsequence search with lcss / erp / edr / msm / twe and the generic scaled_<metric>
metrics (ScaledSubsequenceMetricWrap), where the reference's early abandoning decides which window is reported.

CPU tests pin the oracle restatement to golden vectors generated from the unmodified reference
(tests/golden/make_golden_scan.py); `-m gpu` tests compare the CUDA path (public API -> ctypes -> C ABI) with those
vectors and, at larger sizes, with the oracle -- bit for bit, distances and window indices.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from make_golden_scan import SC_CASES, SW_CASES  # noqa: E402


@pytest.fixture(scope="module")
def scan_golden():
    with np.load(os.path.join(ROOT, "tests", "golden", "scan_golden.npz")) as z:
        return {k: z[k] for k in z.files}


@pytest.fixture(scope="module")
def W(wb):
    assert wb.device_count() >= 1, "no CUDA device: the product has no CPU fallback"
    wb.set_devices([0])
    return wb


def _inputs(g, tag, ci):
    keep = g[f"{tag}|{ci}|keep"]
    return g["X"], [g[f"s{k}"] for k in keep]


# ---------------------------------------------------------------------------------------------
# CPU: oracle == reference
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("ci", range(len(SC_CASES)))
def test_oracle_scan_matches_reference_golden(oracle, scan_golden, ci):
    metric, mp = SC_CASES[ci]
    X, ss = _inputs(scan_golden, "sc", ci)
    d, i = oracle.pairwise_subsequence(metric, ss, X, **mp)
    assert np.array_equal(d, scan_golden[f"sc|{ci}|dist"]) and np.array_equal(i, scan_golden[f"sc|{ci}|idx"])


@pytest.mark.parametrize("ci", range(len(SW_CASES)))
def test_oracle_scaled_wrap_matches_reference_golden(oracle, scan_golden, ci):
    metric, mp = SW_CASES[ci]
    X, ss = _inputs(scan_golden, "sw", ci)
    d, i = oracle.pairwise_scaled_subsequence(metric, ss, X, **mp)
    assert np.array_equal(d, scan_golden[f"sw|{ci}|dist"]) and np.array_equal(i, scan_golden[f"sw|{ci}|idx"])


def test_abandoning_decides_the_result(oracle, scan_golden):
    """Why the scan has to be replayed: for these metrics the reference's answer is NOT the plain minimum over the windows."""
    X = scan_golden["X"]
    differs = 0
    for ci, (metric, mp) in enumerate(SC_CASES):
        _, ss = _inputs(scan_golden, "sc", ci)
        for k, s in enumerate(ss):
            m = len(s)
            kw = dict(mp)
            if metric == "edr" and "epsilon" not in kw:
                kw["epsilon"] = oracle._subsequence_mean_std(s)[1] / 4.0
            for i in range(X.shape[0]):
                wins = np.stack([X[i, w:w + m] for w in range(X.shape[1] - m + 1)])
                plain = oracle.pairwise(metric, s, wins, **kw)[0].min()
                differs += plain != scan_golden[f"sc|{ci}|dist"][i, k]
    assert differs > 0


def test_scan_host_logic(wb):
    x = np.zeros((3, 10))
    for metric in ("lcss", "erp", "edr", "msm", "twe", "scaled_adtw", "scaled_twe", "scaled_edr"):
        with pytest.raises(ValueError, match="Invalid subsequnce shape"):
            wb.pairwise_subsequence_distance([np.zeros(11)], x, metric=metric)
    with pytest.raises(TypeError):
        wb.pairwise_subsequence_distance([np.zeros(4)], x, metric="msm", metric_params={"penalty": 1.0})
    with pytest.raises(ValueError):
        wb.pairwise_subsequence_distance([np.zeros(4)], x, metric="scaled_lcss", metric_params={"epsilon": -1.0})
    from wildboar_b200.subsequence import _check_subsequence_metric
    assert _check_subsequence_metric("msm", True) == ("msm", True)
    assert _check_subsequence_metric("scaled_twe", False) == ("twe", True)
    assert _check_subsequence_metric("edr", False) == ("edr", False)


# ---------------------------------------------------------------------------------------------
# GPU
# ---------------------------------------------------------------------------------------------
def _check_golden(W, g, tag, ci, metric, mp, name):
    X, ss = _inputs(g, tag, ci)
    d, i = W.pairwise_subsequence_distance(ss, X, metric=name, metric_params=mp, return_index=True)
    bad = np.argwhere((d != g[f"{tag}|{ci}|dist"]) | (i != g[f"{tag}|{ci}|idx"]))
    assert len(bad) == 0, (name, mp, bad[:4], d[tuple(bad[0])], g[f"{tag}|{ci}|dist"][tuple(bad[0])], i[tuple(bad[0])],
                           g[f"{tag}|{ci}|idx"][tuple(bad[0])])
    paired = [ss[q % len(ss)] for q in range(X.shape[0])]
    d, i = W.paired_subsequence_distance(paired, X, metric=name, metric_params=mp, return_index=True)
    assert np.array_equal(d, g[f"{tag}|{ci}|paired_dist"]) and np.array_equal(i, g[f"{tag}|{ci}|paired_idx"]), (name, mp)
    assert W.pairwise_subsequence_distance(ss[0], X, metric=name, metric_params=mp).shape == (X.shape[0],)
    assert isinstance(W.pairwise_subsequence_distance(ss[0], X[0], metric=name, metric_params=mp), float)


@pytest.mark.gpu
@pytest.mark.parametrize("ci", range(len(SC_CASES)))
def test_scan_matches_reference_golden(W, scan_golden, ci):
    metric, mp = SC_CASES[ci]
    _check_golden(W, scan_golden, "sc", ci, metric, mp, metric)


@pytest.mark.gpu
@pytest.mark.parametrize("ci", range(len(SW_CASES)))
def test_scaled_wrap_matches_reference_golden(W, scan_golden, ci):
    metric, mp = SW_CASES[ci]
    _check_golden(W, scan_golden, "sw", ci, metric, mp, "scaled_" + metric)
    if ci == 0:  # scale=True selects the scaled metric like the reference's _infer_scaled_metric
        X, ss = _inputs(scan_golden, "sw", ci)
        d = W.pairwise_subsequence_distance(ss, X, metric=metric, scale=True, metric_params=mp)
        assert np.array_equal(d, scan_golden[f"sw|{ci}|dist"])


@pytest.mark.gpu
def test_scan_larger_matches_oracle(W, oracle):
    """Longer series (several warp tasks per sample, several samples per replay block), a repeated motif (first window wins),
    more than one pass over the samples for the materialised scaled windows, two devices if present."""
    rng = np.random.default_rng(23)
    X = np.cumsum(rng.standard_normal((37, 300)), axis=1)
    motif = X[3, 50:90].copy()
    X[3, 200:240] = motif
    subs = [motif, np.cumsum(rng.standard_normal(100)), X[11, 260:300].copy(), np.cumsum(rng.standard_normal(9))]
    for metric, mp in (("lcss", {"r": 0.1, "epsilon": 0.8}), ("erp", {"r": 0.1}), ("edr", {"r": 0.1}), ("msm", {"r": 0.1}),
                       ("twe", {"r": 0.1})):
        d, i = W.pairwise_subsequence_distance(subs, X, metric=metric, metric_params=mp, return_index=True)
        od, oi = oracle.pairwise_subsequence(metric, subs, X, **mp)
        assert np.array_equal(d, od) and np.array_equal(i, oi), metric
        if metric != "lcss":
            assert d[3, 0] == 0.0 and i[3, 0] == 50 and d[11, 2] == 0.0 and i[11, 2] == 260
    os.environ["WILDBOAR_CUDA_SCAN_WINDOW_BUDGET"] = str(300 * 100 * 5)  # forces several passes over the 37 samples
    try:
        for metric, mp in (("adtw", {"r": 0.1, "p": 0.2}), ("wdtw", {"r": 0.1}), ("ddtw", {"r": 0.1}), ("wddtw", {"r": 0.2}),
                           ("lcss", {"r": 0.1, "epsilon": 0.3}), ("erp", {"r": 0.1}), ("edr", {"r": 0.1}), ("msm", {"r": 0.1}),
                           ("twe", {"r": 0.1})):
            d, i = W.pairwise_subsequence_distance(subs, X, metric="scaled_" + metric, metric_params=mp, return_index=True)
            od, oi = oracle.pairwise_scaled_subsequence(metric, subs, X, **mp)
            assert np.array_equal(d, od) and np.array_equal(i, oi), metric
    finally:
        del os.environ["WILDBOAR_CUDA_SCAN_WINDOW_BUDGET"]
    # the head-then-abandon scheme forced on for every metric that has row minima (default: msm / twe / erp unscaled only) and off
    for flag in ("1", "0"):
        os.environ["WILDBOAR_CUDA_SCAN_ABANDON"] = flag
        try:
            for name, metric, mp, scaled in (("edr", "edr", {"r": 0.1}, False), ("scaled_msm", "msm", {"r": 0.1}, True),
                                             ("scaled_edr", "edr", {"r": 0.2}, True), ("twe", "twe", {"r": 0.1}, False),
                                             ("scaled_erp", "erp", {"r": 0.1}, True)):
                d, i = W.pairwise_subsequence_distance(subs, X, metric=name, metric_params=mp, return_index=True)
                od, oi = (oracle.pairwise_scaled_subsequence if scaled else oracle.pairwise_subsequence)(metric, subs, X, **mp)
                assert np.array_equal(d, od) and np.array_equal(i, oi), (name, flag)
        finally:
            del os.environ["WILDBOAR_CUDA_SCAN_ABANDON"]
    if W.device_count() >= 2:
        W.set_devices([0, 1])
        try:
            d2, i2 = W.pairwise_subsequence_distance(subs, X, metric="scaled_msm", metric_params={"r": 0.1}, return_index=True)
        finally:
            W.set_devices([0])
        od, oi = oracle.pairwise_scaled_subsequence("msm", subs, X, r=0.1)
        assert np.array_equal(d2, od) and np.array_equal(i2, oi)


@pytest.mark.gpu
def test_scan_negative_adtw_penalty(W, oracle):
    """adtw with a negative penalty: row minima may decrease, so the abandoning of the scan matters here too."""
    rng = np.random.default_rng(5)
    X = np.cumsum(rng.standard_normal((6, 80)), axis=1)
    subs = [np.cumsum(rng.standard_normal(m)) for m in (10, 25)]
    mp = {"r": 0.3, "p": -0.05}
    d, i = W.pairwise_subsequence_distance(subs, X, metric="adtw", metric_params=mp, return_index=True)
    od, oi = oracle.pairwise_subsequence("adtw", subs, X, **mp)
    assert np.array_equal(d, od, equal_nan=True) and np.array_equal(i, oi)
    d, i = W.pairwise_subsequence_distance(subs, X, metric="scaled_adtw", metric_params=mp, return_index=True)
    od, oi = oracle.pairwise_scaled_subsequence("adtw", subs, X, **mp)
    assert np.array_equal(d, od, equal_nan=True) and np.array_equal(i, oi)


# ---------------------------------------------------------------------------------------------
# subsequence_match / paired_subsequence_match / distance_profile
# ---------------------------------------------------------------------------------------------
from make_golden_scan import SM_CASES, SM_SUBS, pad  # noqa: E402


def _same(a, b):
    return a.shape == b.shape and np.array_equal(a, b, equal_nan=True)


def _dense_from_padded(idx, dist, nw):
    out = np.full((idx.shape[0], nw), np.nan)
    for i in range(idx.shape[0]):
        sel = idx[i] >= 0
        out[i, idx[i][sel]] = dist[i][sel]
    return out


def _view_mean_std(s):
    ex, ex2 = np.cumsum(s)[-1], np.cumsum(s * s)[-1]
    mean = ex / len(s)
    var = ex2 / len(s) - mean * mean
    return float(mean), float(np.sqrt(var) if var > 1e-13 else 0.0)


@pytest.mark.parametrize("scaled", [False, True])
@pytest.mark.parametrize("ci", range(len(SM_CASES)))
def test_oracle_matches_and_profile_match_reference_golden(oracle, scan_golden, ci, scaled):
    g = scan_golden
    metric, mp = SM_CASES[ci]
    name = ("scaled_" if scaled else "") + metric
    X = g["X"]
    for k in SM_SUBS:
        s = g[f"s{k}"]
        for ti in range(2):
            key = f"sm|{name}|{ci}|{k}|thr{ti}"
            want = _dense_from_padded(g[key + "|idx"], g[key + "|dist"], X.shape[1] - len(s) + 1)
            got = oracle.subsequence_matches(metric, s, X, float(g[key]), scaled, **mp)
            assert _same(got, want), (name, k, ti)
    n = X.shape[0]
    Y = np.stack([X[(q + 1) % n, 7 + q:22 + q] for q in range(n)])
    got = oracle.subsequence_matches(metric, Y, X, np.inf, scaled, mean_std=_view_mean_std, **mp)
    assert _same(got, g[f"sm|{name}|{ci}|profile"]), name


def test_match_host_logic(wb):
    x = np.zeros((3, 10))
    with pytest.raises(ValueError, match="single subsequence"):
        wb.subsequence_match([np.zeros(3), np.zeros(3)], x, metric="dtw")
    with pytest.raises(ValueError, match="Invalid subsequnce shape"):
        wb.subsequence_match(np.zeros(11), x, metric="msm")
    with pytest.raises(TypeError, match="threshold must be"):
        wb.subsequence_match(np.zeros(4), x, threshold=object(), metric="dtw")
    with pytest.raises(ValueError, match="must match the number of samples"):
        wb.subsequence_match(np.zeros(4), x, threshold=[1.0, 2.0], metric="dtw")
    with pytest.raises(ValueError, match="unsupported metric"):
        wb.subsequence_match(np.zeros(4), x, threshold=1.0, metric="euclidean")
    with pytest.raises(ValueError, match="must be the same"):
        wb.paired_subsequence_match([np.zeros(4)], x, metric="dtw")
    with pytest.raises(ValueError, match="dilation must be"):
        wb.distance_profile(np.zeros(3), x, dilation=0, metric="dtw")
    with pytest.raises(ValueError, match="larger than input"):
        wb.distance_profile(np.zeros((3, 11)), x, metric="dtw")
    with pytest.raises(ValueError, match="same number of samples"):
        wb.distance_profile(np.zeros((2, 4)), x, metric="dtw")
    from wildboar_b200.subsequence import _keep_nontrivial, _rows_to_matches
    idx, dist = _rows_to_matches(np.array([[np.nan, 1.0, 0.5, np.nan], [np.nan] * 4]))
    assert np.array_equal(idx[0], [1, 2]) and np.array_equal(dist[0], [1.0, 0.5]) and idx[1] is None and dist[1] is None
    keep = _keep_nontrivial(2)(0, np.array([0, 1, 2, 5]), np.array([0.3, 0.1, 0.2, 0.9]))
    assert np.array_equal(keep, [False, True, False, True])


@pytest.mark.gpu
@pytest.mark.parametrize("scaled", [False, True])
@pytest.mark.parametrize("ci", range(len(SM_CASES)))
def test_subsequence_match_and_profile_match_reference_golden(W, scan_golden, ci, scaled):
    g = scan_golden
    metric, mp = SM_CASES[ci]
    name = ("scaled_" if scaled else "") + metric
    X = g["X"]
    n = X.shape[0]
    for k in SM_SUBS:
        s = g[f"s{k}"]
        key = f"sm|{name}|{ci}|{k}"
        for ti in range(2):
            i_, d_ = W.subsequence_match(s, X, threshold=float(g[f"{key}|thr{ti}"]), metric=name, metric_params=mp, return_distance=True)
            assert _same(pad(i_, -1, np.int64), g[f"{key}|thr{ti}|idx"]) and _same(pad(d_, np.nan, float), g[f"{key}|thr{ti}|dist"]), (name, k, ti)
        i_, d_ = W.subsequence_match(s, X, metric=name, metric_params=mp, return_distance=True)
        assert _same(pad(i_, -1, np.int64), g[f"{key}|top|idx"]) and _same(pad(d_, np.nan, float), g[f"{key}|top|dist"]), (name, k, "top")
        i_, d_ = W.subsequence_match(s, X, threshold="auto", metric=name, metric_params=mp, return_distance=True)
        assert _same(pad(i_, -1, np.int64), g[f"{key}|auto|idx"]) and _same(pad(d_, np.nan, float), g[f"{key}|auto|dist"]), (name, k, "auto")
        thr = float(g[f"{key}|thr0"])
        i_, d_ = W.subsequence_match(s, X, threshold=thr, exclude=0.5, max_matches=4, metric=name, metric_params=mp, return_distance=True)
        assert _same(pad(i_, -1, np.int64), g[f"{key}|excl|idx"]) and _same(pad(d_, np.nan, float), g[f"{key}|excl|dist"]), (name, k, "excl")
    paired = [g[f"s{SM_SUBS[q % len(SM_SUBS)]}"] for q in range(n)]
    i_, d_ = W.paired_subsequence_match(paired, X, metric=name, metric_params=mp, return_distance=True, max_matches=5)
    assert _same(pad(i_, -1, np.int64), g[f"sm|{name}|{ci}|paired|idx"]) and _same(pad(d_, np.nan, float), g[f"sm|{name}|{ci}|paired|dist"]), name
    Y = np.stack([X[(q + 1) % n, 7 + q:22 + q] for q in range(n)])
    assert _same(W.distance_profile(Y, X, metric=name, metric_params=mp), g[f"sm|{name}|{ci}|profile"]), name
    # scale=True spelling and the single-series forms
    if scaled:
        assert _same(W.distance_profile(Y, X, metric=metric, scale=True, metric_params=mp), g[f"sm|{name}|{ci}|profile"])
    assert W.distance_profile(Y[0], X[0], metric=name, metric_params=mp).shape == (X.shape[1] - 15 + 1,)


@pytest.mark.gpu
def test_profile_larger_matches_oracle(W, oracle):
    """Longer series, several passes over the samples (scaled windows), per-sample thresholds, two devices if present."""
    rng = np.random.default_rng(31)
    X = np.cumsum(rng.standard_normal((33, 260)), axis=1)
    s = X[5, 40:100].copy()
    os.environ["WILDBOAR_CUDA_SCAN_WINDOW_BUDGET"] = str(201 * 60 * 7)
    try:
        for metric, mp in (("dtw", {"r": 0.1}), ("ddtw", {"r": 0.1}), ("msm", {"r": 0.1}), ("twe", {"r": 0.05}), ("lcss", {"r": 0.1, "epsilon": 0.4}),
                           ("erp", {"r": 0.1}), ("edr", {"r": 0.1})):
            for scaled in (False, True):
                name = ("scaled_" if scaled else "") + metric
                full = oracle.subsequence_matches(metric, s, X, np.inf, scaled, **mp)
                thr = float(np.nanquantile(full, 0.2))
                want = oracle.subsequence_matches(metric, s, X, thr, scaled, **mp)
                i_, d_ = W.subsequence_match(s, X, threshold=thr, metric=name, metric_params=mp, return_distance=True)
                got = _dense_from_padded(pad(i_, -1, np.int64), pad(d_, np.nan, float), full.shape[1])
                assert _same(got, want), name
                Y = np.stack([X[(q + 3) % 33, q:q + 60] for q in range(33)])
                want = oracle.subsequence_matches(metric, Y, X, np.inf, scaled, mean_std=_view_mean_std, **mp)
                assert _same(W.distance_profile(Y, X, metric=name, metric_params=mp), want), name
    finally:
        del os.environ["WILDBOAR_CUDA_SCAN_WINDOW_BUDGET"]
    thr = np.linspace(2.0, 12.0, 33)
    i_, d_ = W.subsequence_match(s, X, threshold=thr, metric="dtw", metric_params={"r": 0.1}, return_distance=True)
    full = oracle.subsequence_matches("dtw", s, X, np.inf, False, r=0.1)
    want = np.where(full <= thr[:, None], full, np.nan)
    assert _same(_dense_from_padded(pad(i_, -1, np.int64), pad(d_, np.nan, float), full.shape[1]), want)
    assert i_[5][np.argmin(d_[5])] == 40 and d_[5].min() == 0.0
    if W.device_count() >= 2:
        W.set_devices([0, 1])
        try:
            dp2 = W.distance_profile(np.stack([s] * 33), X, metric="scaled_msm", metric_params={"r": 0.1})
        finally:
            W.set_devices([0])
        assert _same(dp2, oracle.subsequence_matches("msm", np.stack([s] * 33), X, np.inf, True, mean_std=_view_mean_std, r=0.1))


# ---------------------------------------------------------------------------------------------
# argmin_subsequence_distance
# ---------------------------------------------------------------------------------------------
from make_golden_scan import AS_CASES  # noqa: E402


def _as_inputs(g):
    X = g["X"]
    n = X.shape[0]
    Y = np.stack([X[(q + 1) % n, 7 + q:22 + q] for q in range(n)])
    ragged = [g[f"s{SM_SUBS[q % len(SM_SUBS)]}"] for q in range(n)]
    return X, Y, ragged


@pytest.mark.parametrize("ci", range(len(AS_CASES)))
def test_oracle_argmin_subsequence_matches_reference_golden(oracle, scan_golden, ci):
    g = scan_golden
    metric, mp = AS_CASES[ci]
    X, Y, ragged = _as_inputs(g)
    for scale in (False, True):
        for k in (1, 4):
            i_, d_ = oracle.argmin_subsequence(metric, list(Y), X, k=k, scaled=scale, **mp)
            assert np.array_equal(i_, g[f"as|{ci}|{int(scale)}|{k}|idx"]) and np.array_equal(d_, g[f"as|{ci}|{int(scale)}|{k}|dist"]), (metric, scale, k)
        i_, d_ = oracle.argmin_subsequence(metric, ragged, X, k=3, scaled=scale, **mp)
        assert np.array_equal(i_, g[f"as|{ci}|{int(scale)}|ragged|idx"]) and np.array_equal(d_, g[f"as|{ci}|{int(scale)}|ragged|dist"])


def test_argmin_subsequence_host_logic(wb):
    x = np.zeros((3, 10))
    with pytest.raises(ValueError, match="same number of samples"):
        wb.argmin_subsequence_distance(np.zeros((2, 4)), x, metric="dtw")
    with pytest.raises(ValueError, match="longest subsequence"):
        wb.argmin_subsequence_distance(np.zeros((3, 11)), x, metric="dtw")
    with pytest.raises(ValueError, match="k must be less"):
        wb.argmin_subsequence_distance(np.zeros((3, 4)), x, k=8, metric="dtw")
    with pytest.raises(ValueError, match="1d-array"):
        wb.argmin_subsequence_distance([np.zeros((2, 2))] * 3, x, metric="dtw")
    with pytest.raises(ValueError):
        wb.argmin_subsequence_distance(np.zeros((3, 4)), x, metric="euclidean")


@pytest.mark.gpu
@pytest.mark.parametrize("ci", range(len(AS_CASES)))
def test_argmin_subsequence_matches_reference_golden(W, scan_golden, ci):
    g = scan_golden
    metric, mp = AS_CASES[ci]
    X, Y, ragged = _as_inputs(g)
    for scale in (False, True):
        for k in (1, 4):
            i_, d_ = W.argmin_subsequence_distance(Y, X, k=k, metric=metric, metric_params=mp, scale=scale, return_distance=True)
            assert np.array_equal(i_, g[f"as|{ci}|{int(scale)}|{k}|idx"]) and np.array_equal(d_, g[f"as|{ci}|{int(scale)}|{k}|dist"]), (metric, scale, k)
        i_, d_ = W.argmin_subsequence_distance(ragged, X, k=3, metric=metric, metric_params=mp, scale=scale, return_distance=True)
        assert np.array_equal(i_, g[f"as|{ci}|{int(scale)}|ragged|idx"]) and np.array_equal(d_, g[f"as|{ci}|{int(scale)}|ragged|dist"])
    assert np.array_equal(W.argmin_subsequence_distance(Y, X, k=4, metric="scaled_" + metric, metric_params=mp), g[f"as|{ci}|1|4|idx"])


@pytest.mark.gpu
def test_argmin_subsequence_larger_matches_oracle(W, oracle):
    rng = np.random.default_rng(41)
    X = np.cumsum(rng.standard_normal((45, 240)), axis=1)
    Y = np.stack([X[(q + 2) % 45, q:q + 50] for q in range(45)])
    os.environ["WILDBOAR_CUDA_SCAN_WINDOW_BUDGET"] = str(191 * 50 * 9)
    try:
        for metric, mp in (("dtw", {"r": 0.1}), ("ddtw", {"r": 0.1}), ("msm", {"r": 0.1}), ("twe", {"r": 0.05}), ("erp", {"r": 0.1}),
                           ("edr", {"r": 0.1}), ("lcss", {"r": 0.1, "epsilon": 0.4})):
            for scale in (False, True):
                i_, d_ = W.argmin_subsequence_distance(Y, X, k=5, metric=metric, metric_params=mp, scale=scale, return_distance=True)
                oi, od = oracle.argmin_subsequence(metric, list(Y), X, k=5, scaled=scale, **mp)
                assert np.array_equal(i_, oi) and np.array_equal(d_, od), (metric, scale)
    finally:
        del os.environ["WILDBOAR_CUDA_SCAN_WINDOW_BUDGET"]


@pytest.mark.gpu
def test_subsequence_family_edge_shapes(W, oracle):
    """One window (m == T), one sample, m == 1, 3-D input with dim (strided rows), float32 / integer input, many lengths at once."""
    rng = np.random.default_rng(77)
    X3 = np.cumsum(rng.standard_normal((5, 3, 30)), axis=2)
    X = np.ascontiguousarray(X3[:, 1, :])
    full = [x.copy() for x in X]                       # m == T: a single window per sample
    for metric, mp in (("msm", {"r": 0.2}), ("scaled_twe", {"r": 0.2}), ("dtw", {"r": 0.2}), ("scaled_erp", {"r": 0.2})):
        scaled = metric.startswith("scaled_")
        base = metric[7:] if scaled else metric
        odist = oracle.pairwise_scaled_subsequence if scaled else oracle.pairwise_subsequence
        d, i = W.pairwise_subsequence_distance(full, X3, dim=1, metric=metric, metric_params=mp, return_index=True)
        od, oi = odist(base, full, X, **mp)
        assert np.array_equal(d, od) and np.array_equal(i, oi) and np.all(i == 0), metric
        # one sample, subsequences of every length 1 .. T in one call (one DP launch per length group)
        subs = [X[2, :m].copy() for m in range(1 if base != "dtw" or not scaled else 3, 31)]
        d, i = W.pairwise_subsequence_distance(subs, X[0], metric=metric, metric_params=mp, return_index=True)
        od, oi = odist(base, subs, X[:1], **mp)
        assert np.array_equal(d, od[0]) and np.array_equal(i, oi[0]), metric
        # profile / matches / argmin with a single window
        # the reference's undilated profile IGNORES dim and reads dimension 0 of x (&self.x[i, 0, 0], CD:1690-1699;
        # checked against the live reference: distance_profile(Y, X3, dim=1) == distance_profile(Y, X3[:, 0]))
        dp = W.distance_profile(X, X3, dim=1, metric=metric, metric_params=mp)
        want = oracle.subsequence_matches(base, X, np.ascontiguousarray(X3[:, 0, :]), np.inf, scaled, mean_std=_view_mean_std, **mp)
        assert _same(np.atleast_1d(dp), want[:, 0]), metric
        with pytest.raises(ValueError, match="dimensions of y"):
            W.distance_profile(X[:, :9], X3, dim=1, dilation=2, metric=metric, metric_params=mp)
        ai, ad = W.argmin_subsequence_distance(X[:, :29], X3, dim=1, k=2, metric=base, scale=scaled, metric_params=mp, return_distance=True)
        oi2, od2 = oracle.argmin_subsequence(base, list(X[:, :29]), X, k=2, scaled=scaled, **mp)
        assert np.array_equal(ai, oi2) and np.array_equal(ad, od2), metric
    # float32 and integer inputs are converted like the reference converts them (check_array(dtype=float))
    Xi = np.round(X * 4).astype(np.int64)
    d = W.pairwise_subsequence_distance([Xi[0, 3:12]], Xi, metric="edr", metric_params={"r": 0.3})
    assert np.array_equal(d, oracle.pairwise_subsequence("edr", [Xi[0, 3:12].astype(float)], Xi.astype(float), r=0.3)[0][:, 0])
    Xf = X.astype(np.float32)
    d = W.pairwise_subsequence_distance([Xf[0, 3:12]], Xf, metric="lcss", metric_params={"r": 0.3})
    assert np.array_equal(d, oracle.pairwise_subsequence("lcss", [Xf[0, 3:12].astype(float)], Xf.astype(float), r=0.3)[0][:, 0])
    with pytest.raises(ValueError):
        W.pairwise_subsequence_distance([np.array([1.0, np.nan, 2.0])], X, metric="msm")


def test_match_post_filters_equal_the_reference_helpers(wb):
    """CPU: the jagged-list post-filters of subsequence_match (exclusion zone, max_matches, threshold functions) against the
    reference's own helpers (_distance.py:376-511) on random match lists, None entries included."""
    from oracle import ref
    if ref.load() is None:
        pytest.skip("oracle/_ref not built")
    from wildboar.distance import _distance as RD
    from wildboar_b200 import subsequence as S
    rng = np.random.default_rng(3)

    def same(a, b):
        return len(a) == len(b) and all((p is None and q is None) or (p is not None and q is not None and np.array_equal(p, q))
                                        for p, q in zip(a, b))
    for trial in range(40):
        idx, dist = [], []
        for _ in range(6):
            n = int(rng.integers(0, 15))
            if n == 0:
                idx.append(None); dist.append(None)
            else:
                idx.append(np.sort(rng.choice(60, n, replace=False)).astype(np.intp))
                dist.append(np.round(rng.random(n), 1))          # ties on purpose
        ex = int(rng.integers(1, 9))
        ri, rd = RD._exclude_trivial_matches(idx, dist, ex)
        oi, od_ = S._filter(idx, dist, S._keep_nontrivial(ex))
        assert same(ri, oi) and same(rd, od_)
        mm = int(rng.integers(1, 7))
        ri, rd = RD._filter_by_max_matches(idx, dist, mm)
        oi, od_ = S._filter(idx, dist, lambda _, __, d: np.argsort(d)[:mm])
        assert same(ri, oi) and same(rd, od_)
        fn = RD._THRESHOLD["auto"]
        ri, rd = RD._filter_by_max_dist(idx, dist, lambda _, d: d <= fn(d))
        thr, _, max_dist = S._resolve_threshold("auto", None, 6, allow_array=True)
        oi, od_ = S._filter(idx, dist, lambda i, _, d: max_dist(i, d))
        assert np.isinf(thr) and same(ri, oi) and same(rd, od_)
    assert S._resolve_threshold(None, None, 3, True)[:2] == (np.inf, 10)
    assert S._resolve_threshold(0.5, None, 3, True) == (0.5, None, None)


# ---------------------------------------------------------------------------------------------
# distance_profile with dilation / padding
# ---------------------------------------------------------------------------------------------
from make_golden_scan import DD_CASES, DD_GEOMETRY  # noqa: E402


def _dd_inputs(g):
    X = g["X"]
    n = X.shape[0]
    return X, np.stack([X[(q + 2) % n, 11 + q:18 + q] for q in range(n)])


def test_dilated_profile_orchestration_matches_reference_golden(wb, oracle, scan_golden, monkeypatch):
    """CPU: everything of the dilated / padded profile except the device call -- window and kernel indices at the padded
    borders, grouping by the number of points, the sequential window statistics divided by the FULL kernel length, the
    float divisor -- with the oracle standing in for wb_cuda_subsequence_argmin."""
    from wildboar_b200 import _shim
    ids = {v: k for k, v in oracle.METRIC_IDS.items()}

    def fake(metric_id, params, s, x, k, scaled=False, weight_len=0):
        metric = ids[metric_id]
        kw = {name: getattr(params, name) for name in oracle.DEFAULTS[metric]}
        return oracle.argmin_subsequence(metric, list(s), x, k=k, scaled=scaled, weight_len=weight_len, **kw)
    monkeypatch.setattr(_shim, "subsequence_argmin", fake)
    X, Yd = _dd_inputs(scan_golden)
    for ci, (metric, mp) in enumerate(DD_CASES):
        for scale in (False, True):
            for di, (dil, pad_) in enumerate(DD_GEOMETRY):
                got = wb.distance_profile(Yd, X, metric=metric, metric_params=mp, scale=scale, dilation=dil, padding=pad_)
                assert _same(got, scan_golden[f"dd|{ci}|{int(scale)}|{di}"]), (metric, scale, dil, pad_)
    with pytest.raises(ValueError, match="odd subsequence length"):
        wb.distance_profile(Yd[:, :6], X, metric="dtw", padding="same")
    with pytest.raises(ValueError, match="larger than input"):
        wb.distance_profile(Yd, X[:, :10], metric="dtw", dilation=3)


@pytest.mark.gpu
@pytest.mark.parametrize("ci", range(len(DD_CASES)))
def test_dilated_profile_matches_reference_golden(W, scan_golden, ci):
    metric, mp = DD_CASES[ci]
    X, Yd = _dd_inputs(scan_golden)
    for scale in (False, True):
        for di, (dil, pad_) in enumerate(DD_GEOMETRY):
            got = W.distance_profile(Yd, X, metric=("scaled_" if scale else "") + metric, metric_params=mp, dilation=dil, padding=pad_)
            assert _same(got, scan_golden[f"dd|{ci}|{int(scale)}|{di}"]), (metric, scale, dil, pad_)
