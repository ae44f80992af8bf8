"""TEST INFRASTRUCTURE -- ctypes front-end of oracle/_build/liboracle.so (elastic_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) import
this module.  It mirrors the reference's three entry points on 2-D float64 arrays
(dim already selected):

  pairwise(metric, x, y=None, **params)  -> (nx, ny)   reference: _distance.py:1178
  paired(metric, x, y, **params)         -> (n,)       reference: _distance.py:1082 (swapped operands)
  argmin(metric, x, y, k, lower_bound, **params) -> (idx, dist) in heap order, _distance.py:1320
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")

METRIC_IDS = {
    "dtw": 0, "wdtw": 1, "ddtw": 2, "adtw": 3, "lcss": 4, "erp": 5,
    "edr": 6, "msm": 7, "twe": 8, "wddtw": 9, "wlcss": 10,
}

# defaults of the reference constructors
DEFAULTS = {
    "dtw": dict(r=1.0), "wdtw": dict(r=1.0, g=0.05), "ddtw": dict(r=1.0), "adtw": dict(r=1.0, p=1.0),
    "lcss": dict(r=1.0, epsilon=1.0), "erp": dict(r=1.0, g=0.0), "edr": dict(r=1.0, epsilon=float("nan")),
    "msm": dict(r=1.0, c=1.0), "twe": dict(r=1.0, penalty=1.0, stiffness=0.001),
    "wddtw": dict(r=1.0, g=0.05), "wlcss": dict(r=1.0, epsilon=1.0, g=0.05),
}


class Params(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("r", "g", "p", "c", "epsilon", "penalty", "stiffness")]


def build(force=False):
    if force or not os.path.exists(_LIB_PATH) or (
        os.path.getmtime(_LIB_PATH) < os.path.getmtime(os.path.join(_HERE, "elastic_oracle.c"))
    ):
        subprocess.check_call(["make", "-C", _HERE, "_build/liboracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int64)
        i64 = C.c_int64
        _lib.orc_pairwise.argtypes = [C.c_int, C.POINTER(Params), dp, i64, i64, i64, dp, i64, i64, i64, dp, C.c_int]
        _lib.orc_pairwise_self.argtypes = [C.c_int, C.POINTER(Params), dp, i64, i64, i64, dp, C.c_int]
        _lib.orc_paired.argtypes = [C.c_int, C.POINTER(Params), dp, i64, i64, i64, dp, i64, i64, dp, C.c_int]
        _lib.orc_argmin.argtypes = [C.c_int, C.POINTER(Params), dp, i64, i64, i64, dp, i64, i64, i64, i64, dp, ip, dp, C.c_int]
        _lib.orc_heap_replay.argtypes = [i64, i64, ip, dp, ip, dp, ip]
        _lib.orc_envelope.argtypes = [dp, i64, i64, dp, dp]
        _lib.orc_lb_keogh_one.argtypes = [dp, dp, dp, i64]
        _lib.orc_lb_keogh_one.restype = C.c_double
        _lib.orc_lb_warp_size.argtypes = [i64, C.c_double]
        _lib.orc_lb_warp_size.restype = i64
        _lib.orc_lb_keogh_matrix.argtypes = [dp, i64, dp, i64, i64, C.c_double, C.c_int, dp]
        _lib.orc_lb_kim_matrix.argtypes = [dp, i64, dp, i64, i64, dp]
        _lib.orc_compute_r.argtypes = [i64, C.c_double]
        _lib.orc_compute_r.restype = i64
        _lib.orc_std.argtypes = [dp, i64]
        _lib.orc_std.restype = C.c_double
    return _lib


def make_params(metric, **kw):
    d = dict(r=1.0, g=0.0, p=1.0, c=1.0, epsilon=1.0, penalty=1.0, stiffness=0.001)
    d.update(DEFAULTS[metric])
    for k, v in kw.items():
        if k not in DEFAULTS[metric]:
            raise TypeError(f"unexpected metric param {k!r} for {metric}")
        d[k] = float(v)
    return Params(**d)


def _arr(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if a.ndim == 1:
        a = a.reshape(1, -1)
    assert a.ndim == 2
    return a


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def pairwise(metric, x, y=None, n_jobs=1, **params):
    p = make_params(metric, **params)
    x = _arr(x)
    if y is None:
        out = np.empty((x.shape[0], x.shape[0]))
        rc = lib().orc_pairwise_self(METRIC_IDS[metric], C.byref(p), _dp(x), x.shape[0], x.shape[1], x.shape[1], _dp(out), n_jobs)
    else:
        y = _arr(y)
        out = np.empty((x.shape[0], y.shape[0]))
        rc = lib().orc_pairwise(METRIC_IDS[metric], C.byref(p), _dp(x), x.shape[0], x.shape[1], x.shape[1],
                                _dp(y), y.shape[0], y.shape[1], y.shape[1], _dp(out), n_jobs)
    assert rc == 0
    return out


def paired(metric, x, y, n_jobs=1, **params):
    p = make_params(metric, **params)
    x, y = _arr(x), _arr(y)
    assert x.shape[0] == y.shape[0]
    out = np.empty(x.shape[0])
    rc = lib().orc_paired(METRIC_IDS[metric], C.byref(p), _dp(x), x.shape[0], x.shape[1], x.shape[1],
                          _dp(y), y.shape[1], y.shape[1], _dp(out), n_jobs)
    assert rc == 0
    return out


def argmin(metric, x, y, k=1, lower_bound=None, n_jobs=1, **params):
    p = make_params(metric, **params)
    x, y = _arr(x), _arr(y)
    k = min(k, y.shape[0])
    idx = np.zeros((x.shape[0], k), dtype=np.int64)
    dist = np.zeros((x.shape[0], k))
    lb = None
    if lower_bound is not None:
        lb = np.ascontiguousarray(lower_bound, dtype=np.float64)
        assert lb.shape == (x.shape[0], y.shape[0])
    rc = lib().orc_argmin(METRIC_IDS[metric], C.byref(p), _dp(x), x.shape[0], x.shape[1], x.shape[1],
                          _dp(y), y.shape[0], y.shape[1], y.shape[1], k, _dp(lb) if lb is not None else None,
                          idx.ctypes.data_as(C.POINTER(C.c_int64)), _dp(dist), n_jobs)
    assert rc == 0
    return idx, dist


def heap_replay(k, index, value):
    index = np.ascontiguousarray(index, dtype=np.int64)
    value = np.ascontiguousarray(value, dtype=np.float64)
    oi = np.zeros(k, dtype=np.int64)
    ov = np.zeros(k)
    n = C.c_int64(0)
    ip = C.POINTER(C.c_int64)
    lib().orc_heap_replay(k, len(index), index.ctypes.data_as(ip), _dp(value), oi.ctypes.data_as(ip), _dp(ov), C.byref(n))
    return oi, ov, n.value


def envelope(t, w):
    t = np.ascontiguousarray(t, dtype=np.float64)
    lo, hi = np.empty_like(t), np.empty_like(t)
    lib().orc_envelope(_dp(t), len(t), int(w), _dp(lo), _dp(hi))
    return lo, hi


def lb_keogh_one(q, lower, upper):
    q = np.ascontiguousarray(q, dtype=np.float64)
    return lib().orc_lb_keogh_one(_dp(q), _dp(np.ascontiguousarray(lower)), _dp(np.ascontiguousarray(upper)), len(q))


LB_KINDS = {"both": 0, "left": 1, "right": 2}


def lb_warp_size(T, r):
    return lib().orc_lb_warp_size(int(T), float(r))


def lb_keogh(q, x, r=1.0, kind="both"):
    """DtwKeoghLowerBound(r, kind).fit(x).transform(q) -> (nq, nx), lb.py:359-432."""
    q, x = _arr(q), _arr(x)
    assert q.shape[1] == x.shape[1]
    out = np.empty((q.shape[0], x.shape[0]))
    rc = lib().orc_lb_keogh_matrix(_dp(q), q.shape[0], _dp(x), x.shape[0], x.shape[1], float(r), LB_KINDS[kind], _dp(out))
    assert rc == 0
    return out


def lb_kim(q, x):
    """DtwKimLowerBound().fit(x).transform(q) -> (nq, nx), lb.py:224-311."""
    q, x = _arr(q), _arr(x)
    assert q.shape[1] == x.shape[1]
    out = np.empty((q.shape[0], x.shape[0]))
    rc = lib().orc_lb_kim_matrix(_dp(q), q.shape[0], _dp(x), x.shape[0], x.shape[1], _dp(out))
    assert rc == 0
    return out


def compute_r(n, r):
    return lib().orc_compute_r(int(n), float(r))


def cells_per_pair(Tx, Ty, r, metric="dtw"):
    """DP cells the reference evaluates for one pair: sum_i (j_stop(i) - j_start(i))."""
    R = compute_r(min(Tx, Ty), r)
    if metric in ("ddtw", "wddtw"):
        Tx, Ty = Tx - 2, Ty - 2
        if min(Tx, Ty) < 1:
            return 0
    max_len = max(0, Ty - Tx) + R
    min_len = max(0, Tx - Ty)
    i = np.arange(Tx)
    js = np.maximum(0, i - min_len - R + 1)
    je = np.minimum(Ty, i + max_len)
    return int(np.sum(np.maximum(je - js, 0)))


# ---------------------------------------------------------------------------------------------
# DTW alignment / warping path / DBA (reference: distance/dtw.py:246-690)
# ---------------------------------------------------------------------------------------------
def warp_size(x_size, r, y_size=0):
    """dtw.py:38-40 `_compute_warp_size` (max, not min, of the two lengths)."""
    import math
    return max(math.floor(max(x_size, y_size) * r), 1)


def jeong_weight(n, g=0.05):
    """dtw.py:347-374 (numpy's exp, as in the reference)."""
    return 1.0 / (1.0 + np.exp(-g * (np.arange(n, dtype=float) - n / 2.0)))


def dtw_alignment(x, y, r=1.0, weight=None):
    """`_dtw_alignment` (EL:1011-1073): cells the reference leaves uninitialised (np.empty) are NaN here."""
    L = lib()
    L.orc_dtw_alignment.argtypes = [C.POINTER(C.c_double), C.c_int64, C.POINTER(C.c_double), C.c_int64, C.c_int64,
                                    C.POINTER(C.c_double), C.POINTER(C.c_double)]
    x = np.ascontiguousarray(x, dtype=np.float64).ravel()
    y = np.ascontiguousarray(y, dtype=np.float64).ravel()
    out = np.full((x.shape[0], y.shape[0]), np.nan)
    wp = None
    if weight is not None:
        weight = np.ascontiguousarray(weight, dtype=np.float64)
        wp = _dp(weight)
    L.orc_dtw_alignment(_dp(x), x.shape[0], _dp(y), y.shape[0], warp_size(x.shape[0], r, y.shape[0]), wp, _dp(out))
    return out


def dtw_path(alignment):
    """First / last column of the optimal path per row (`dtw_mapping`, dtw.py:385-413)."""
    L = lib()
    ip32 = C.POINTER(C.c_int32)
    L.orc_dtw_path.argtypes = [C.POINTER(C.c_double), C.c_int64, C.c_int64, ip32, ip32]
    a = np.ascontiguousarray(alignment, dtype=np.float64)
    lo = np.zeros(a.shape[0], dtype=np.int32)
    hi = np.zeros(a.shape[0], dtype=np.int32)
    L.orc_dtw_path(_dp(a), a.shape[0], a.shape[1], lo.ctypes.data_as(ip32), hi.ctypes.data_as(ip32))
    return lo, hi


def dtw_average_mm(X, r=1.0, g=None, init=None, sample_weight=None, tol=1e-5, max_epoch=50):
    """`_mm_dtw_average` (dtw.py:655-690) with the cost function of `dtw_average` (dtw.py:590-605)."""
    X = _arr(X)
    mean = np.array(init, dtype=float, copy=True)
    metric, mp = ("dtw", dict(r=r)) if g is None else ("wdtw", dict(r=r, g=g))

    def costfn(mean):
        cost = pairwise(metric, mean.reshape(1, -1), X, **mp)[0]
        return np.mean(cost) if sample_weight is None else np.average(cost, weights=sample_weight)

    cost = costfn(mean)
    for _ in range(max_epoch):
        z = np.zeros(mean.shape[0])
        V = np.zeros(mean.shape[0])
        for i in range(X.shape[0]):
            weight = None if g is None else jeong_weight(max(mean.shape[0], X.shape[1]), g)
            lo, hi = dtw_path(dtw_alignment(mean, X[i], r=r, weight=weight))
            w = 1.0 if sample_weight is None else sample_weight[i]
            for m in range(mean.shape[0]):
                for xx in range(lo[m], hi[m] + 1):
                    V[m] += w
                    z[m] += X[i, xx] * w
        mean = z / V
        prev_cost, cost = cost, costfn(mean)
        if abs(prev_cost - cost) < tol:
            break
    return mean, cost


# ---------------------------------------------------------------------------------------------
# subsequence search, DTW family (reference: _distance.py:543-729, _elastic.pyx:622-815, 2206-2615)
# ---------------------------------------------------------------------------------------------
def pairwise_subsequence(metric, subsequences, x, **params):
    """(dist, idx) of shape (n_samples, n_subsequences): minimum over the sliding windows, first best window."""
    L = lib()
    L.orc_subsequence_distance.argtypes = [C.c_int, C.POINTER(Params), C.POINTER(C.c_double), C.c_int64,
                                           C.POINTER(C.c_double), C.c_int64, C.POINTER(C.c_int64)]
    L.orc_subsequence_distance.restype = C.c_double
    x = _arr(x)
    p = make_params(metric, **params)
    dist = np.empty((x.shape[0], len(subsequences)))
    idx = np.zeros((x.shape[0], len(subsequences)), dtype=np.int64)
    eps_auto = metric == "edr" and np.isnan(p.epsilon)
    for k, s in enumerate(subsequences):
        s = np.ascontiguousarray(s, dtype=np.float64)
        if eps_auto:  # EdrSubsequenceMetric._distance EL:2762-2765: s_std / 4 with the subsequence's own std
            p.epsilon = _subsequence_mean_std(s)[1] / 4.0
        for i in range(x.shape[0]):
            j = C.c_int64(0)
            dist[i, k] = L.orc_subsequence_distance(METRIC_IDS[metric], C.byref(p), _dp(s), s.shape[0], _dp(x[i]), x.shape[1],
                                                    C.byref(j))
            idx[i, k] = j.value
    return dist, idx


def _subsequence_mean_std(s):
    """ScaledSubsequenceMetric.from_array (_cdistance.pyx:453-467) + `std if std != 0 else 1.0` (:283-298)."""
    mean, std = np.mean(s), np.std(s)
    if std <= 1e-13:
        std = 0.0
    return float(mean), float(std if std != 0 else 1.0)


def pairwise_scaled_subsequence(metric, subsequences, x, **params):
    """`scaled_<metric>` = ScaledSubsequenceMetricWrap(Metric) (_cdistance.pyx:470-551): (dist, idx), (n_samples, n_subsequences)."""
    L = lib()
    L.orc_scaled_subsequence_distance.argtypes = [C.c_int, C.POINTER(Params), C.POINTER(C.c_double), C.c_int64, C.c_double,
                                                  C.c_double, C.POINTER(C.c_double), C.c_int64, C.POINTER(C.c_int64)]
    L.orc_scaled_subsequence_distance.restype = C.c_double
    x = _arr(x)
    p = make_params(metric, **params)
    dist = np.empty((x.shape[0], len(subsequences)))
    idx = np.zeros((x.shape[0], len(subsequences)), dtype=np.int64)
    for k, s in enumerate(subsequences):
        s = np.ascontiguousarray(s, dtype=np.float64)
        mean, std = _subsequence_mean_std(s)
        for i in range(x.shape[0]):
            j = C.c_int64(0)
            dist[i, k] = L.orc_scaled_subsequence_distance(METRIC_IDS[metric], C.byref(p), _dp(s), s.shape[0], mean, std,
                                                           _dp(x[i]), x.shape[1], C.byref(j))
            idx[i, k] = j.value
    return dist, idx


def inc_window_stats(x, m):
    """(mean, std) of every window of length m of the 1-D series x as ScaledSubsequenceMetricWrap computes them."""
    L = lib()
    dp = C.POINTER(C.c_double)
    L.orc_inc_window_stats.argtypes = [dp, C.c_int64, C.c_int64, dp, dp]
    L.orc_inc_window_stats.restype = None
    x = np.ascontiguousarray(x, dtype=np.float64)
    nw = x.shape[0] - m + 1
    mean, std = np.empty(nw), np.empty(nw)
    L.orc_inc_window_stats(_dp(x), x.shape[0], m, _dp(mean), _dp(std))
    return mean, std


def pairwise_scaled_dtw_subsequence(subsequences, x, r=1.0):
    """scaled_dtw subsequence search (ScaledDtwSubsequenceMetric, EL:1928-2060): (dist, idx), (n_samples, n_subsequences)."""
    L = lib()
    L.orc_scaled_dtw_subsequence.argtypes = [C.POINTER(C.c_double), C.c_int64, C.c_double, C.c_double, C.POINTER(C.c_double),
                                             C.c_int64, C.c_double, C.POINTER(C.c_int64)]
    L.orc_scaled_dtw_subsequence.restype = C.c_double
    x = _arr(x)
    dist = np.empty((x.shape[0], len(subsequences)))
    idx = np.zeros((x.shape[0], len(subsequences)), dtype=np.int64)
    for k, s in enumerate(subsequences):
        s = np.ascontiguousarray(s, dtype=np.float64)
        mean, std = np.mean(s), np.std(s)      # _cdistance.pyx:453-467 (EPSILON = 1e-13), :370
        if std <= 1e-13:
            std = 0.0
        std = std if std != 0 else 1.0
        for i in range(x.shape[0]):
            j = C.c_int64(0)
            dist[i, k] = L.orc_scaled_dtw_subsequence(_dp(s), s.shape[0], float(mean), float(std), _dp(x[i]), x.shape[1], float(r),
                                                      C.byref(j))
            idx[i, k] = j.value
    return dist, idx


def subsequence_matches(metric, s, x, threshold=float("inf"), scaled=False, mean_std=None, **params):
    """Dense form of SubsequenceMetric._matches for ONE subsequence against every sample of x (or, with s 2-D, the i:th
    subsequence against the i:th sample): (n_samples, n_windows), the window's distance where the reference reports a
    match under `threshold`, NaN elsewhere.  mean_std: (mean, std) the reference hands to the metric (default: the
    from_array values, _cdistance.pyx:453-467); edr's default epsilon (unscaled) is std / 4 of it."""
    L = lib()
    dp = C.POINTER(C.c_double)
    L.orc_subsequence_matches.argtypes = [C.c_int, C.POINTER(Params), dp, C.c_int64, C.c_double, C.c_double, dp, C.c_int64,
                                          C.c_int, C.c_double, dp]
    L.orc_subsequence_matches.restype = C.c_int64
    x = _arr(x)
    s = np.ascontiguousarray(s, dtype=np.float64)
    subs = [s] * x.shape[0] if s.ndim == 1 else list(s)
    assert len(subs) == x.shape[0]
    p = make_params(metric, **params)
    eps_auto = metric == "edr" and np.isnan(p.epsilon) and not scaled
    m = subs[0].shape[0]
    out = np.empty((x.shape[0], x.shape[1] - m + 1))
    for i, si in enumerate(subs):
        si = np.ascontiguousarray(si)
        mean, std = _subsequence_mean_std(si) if mean_std is None else mean_std(si)
        if eps_auto:
            p.epsilon = std / 4.0
        L.orc_subsequence_matches(METRIC_IDS[metric], C.byref(p), _dp(si), m, mean, std, _dp(x[i]), x.shape[1],
                                  1 if scaled else 0, float(threshold), _dp(out[i]))
    return out


def seq_mean_std(s):
    """fast_mean_std (utils/_stats.pyx:22-42): sequential sums, std 0 when the variance is <= 1e-13."""
    s = np.ascontiguousarray(s, dtype=np.float64)
    return float(np.cumsum(s)[-1] / s.shape[0]), float(lib().orc_std(_dp(s), s.shape[0]))


def argmin_subsequence(metric, subs, x, k=1, scaled=False, weight_len=0, **params):
    """argmin_subsequence_distance (_distance.py:1636-1790, _cdistance.pyx:1380-1600): the k closest windows of the i:th
    sample to the i:th subsequence under the sequential scan with Metric._eadistance; (idx, dist) of shape (n, k) in heap
    order.  scaled: subsequence z-normalised with fast_mean_std (std 0 -> 1), windows with the running IncStats."""
    L = lib()
    dp = C.POINTER(C.c_double)
    L.orc_argmin_subsequence_w.argtypes = [C.c_int, C.POINTER(Params), dp, C.c_int64, C.c_double, C.c_double, dp, C.c_int64,
                                           C.c_int, C.c_int64, C.c_int64, C.POINTER(C.c_int64), dp]
    L.orc_argmin_subsequence_w.restype = C.c_int64
    x = _arr(x)
    assert len(subs) == x.shape[0]
    p = make_params(metric, **params)
    idx = np.zeros((x.shape[0], k), dtype=np.int64)
    dist = np.zeros((x.shape[0], k))
    for i, s in enumerate(subs):
        s = np.ascontiguousarray(s, dtype=np.float64)
        mean, std = seq_mean_std(s) if scaled else (0.0, 1.0)
        if std == 0.0:
            std = 1.0
        n = L.orc_argmin_subsequence_w(METRIC_IDS[metric], C.byref(p), _dp(s), s.shape[0], mean, std, _dp(x[i]), x.shape[1],
                                       1 if scaled else 0, k, int(weight_len), idx[i].ctypes.data_as(C.POINTER(C.c_int64)), _dp(dist[i]))
        assert n == k or metric in ("ddtw", "wddtw", "lcss", "erp", "edr", "msm", "twe", "adtw"), (n, k)
    return idx, dist
