"""ctypes shim over the C ABI of libwbcuda.so (include/wb_cuda.h).

float64 numpy buffers go in and out as plain pointers + sizes; nothing here computes.  The
library is loaded lazily and loading FAILS LOUDLY if the shared object is missing -- there is
no CPU fallback (the oracle under oracle/ is test infrastructure and is never imported here).
"""
import ctypes as C
import os
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libwbcuda.so")

METRIC_IDS = {
    "dtw": 0, "wdtw": 1, "ddtw": 2, "adtw": 3, "lcss": 4, "erp": 5,
    "edr": 6, "msm": 7, "twe": 8, "wddtw": 9, "wlcss": 10,
}


class WbParams(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("r", "g", "p", "c", "epsilon", "penalty", "stiffness")] + [
        ("engine", C.c_int32), ("precision", C.c_int32)]


class WbStats(C.Structure):
    _fields_ = [("kernel_ms", C.c_double), ("total_ms", C.c_double), ("cells", C.c_int64), ("pairs", C.c_int64),
                ("launches", C.c_int32), ("engine", C.c_int32), ("lb_kim_pruned", C.c_int64), ("lb_keogh_pruned", C.c_int64),
                ("strip_w", C.c_int32), ("strip_nr", C.c_int32), ("strip_warps", C.c_int32), ("strip_gring", C.c_int32),
                ("ambiguous", C.c_int64)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


_lib = None
_lock = threading.Lock()
_devices = None
_tls = threading.local()

_DP = C.POINTER(C.c_double)
_IP = C.POINTER(C.c_int64)


def library_path():
    return _LIB_PATH


def lib():
    """Load libwbcuda.so (raises if it has not been built: no fallback)."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(_LIB_PATH):
                    raise RuntimeError(
                        f"{_LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(nvcc, sm_100a).  wildboar_b200 has no CPU fallback."
                    )
                L = C.CDLL(_LIB_PATH)
                i64, ci = C.c_int64, C.c_int
                PP, SP, DV = C.POINTER(WbParams), C.POINTER(WbStats), C.POINTER(C.c_int)
                L.wb_cuda_device_count.restype = ci
                L.wb_cuda_last_error.restype = C.c_char_p
                L.wb_cuda_pairwise.argtypes = [ci, PP, _DP, i64, i64, i64, _DP, i64, i64, i64, _DP, DV, ci, SP]
                L.wb_cuda_pairwise_self.argtypes = [ci, PP, _DP, i64, i64, i64, _DP, DV, ci, SP]
                L.wb_cuda_paired.argtypes = [ci, PP, _DP, i64, i64, i64, _DP, i64, i64, _DP, DV, ci, SP]
                L.wb_cuda_pairwise_nd.argtypes = [ci, PP, _DP, i64, i64, i64, i64, i64, _DP, i64, i64, i64, i64, ci, _DP, DV,
                                                  ci, SP]
                L.wb_cuda_paired_nd.argtypes = [ci, PP, _DP, i64, i64, i64, i64, i64, _DP, i64, i64, i64, ci, _DP, DV, ci,
                                                SP]
                L.wb_cuda_fit.argtypes = [_DP, i64, i64, i64, i64, i64, DV, ci, C.POINTER(C.c_void_p)]
                L.wb_cuda_fit_free.argtypes = [C.c_void_p]
                L.wb_cuda_fit_free.restype = None
                L.wb_cuda_pairwise_fitted.argtypes = [ci, PP, _DP, i64, i64, i64, i64, i64, C.c_void_p, ci, _DP, SP]
                L.wb_cuda_argmin_fitted.argtypes = [ci, PP, _DP, i64, i64, i64, C.c_void_p, i64, _DP, ci, _IP, _DP, SP]
                I32P = C.POINTER(C.c_int32)
                L.wb_cuda_dtw_paths.argtypes = [_DP, i64, i64, i64, _DP, i64, i64, i64, _IP, _IP, i64, C.c_double, _DP, I32P,
                                                I32P, _DP, _DP, ci, SP]
                L.wb_cuda_dba_epoch.argtypes = [C.c_void_p, ci, PP, _DP, i64, i64, _IP, _IP, _DP, _DP, ci, _DP, _DP, SP]
                L.wb_cuda_subsequence_profile.argtypes = [ci, PP, _DP, i64, i64, _DP, i64, i64, i64, ci, _DP, C.c_double, _DP, DV, ci, SP]
                L.wb_cuda_subsequence_argmin.argtypes = [ci, PP, _DP, i64, i64, _DP, i64, i64, i64, ci, i64, i64, _IP, _DP, DV, ci, SP]
                L.wb_cuda_subsequence.argtypes = [ci, PP, _DP, _IP, i64, _DP, i64, i64, i64, ci, ci, _DP, _DP, _IP, DV, ci, SP]
                L.wb_cuda_argmin.argtypes = [ci, PP, _DP, i64, i64, i64, _DP, i64, i64, i64, i64, _DP, ci, _IP, _DP,
                                             DV, ci, SP]
                L.wb_cuda_pairwise_dev.argtypes = [ci, PP, C.c_void_p, i64, i64, C.c_void_p, i64, i64, C.c_void_p,
                                                   C.c_void_p, SP]
                L.wb_cuda_lb_keogh.argtypes = [_DP, i64, i64, _DP, i64, i64, i64, C.c_double, ci, _DP, ci, SP]
                L.wb_cuda_lb_kim.argtypes = [_DP, i64, i64, _DP, i64, i64, i64, _DP, ci, SP]
                L.wb_cuda_fp64_peak.argtypes = [ci, _DP, _DP]
                L.wb_cuda_dtw_envelope.argtypes = [_DP, i64, i64, i64, i64, _DP, _DP, ci, SP]
                L.wb_cuda_dtw_lb_keogh_terms.argtypes = [_DP, _DP, _DP, i64, i64, _DP, _DP, ci, SP]
                L.wb_cuda_host_alloc.argtypes = [C.c_size_t]
                L.wb_cuda_host_alloc.restype = C.c_void_p
                L.wb_cuda_host_free.argtypes = [C.c_void_p]
                L.wb_cuda_host_free.restype = None
                _lib = L
    return _lib


def device_count():
    return int(lib().wb_cuda_device_count())


def set_devices(devices):
    """Select the CUDA devices the host entry points shard rows over (None = default policy)."""
    global _devices
    _devices = None if devices is None else [int(d) for d in devices]


def _resolve_devices(work_cells):
    if _devices is not None:
        return _devices
    env = os.environ.get("WILDBOAR_CUDA_DEVICES")
    if env:
        if env.strip().lower() == "all":
            return list(range(max(device_count(), 1)))
        return [int(t) for t in env.split(",") if t.strip() != ""]
    # default policy: one device for small jobs, every visible device for big ones
    if work_cells >= 5e10:
        n = device_count()
        if n > 1:
            return list(range(n))
    return [0]


_precision = None
# wb_params.precision: bit-exact fp64 / optional fp32 mode / fp64 with fused multiply-add in the DTW-family cell
_PRECISIONS = {"fp64": 0, "fp32": 1, "fp64_fma": 2}


def set_precision(precision):
    """Arithmetic of the DP kernels: "fp64" (default: bit-equal to the reference) or "fp32"
    (optional mode of the north star: <= 1e-4 relative error, about 3x the throughput; lcss, wlcss and
    edr are step functions of a threshold test and always run in fp64) or "fp64_fma" (fp64 with the DTW-family
    cost folded into the running minimum by one fused multiply-add: <= 1e-12 relative, ~15 % faster, not bit-equal
    because the reference build has no FMA).  None = take
    WILDBOAR_CUDA_PRECISION from the environment (default fp64)."""
    global _precision
    if precision is not None and precision not in _PRECISIONS:
        raise ValueError("precision must be 'fp64', 'fp32', 'fp64_fma' or None")
    _precision = precision


def get_precision():
    p = _precision if _precision is not None else os.environ.get("WILDBOAR_CUDA_PRECISION", "fp64").strip().lower()
    if p not in _PRECISIONS:
        raise ValueError("WILDBOAR_CUDA_PRECISION must be fp64, fp32 or fp64_fma")
    return p


def last_stats():
    """wb_stats of the last call on this thread (dict) or None."""
    return getattr(_tls, "stats", None)


def apply_engine_override(params):
    """WILDBOAR_CUDA_ENGINE=rowscan|strip|band|coop forces one DP engine (testing / cross-checks); also
    stamps the selected precision into the parameter block."""
    e = os.environ.get("WILDBOAR_CUDA_ENGINE", "").strip().lower()
    params.engine = {"rowscan": 1, "strip": 2, "band": 3, "coop": 4}.get(e, 0)
    params.precision = _PRECISIONS[get_precision()]
    return params


def _check(rc):
    if rc != 0:
        msg = lib().wb_cuda_last_error()
        raise RuntimeError("wildboar_b200 CUDA call failed: " + (msg.decode() if msg else "unknown error"))


def _rows(a):
    """(pointer, n, T, stride-in-elements) of a 2-D float64 array whose last axis is contiguous."""
    assert a.dtype == np.float64 and a.ndim == 2 and (a.shape[1] <= 1 or a.strides[1] == 8)
    stride = a.strides[0] // 8 if a.shape[0] > 1 else a.shape[1]
    if a.strides[0] % 8 != 0 or stride < a.shape[1]:
        a = np.ascontiguousarray(a)
        stride = a.shape[1]
    return a, a.ctypes.data_as(_DP), a.shape[0], a.shape[1], stride


class _PinnedBlock:
    """Page-locked result memory from the library's pool (wb_cuda_host_alloc), exposed through the array interface;
    the block goes back to the pool when the last array viewing it is gone."""

    def __init__(self, ptr, shape):
        self._ptr = ptr
        self.__array_interface__ = {"shape": tuple(shape), "typestr": "<f8", "data": (ptr, False), "version": 3}

    def __del__(self):
        p, self._ptr = self._ptr, None
        if p and _lib is not None:
            _lib.wb_cuda_host_free(p)


_PINNED_MIN_BYTES = 1 << 20


def result_array(shape):
    """float64 result array: page-locked for results of 1 MB and more (the device writes them back by asynchronous DMA
    at PCIe speed; see wb_cuda_host_alloc), ordinary numpy memory otherwise or when no page-locked memory is to be had.
    WILDBOAR_CUDA_PINNED_RESULTS=0 turns the page-locked results off."""
    n = int(np.prod(shape)) * 8
    if n >= _PINNED_MIN_BYTES and os.environ.get("WILDBOAR_CUDA_PINNED_RESULTS", "1") != "0":
        ptr = lib().wb_cuda_host_alloc(n)
        if ptr:
            return np.asarray(_PinnedBlock(ptr, shape))
    return np.empty(shape, dtype=np.float64)


def pinned_copy(a):
    """A C-contiguous float64 copy of `a` in page-locked memory (the library's pool, like the result arrays): inputs that
    live there reach the device by asynchronous DMA at PCIe speed instead of through the driver's pageable staging.  Falls
    back to an ordinary copy when no page-locked memory is to be had (or for arrays under 1 MB)."""
    a = np.asarray(a, dtype=np.float64)
    out = result_array(a.shape)
    out[...] = a
    return out


def _dev_array(devs):
    return (C.c_int * len(devs))(*devs), len(devs)


def _est_cells(n_pairs, Tx, Ty, r):
    R = max(int(min(Tx, Ty) * r), 1)
    return float(n_pairs) * min(Tx, Ty) * min(2 * R + abs(Tx - Ty), max(Tx, Ty))


def pairwise(metric_id, params, x, y):
    apply_engine_override(params)
    x, xp, nx, Tx, xs = _rows(x)
    st = WbStats()
    if y is None:
        out = result_array((nx, nx))
        dv, nd = _dev_array(_resolve_devices(_est_cells(nx * nx / 2, Tx, Tx, params.r)))
        _check(lib().wb_cuda_pairwise_self(metric_id, C.byref(params), xp, nx, Tx, xs, out.ctypes.data_as(_DP), dv, nd,
                                           C.byref(st)))
    else:
        y, yp, ny, Ty, ys = _rows(y)
        out = result_array((nx, ny))
        dv, nd = _dev_array(_resolve_devices(_est_cells(nx * ny, Tx, Ty, params.r)))
        _check(lib().wb_cuda_pairwise(metric_id, C.byref(params), xp, nx, Tx, xs, yp, ny, Ty, ys,
                                      out.ctypes.data_as(_DP), dv, nd, C.byref(st)))
    _tls.stats = st.as_dict()
    return out


def paired(metric_id, params, x, y):
    apply_engine_override(params)
    x, xp, n, Tx, xs = _rows(x)
    y, yp, ny, Ty, ys = _rows(y)
    assert n == ny
    out = np.empty(n, dtype=np.float64)
    st = WbStats()
    dv, nd = _dev_array(_resolve_devices(_est_cells(n, Tx, Ty, params.r)))
    _check(lib().wb_cuda_paired(metric_id, C.byref(params), xp, n, Tx, xs, yp, Ty, ys, out.ctypes.data_as(_DP), dv, nd,
                                C.byref(st)))
    _tls.stats = st.as_dict()
    return out


def _samples(a):
    """(array, pointer, n, n_dims, T, sample stride, dim stride) of a 3-D float64 TSArray (strides in elements)."""
    assert a.dtype == np.float64 and a.ndim == 3
    n, nd, T = a.shape
    ok = (T <= 1 or a.strides[2] == 8) and a.strides[0] % 8 == 0 and a.strides[1] % 8 == 0
    ss, ds = a.strides[0] // 8, a.strides[1] // 8
    if not ok or (nd > 1 and ds < T) or (n > 1 and ss < T):
        a = np.ascontiguousarray(a)
        ss, ds = nd * T, T
    if n == 1:
        ss = max(ss, nd * T)
    return a, a.ctypes.data_as(_DP), n, nd, T, ss, ds


def pairwise_nd(metric_id, params, x, y, combine):
    """Multivariate pairwise / self join; combine "mean" -> (nx, ny), "full" -> (n_dims, nx, ny)."""
    apply_engine_override(params)
    x, xp, nx, nd, Tx, xss, xds = _samples(x)
    full = combine == "full"
    st = WbStats()
    if y is None:
        out = result_array((nd, nx, nx) if full else (nx, nx))
        dv, ndv = _dev_array(_resolve_devices(nd * _est_cells(nx * nx / 2, Tx, Tx, params.r)))
        _check(lib().wb_cuda_pairwise_nd(metric_id, C.byref(params), xp, nx, nd, Tx, xss, xds, None, 0, 0, 0, 0,
                                         1 if full else 0, out.ctypes.data_as(_DP), dv, ndv, C.byref(st)))
    else:
        y, yp, ny, ndy, Ty, yss, yds = _samples(y)
        assert nd == ndy
        out = result_array((nd, nx, ny) if full else (nx, ny))
        dv, ndv = _dev_array(_resolve_devices(nd * _est_cells(nx * ny, Tx, Ty, params.r)))
        _check(lib().wb_cuda_pairwise_nd(metric_id, C.byref(params), xp, nx, nd, Tx, xss, xds, yp, ny, Ty, yss, yds,
                                         1 if full else 0, out.ctypes.data_as(_DP), dv, ndv, C.byref(st)))
    _tls.stats = st.as_dict()
    return out


def paired_nd(metric_id, params, x, y, combine):
    apply_engine_override(params)
    x, xp, n, nd, Tx, xss, xds = _samples(x)
    y, yp, ny, ndy, Ty, yss, yds = _samples(y)
    assert n == ny and nd == ndy
    full = combine == "full"
    out = np.empty((nd, n) if full else (n,), dtype=np.float64)
    st = WbStats()
    dv, ndv = _dev_array(_resolve_devices(nd * _est_cells(n, Tx, Ty, params.r)))
    _check(lib().wb_cuda_paired_nd(metric_id, C.byref(params), xp, n, nd, Tx, xss, xds, yp, Ty, yss, yds,
                                   1 if full else 0, out.ctypes.data_as(_DP), dv, ndv, C.byref(st)))
    _tls.stats = st.as_dict()
    return out


def _set_mode_result_ok(dist, ordered):
    """Result of a neighbour-set call (include/wb_cuda.h, use_device_lb bit 1): usable when no query was reported as
    ambiguous and -- if the caller is going to SORT the neighbours by distance -- no row holds two equal distances (their
    order after sorting would depend on the heap order, which the mode does not preserve)."""
    if _tls.stats.get("ambiguous", 0) != 0:
        return False
    if ordered and dist.shape[1] > 1:
        sd = np.sort(dist, axis=1)
        if bool((sd[:, 1:] == sd[:, :-1]).any()):
            return False
    return True


def argmin(metric_id, params, x, y, k, lower_bound=None, use_device_lb=False, neighbour_set=False, ordered=False):
    """neighbour_set: the caller consumes the k nearest as a set, or (ordered=True) sorted by distance -- never in heap
    order; the exact scan answers instead whenever the set-mode result would not be unique."""
    if neighbour_set and use_device_lb and 1 < k <= 8:
        idx, dist = _argmin(metric_id, params, x, y, k, lower_bound, 3)
        if _set_mode_result_ok(dist, ordered):
            return idx, dist
    return _argmin(metric_id, params, x, y, k, lower_bound, 1 if use_device_lb else 0)


def _argmin(metric_id, params, x, y, k, lower_bound, lb_flags):
    apply_engine_override(params)
    x, xp, nx, Tx, xs = _rows(x)
    y, yp, ny, Ty, ys = _rows(y)
    idx = np.zeros((nx, k), dtype=np.int64)
    dist = np.zeros((nx, k), dtype=np.float64)
    lbp = None
    if lower_bound is not None:
        lower_bound = np.ascontiguousarray(lower_bound, dtype=np.float64)
        lbp = lower_bound.ctypes.data_as(_DP)
    st = WbStats()
    dv, nd = _dev_array(_resolve_devices(_est_cells(nx * ny, Tx, Ty, params.r)))
    _check(lib().wb_cuda_argmin(metric_id, C.byref(params), xp, nx, Tx, xs, yp, ny, Ty, ys, k, lbp,
                                lb_flags, idx.ctypes.data_as(_IP), dist.ctypes.data_as(_DP), dv, nd,
                                C.byref(st)))
    _tls.stats = st.as_dict()
    return idx.astype(np.intp, copy=False), dist


class FittedSet:
    """A training set kept resident on the device(s) (wb_cuda_fit): estimators upload `_fit_X` once and
    every later query moves only the queries and the result.  x: (n, n_dims, T) float64."""

    def __init__(self, x, devices=None):
        x, xp, n, nd, T, ss, ds = _samples(x)
        self.shape = (n, nd, T)
        devs = list(devices) if devices is not None else _resolve_devices(0.0 if _devices is not None else 1e30)
        dv, ndv = _dev_array(devs)
        h = C.c_void_p()
        _check(lib().wb_cuda_fit(xp, n, nd, T, ss, ds, dv, ndv, C.byref(h)))
        self._h = h
        self.devices = devs

    def close(self):
        h, self._h = getattr(self, "_h", None), None
        if h is not None and _lib is not None:
            _lib.wb_cuda_fit_free(h)

    __del__ = close

    def _handle(self):
        if self._h is None:
            raise RuntimeError("the fitted set has been released")
        return self._h


def pairwise_fitted(metric_id, params, x, fitted, combine="mean"):
    """x (nx, n_dims, Tx) against a FittedSet; (nx, n) or, combine="full", (n_dims, nx, n)."""
    apply_engine_override(params)
    x, xp, nx, nd, Tx, xss, xds = _samples(x)
    full = combine == "full"
    n = fitted.shape[0]
    out = result_array((nd, nx, n) if full else (nx, n))
    st = WbStats()
    _check(lib().wb_cuda_pairwise_fitted(metric_id, C.byref(params), xp, nx, nd, Tx, xss, xds, fitted._handle(),
                                         1 if full else 0, out.ctypes.data_as(_DP), C.byref(st)))
    _tls.stats = st.as_dict()
    return out


def argmin_fitted(metric_id, params, x, fitted, k, lower_bound=None, use_device_lb=False, neighbour_set=False, ordered=False):
    """neighbour_set: the caller only needs the k nearest as a SET, or (ordered=True) sorted by distance (include/wb_cuda.h,
    use_device_lb bit 1): the call is repeated with the exact scan when the set-mode result would not be unique."""
    if neighbour_set and use_device_lb and 1 < k <= 8:
        idx, dist = _argmin_fitted(metric_id, params, x, fitted, k, lower_bound, 3)
        if _set_mode_result_ok(dist, ordered):
            return idx, dist
    return _argmin_fitted(metric_id, params, x, fitted, k, lower_bound, 1 if use_device_lb else 0)


def _argmin_fitted(metric_id, params, x, fitted, k, lower_bound, lb_flags):
    apply_engine_override(params)
    x, xp, nx, Tx, xs = _rows(x)
    idx = np.zeros((nx, k), dtype=np.int64)
    dist = np.zeros((nx, k), dtype=np.float64)
    lbp = None
    if lower_bound is not None:
        lower_bound = np.ascontiguousarray(lower_bound, dtype=np.float64)
        lbp = lower_bound.ctypes.data_as(_DP)
    st = WbStats()
    _check(lib().wb_cuda_argmin_fitted(metric_id, C.byref(params), xp, nx, Tx, xs, fitted._handle(), k, lbp,
                                       lb_flags, idx.ctypes.data_as(_IP), dist.ctypes.data_as(_DP),
                                       C.byref(st)))
    _tls.stats = st.as_dict()
    return idx.astype(np.intp, copy=False), dist


def dtw_paths(a, b, r, weights=None, ia=None, ib=None, want_cost=False, want_matrix=False):
    """Warping paths of pairs (a[ia[p]], b[ib[p]]): (lo, hi[, cost][, matrix]), include/wb_cuda.h."""
    a, ap, na, Ta, as_ = _rows(a)
    b, bp, nb, Tb, bs_ = _rows(b)
    iap = ibp = None
    n_pairs = min(na, nb)
    if ia is not None:
        ia = np.ascontiguousarray(ia, dtype=np.int64)
        iap, n_pairs = ia.ctypes.data_as(_IP), ia.shape[0]
    if ib is not None:
        ib = np.ascontiguousarray(ib, dtype=np.int64)
        ibp, n_pairs = ib.ctypes.data_as(_IP), ib.shape[0]
    if ia is not None and ib is not None:
        assert ia.shape == ib.shape
    wp = None
    if weights is not None:
        weights = np.ascontiguousarray(weights, dtype=np.float64)
        assert weights.shape[0] == max(Ta, Tb)
        wp = weights.ctypes.data_as(_DP)
    I32P = C.POINTER(C.c_int32)
    lo = np.empty((n_pairs, Ta), dtype=np.int32)
    hi = np.empty((n_pairs, Ta), dtype=np.int32)
    cost = np.empty(n_pairs, dtype=np.float64) if want_cost else None
    mat = np.empty((n_pairs, Ta, Tb), dtype=np.float64) if want_matrix else None
    st = WbStats()
    _check(lib().wb_cuda_dtw_paths(ap, na, Ta, as_, bp, nb, Tb, bs_, iap, ibp, n_pairs, float(r), wp,
                                   lo.ctypes.data_as(I32P), hi.ctypes.data_as(I32P),
                                   cost.ctypes.data_as(_DP) if want_cost else None,
                                   mat.ctypes.data_as(_DP) if want_matrix else None, _first_device(), C.byref(st)))
    _tls.stats = st.as_dict()
    out = [lo, hi]
    if want_cost:
        out.append(cost)
    if want_matrix:
        out.append(mat)
    return tuple(out)


def dba_epoch(fitted, metric_id, params, means, offsets, members, sample_weight=None, weights=None, update=True):
    """One DBA step for all clusters (wb_cuda_dba_epoch): (new_means (K, Tm), dist (n_members,))."""
    apply_engine_override(params)
    means = np.ascontiguousarray(means, dtype=np.float64)
    K, Tm = means.shape
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    members = np.ascontiguousarray(members, dtype=np.int64)
    assert offsets.shape[0] == K + 1 and offsets[-1] == members.shape[0]
    swp = wp = None
    if sample_weight is not None:
        sample_weight = np.ascontiguousarray(sample_weight, dtype=np.float64)
        assert sample_weight.shape[0] == fitted.shape[0]
        swp = sample_weight.ctypes.data_as(_DP)
    if weights is not None:
        weights = np.ascontiguousarray(weights, dtype=np.float64)
        assert weights.shape[0] == max(Tm, fitted.shape[2])
        wp = weights.ctypes.data_as(_DP)
    out = np.empty((K, Tm), dtype=np.float64)
    dist = np.empty(members.shape[0], dtype=np.float64)
    st = WbStats()
    _check(lib().wb_cuda_dba_epoch(fitted._handle(), metric_id, C.byref(params), means.ctypes.data_as(_DP), K, Tm,
                                   offsets.ctypes.data_as(_IP), members.ctypes.data_as(_IP), swp, wp, 1 if update else 0,
                                   out.ctypes.data_as(_DP), dist.ctypes.data_as(_DP), C.byref(st)))
    _tls.stats = st.as_dict()
    return out, dist


def subsequence(metric_id, params, subsequences, x, paired=False, scaled=False, s_epsilon=None):
    """Minimum sliding-window distance and first best window of every (sample, subsequence) pair
    (wb_cuda_subsequence): (dist, idx) of shape (nx, n_s) or, paired, (nx,).  s_epsilon: edr's per-subsequence default
    epsilon (std / 4) or None."""
    apply_engine_override(params)
    params.precision = 0
    x, xp, nx, T, xs = _rows(x)
    lens = np.array([len(s_) for s_ in subsequences], dtype=np.int64)
    offsets = np.zeros(len(subsequences) + 1, dtype=np.int64)
    np.cumsum(lens, out=offsets[1:])
    flat = np.ascontiguousarray(np.concatenate([np.asarray(s_, dtype=np.float64).ravel() for s_ in subsequences]))
    shape = (nx,) if paired else (nx, len(subsequences))
    dist = np.empty(shape, dtype=np.float64)
    idx = np.empty(shape, dtype=np.int64)
    st = WbStats()
    cells = float(nx) * T * sum(_est_cells(1, int(m), int(m), params.r) for m in lens) / (nx if paired else 1)
    dv, nd = _dev_array(_resolve_devices(cells))
    eps = None
    if s_epsilon is not None:
        eps = np.ascontiguousarray(s_epsilon, dtype=np.float64)
        if eps.shape != (len(subsequences),):
            raise ValueError("s_epsilon needs one value per subsequence")
    _check(lib().wb_cuda_subsequence(metric_id, C.byref(params), flat.ctypes.data_as(_DP), offsets.ctypes.data_as(_IP),
                                     len(subsequences), xp, nx, T, xs, 1 if paired else 0, 1 if scaled else 0,
                                     eps.ctypes.data_as(_DP) if eps is not None else None, dist.ctypes.data_as(_DP),
                                     idx.ctypes.data_as(_IP), dv, nd, C.byref(st)))
    _tls.stats = st.as_dict()
    return dist, idx.astype(np.intp, copy=False)


def subsequence_profile(metric_id, params, s, x, scaled=False, s_epsilon=None, threshold=float("inf")):
    """Dense matches / distance profile (wb_cuda_subsequence_profile): s is (m,) (against every sample) or (nx, m)
    (paired); returns (nx, T - m + 1) with NaN where the reference reports no match under `threshold`."""
    apply_engine_override(params)
    params.precision = 0
    x, xp, nx, T, xs = _rows(x)
    s = np.ascontiguousarray(np.atleast_2d(np.asarray(s, dtype=np.float64)))
    n_s, m = s.shape
    out = np.empty((nx, T - m + 1), dtype=np.float64)
    st = WbStats()
    eps = None
    if s_epsilon is not None:
        eps = np.ascontiguousarray(np.atleast_1d(s_epsilon), dtype=np.float64)
        if eps.shape != (n_s,):
            raise ValueError("s_epsilon needs one value per subsequence")
    cells = float(nx) * (T - m + 1) * _est_cells(1, int(m), int(m), params.r)
    dv, nd = _dev_array(_resolve_devices(cells))
    _check(lib().wb_cuda_subsequence_profile(metric_id, C.byref(params), s.ctypes.data_as(_DP), n_s, m, xp, nx, T, xs,
                                             1 if scaled else 0, eps.ctypes.data_as(_DP) if eps is not None else None,
                                             float(threshold), out.ctypes.data_as(_DP), dv, nd, C.byref(st)))
    _tls.stats = st.as_dict()
    return out


def subsequence_argmin(metric_id, params, s, x, k, scaled=False, weight_len=0):
    """k closest windows of sample i to subsequence i (wb_cuda_subsequence_argmin): (idx, dist), (nx, k), heap order."""
    apply_engine_override(params)
    params.precision = 0
    x, xp, nx, T, xs = _rows(x)
    s = np.ascontiguousarray(np.atleast_2d(np.asarray(s, dtype=np.float64)))
    n_s, m = s.shape
    idx = np.empty((nx, k), dtype=np.int64)
    dist = np.empty((nx, k), dtype=np.float64)
    st = WbStats()
    cells = float(nx) * (T - m + 1) * _est_cells(1, int(m), int(m), params.r)
    dv, nd = _dev_array(_resolve_devices(cells))
    _check(lib().wb_cuda_subsequence_argmin(metric_id, C.byref(params), s.ctypes.data_as(_DP), n_s, m, xp, nx, T, xs,
                                            1 if scaled else 0, int(k), int(weight_len), idx.ctypes.data_as(_IP), dist.ctypes.data_as(_DP),
                                            dv, nd, C.byref(st)))
    _tls.stats = st.as_dict()
    return idx.astype(np.intp, copy=False), dist


def _first_device():
    devs = _resolve_devices(0.0)
    return devs[0] if devs else 0


def lb_keogh(q, x, r, kind):
    """(nq, nx) LB_Keogh matrix; kind 0 both / 1 left / 2 right (include/wb_cuda.h)."""
    q, qp, nq, T, qs = _rows(q)
    x, xp, nx, Tx, xs = _rows(x)
    assert T == Tx
    out = result_array((nq, nx))
    st = WbStats()
    _check(lib().wb_cuda_lb_keogh(qp, nq, qs, xp, nx, xs, T, float(r), int(kind), out.ctypes.data_as(_DP), _first_device(),
                                  C.byref(st)))
    _tls.stats = st.as_dict()
    return out


def lb_kim(q, x):
    q, qp, nq, T, qs = _rows(q)
    x, xp, nx, Tx, xs = _rows(x)
    assert T == Tx
    out = result_array((nq, nx))
    st = WbStats()
    _check(lib().wb_cuda_lb_kim(qp, nq, qs, xp, nx, xs, T, out.ctypes.data_as(_DP), _first_device(), C.byref(st)))
    _tls.stats = st.as_dict()
    return out


def dtw_envelope(x, w):
    """(lower, upper) envelopes of half-width w for every row of x (n, T)."""
    x, xp, n, T, xs = _rows(x)
    lo = np.empty((n, T), dtype=np.float64)
    hi = np.empty((n, T), dtype=np.float64)
    st = WbStats()
    _check(lib().wb_cuda_dtw_envelope(xp, n, T, xs, int(w), lo.ctypes.data_as(_DP), hi.ctypes.data_as(_DP), _first_device(),
                                      C.byref(st)))
    _tls.stats = st.as_dict()
    return lo, hi


def dtw_lb_keogh_terms(x, lower, upper):
    """(min_dist (n,), cb (n, T)): LB_Keogh of every row of x against the envelope rows, with the per-step terms."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    lower = np.ascontiguousarray(lower, dtype=np.float64)
    upper = np.ascontiguousarray(upper, dtype=np.float64)
    n, T = x.shape
    assert lower.shape == (n, T) and upper.shape == (n, T)
    md = np.empty(n, dtype=np.float64)
    cb = np.empty((n, T), dtype=np.float64)
    st = WbStats()
    _check(lib().wb_cuda_dtw_lb_keogh_terms(x.ctypes.data_as(_DP), lower.ctypes.data_as(_DP), upper.ctypes.data_as(_DP), n, T,
                                            md.ctypes.data_as(_DP), cb.ctypes.data_as(_DP), _first_device(), C.byref(st)))
    _tls.stats = st.as_dict()
    return md, cb


def pairwise_dev(metric_id, params, x_ptr, nx, Tx, y_ptr, ny, Ty, out_ptr, stream=0, want_stats=True):
    """Device-resident pairwise (raw device pointers, e.g. torch.Tensor.data_ptr())."""
    st = WbStats()
    _check(lib().wb_cuda_pairwise_dev(metric_id, C.byref(params), x_ptr, nx, Tx, y_ptr, ny, Ty, out_ptr, stream,
                                      C.byref(st) if want_stats else None))
    return st.as_dict() if want_stats else None


def fp64_peak(mix=0):
    a, b = C.c_double(0), C.c_double(0)
    _check(lib().wb_cuda_fp64_peak(mix, C.byref(a), C.byref(b)))
    return a.value, b.value
