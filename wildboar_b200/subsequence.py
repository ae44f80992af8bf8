"""Subsequence search with the elastic metrics on the CUDA path.

Mirrors ``wildboar.distance.pairwise_subsequence_distance`` and ``paired_subsequence_distance``
(reference: src/wildboar/distance/_distance.py:543-636, 639-729), ``subsequence_match`` / ``paired_subsequence_match``
(:732-1080), ``distance_profile`` (:1477-1600, dilation 1 / no padding) and ``argmin_subsequence_distance`` (:1636-1790)
for every ELASTIC entry of the reference's
``_SUBSEQUENCE_METRICS`` (_distance.py:143-178): ``dtw, wdtw, adtw, ddtw, wddtw, lcss, erp, edr, msm, twe`` and their
``scaled_`` forms (``scale=True``; ``scaled_dtw`` is the UCR-suite search, the others ScaledSubsequenceMetricWrap):
same arguments, return shapes (``_format_return``), index of the first best window under the reference's scan.

B200-first: instead of one early-abandoning scan per (sample, subsequence) pair, all sliding windows of all
samples are the second operand of one pairwise DP launch per subsequence (unscaled windows addressed in place with
stride 1, so a warp's 32 consecutive windows read consecutive addresses; scaled windows z-normalised on the device),
followed by a first-minimum reduction per sample -- or, where the reference's early abandoning decides the result
(lcss / erp / edr / msm / twe and every scaled metric), by an exact replay of its scan (one warp per sample).
The non-elastic subsequence metrics (euclidean, manhattan, mass, ...) are not part of this path; there is no CPU
fallback.
"""
import math
import numbers
import warnings

import numpy as np

from . import _shim
from .distance import _check_ts_array, _format_return, _make_metric, check_array

__all__ = ["pairwise_subsequence_distance", "paired_subsequence_distance", "subsequence_match",
           "paired_subsequence_match", "distance_profile", "argmin_subsequence_distance"]

_SUBSEQUENCE_METRICS = ("dtw", "wdtw", "adtw", "ddtw", "wddtw", "lcss", "erp", "edr", "msm", "twe")


def _is_arraylike(e):
    return hasattr(e, "__len__") or hasattr(e, "shape") or hasattr(e, "__array__")


def _validate_subsequence(y):
    """_distance.py:336-373."""
    if len(y) == 0:
        raise ValueError("Subsequence y cannot be empty.")
    if isinstance(y, np.ndarray) and y.dtype != object:
        if y.ndim == 1:
            return [y.astype(float)]
        elif y.ndim == 2:
            y = list(y.astype(float))
        else:
            raise ValueError("Expected 2D array, got {}D array instead:\narray={}.\n".format(y.ndim, y))
    elif any(_is_arraylike(e) for e in y):
        y = [np.array(e, dtype=np.double) for e in y]
    else:
        y = [np.array(y, dtype=np.double)]
    return y


_EPSILON = 1e-13  # _cdistance.pxd EPSILON


def _check_subsequence_metric(metric, scale):
    """Returns (metric name of the DP, scaled?)  (_distance.py:208-228 `_infer_scaled_metric`, :263-300)."""
    if callable(metric):
        raise ValueError("callable subsequence metrics are not accelerated; use wildboar.distance for them")
    if scale and isinstance(metric, str) and not metric.startswith("scaled_"):
        metric = "scaled_" + metric
    scaled = isinstance(metric, str) and metric.startswith("scaled_")
    base = metric[len("scaled_"):] if scaled else metric
    if base not in _SUBSEQUENCE_METRICS:
        raise ValueError(
            "unsupported metric '{}', 'metric' must be a str among {}".format(
                metric, set(_SUBSEQUENCE_METRICS) | {"scaled_" + b for b in _SUBSEQUENCE_METRICS})
        )
    return base, scaled


def _mean_std(s):
    """ScaledSubsequenceMetric.from_array (_cdistance.pyx:453-467) + `std if std != 0 else 1.0` (:283-298)."""
    mean, std = np.mean(s), np.std(s)
    if std <= _EPSILON:
        std = 0.0
    return mean, (std if std != 0 else 1.0)


def _z_normalise(s):
    """The subsequence as the scaled metrics see it: (s - mean) / std (_elastic.pyx:2145-2166, _cdistance.pyx:505-509)."""
    mean, std = _mean_std(s)
    return (s - mean) / std


def _prepare(y, x, dim, metric, metric_params, scale):
    y = _validate_subsequence(y)
    x = check_array(x, allow_3d=True, ensure_2d=False, dtype=np.double)
    for s in y:
        if s.shape[0] > x.shape[-1]:
            raise ValueError("Invalid subsequnce shape (%d > %d)" % (s.shape[0], x.shape[-1]))
        if s.ndim != 1 or s.shape[0] < 1 or not np.all(np.isfinite(s)):
            raise ValueError("every subsequence must be a non-empty, finite 1-D array")
    metric, scaled = _check_subsequence_metric(metric, scale)
    m = _make_metric(metric, metric_params)
    x_ = _check_ts_array(x)
    if isinstance(dim, bool) or not isinstance(dim, numbers.Integral) or not 0 <= dim < x_.shape[1]:
        raise ValueError("The parameter dim must be 0 <= dim < n_dims")
    s_eps = None
    if metric == "edr" and not scaled and np.isnan(m._params().epsilon):
        # EdrSubsequenceMetric._distance (_elastic.pyx:2762-2765): epsilon = s_std / 4 of each subsequence
        s_eps = np.array([_mean_std(s)[1] / 4.0 for s in y], dtype=np.double)
    if scaled:
        if metric == "dtw" and any(s.shape[0] < 3 for s in y):
            raise ValueError("scaled_dtw needs subsequences of at least 3 samples (the reference reads S[1], S[2] unconditionally)")
        y = [_z_normalise(s) for s in y]
    return y, x, x_[:, int(dim), :], m, scaled, s_eps


def pairwise_subsequence_distance(y, x, *, dim=0, metric="dtw", metric_params=None, scale=False, return_index=False,
                                  n_jobs=None):
    """Minimum distance between every subsequence of ``y`` and every sample of ``x`` (_distance.py:543-636).

    Returns an array of shape (n_samples, n_subsequences) (squeezed like the reference) and, with
    ``return_index``, the start of the first best-matching window.
    """
    y, x, xd, m, scaled, s_eps = _prepare(y, x, dim, metric, metric_params, scale)
    min_dist, min_ind = _shim.subsequence(m.metric_id, m._params(), y, xd, paired=False, scaled=scaled, s_epsilon=s_eps)
    if return_index:
        return _format_return(min_dist, len(y), x.ndim), _format_return(min_ind, len(y), x.ndim)
    return _format_return(min_dist, len(y), x.ndim)


def paired_subsequence_distance(y, x, *, dim=0, metric="dtw", metric_params=None, scale=False, return_index=False,
                                n_jobs=None):
    """Minimum distance between the i:th subsequence and the i:th sample (_distance.py:639-729)."""
    y, x, xd, m, scaled, s_eps = _prepare(y, x, dim, metric, metric_params, scale)
    n_samples = x.shape[0] if x.ndim > 1 else 1
    if len(y) != n_samples:
        raise ValueError(
            "The number of subsequences and samples must be the same, got %d subsequences and %d samples." % (len(y), n_samples)
        )
    min_dist, min_ind = _shim.subsequence(m.metric_id, m._params(), y, xd, paired=True, scaled=scaled, s_epsilon=s_eps)
    if return_index:
        return _format_return(min_dist, len(y), x.ndim), _format_return(min_ind, len(y), x.ndim)
    return _format_return(min_dist, len(y), x.ndim)


# ---------------------------------------------------------------------------------------------
# subsequence_match / paired_subsequence_match / distance_profile (_distance.py:732-1080, 1477-1600)
# ---------------------------------------------------------------------------------------------
def _std_below_mean(mult):
    """_distance.py:310-333: max(mean - mult * std, min)."""
    def f(d):
        return max(np.mean(d) - mult * np.std(d), np.min(d))
    return f


_THRESHOLD = {"auto": _std_below_mean(2.0)}


def _jagged(lst):
    arr = np.empty(len(lst), dtype=object)
    arr[:] = lst
    return arr


def _rows_to_matches(dense):
    """(n_samples, n_windows) with NaN = no match  ->  per-sample (indices, distances) arrays, None where nothing matched
    (`_new_match_array` / `_new_distance_array`, _cdistance.pyx:936-953; matches in window order)."""
    indices, distances = [], []
    for row in dense:
        idx = np.flatnonzero(~np.isnan(row))
        if idx.size == 0:
            indices.append(None)
            distances.append(None)
        else:
            indices.append(idx.astype(np.intp))
            distances.append(row[idx])
    return indices, distances


def _filter(indices, distances, keep):
    out_i, out_d = [], []
    for sample, (index, distance) in enumerate(zip(indices, distances)):
        if index is None:
            out_i.append(None)
            out_d.append(None)
        else:
            sel = keep(sample, index, distance)
            out_i.append(index[sel])
            out_d.append(distance[sel])
    return out_i, out_d


def _keep_nontrivial(exclude):
    """_exclude_trivial_matches (_distance.py:400-443): best matches first, drop those within `exclude` of a kept one."""
    def keep(_, index, distance):
        order = np.argsort(distance)
        sel = np.zeros(order.size, dtype=bool)
        kept = []
        for o in order:
            if not any(index[o] - exclude < e < index[o] + exclude for e in kept):
                kept.append(index[o])
                sel[o] = True
        return sel
    return keep


def _resolve_threshold(threshold, max_matches, n_samples, allow_array):
    """threshold / max_matches defaults and the post-filter of the match functions (_distance.py:848-893, 1033-1055)."""
    if threshold is None:
        threshold = np.inf
        if max_matches is None:
            max_matches = 10
    max_dist = None
    if callable(threshold) or isinstance(threshold, str):
        if isinstance(threshold, str):
            if threshold not in _THRESHOLD:
                raise ValueError("threshold must be one of %s, got %r" % (set(_THRESHOLD), threshold))
            fn = _THRESHOLD[threshold]
        else:
            fn = threshold

        def max_dist(_, d):
            return d <= fn(d)
        threshold = np.inf
    elif allow_array and _is_arraylike(threshold):
        arr = np.asarray(threshold)
        if len(arr) != n_samples:
            raise ValueError(f"threshold array length ({len(arr)}) must match the number of samples ({n_samples})")

        def max_dist(i, d):
            return d <= arr[i]
        threshold = np.inf
    elif not isinstance(threshold, numbers.Real):
        raise TypeError("threshold must be str, callable%s or float, not %s"
                        % (", array-like" if allow_array else "", type(threshold).__qualname__))
    return float(threshold), max_matches, max_dist


def _profile(subs, xd, metric, m, scaled, threshold, mean_std):
    """Dense matches of subs ((m,) or (n, m)) against the samples xd through the C ABI.  mean_std: row-wise statistics of
    a C-contiguous (n, m) array -> (mean, std) arrays, the ones the reference uses on the calling path."""
    sub2 = np.ascontiguousarray(np.atleast_2d(subs), dtype=np.double)
    s_eps = None
    if metric == "edr" and not scaled and np.isnan(m._params().epsilon):
        s_eps = mean_std(sub2)[1] / 4.0
    if scaled:
        if metric == "dtw" and sub2.shape[1] < 3:
            raise ValueError("scaled_dtw needs subsequences of at least 3 samples (the reference reads S[1], S[2] unconditionally)")
        mean, std = mean_std(sub2)
        if np.any(std == 0):
            raise ValueError("constant subsequence: the reference divides by its zero standard deviation on this path")
        sub2 = (sub2 - mean[:, None]) / std[:, None]
    return _shim.subsequence_profile(m.metric_id, m._params(), sub2, xd, scaled=scaled, s_epsilon=s_eps, threshold=threshold)


def _from_array_mean_std(a):
    """Row-wise ScaledSubsequenceMetric.from_array (_cdistance.pyx:453-467) + `std if std != 0 else 1.0` (:283-298); numpy's
    reductions over the contiguous last axis are the 1-D reductions of every row."""
    mean, std = np.mean(a, axis=1), np.std(a, axis=1)
    std = np.where(std <= _EPSILON, 0.0, std)
    return mean, np.where(std != 0, std, 1.0)


def _view_mean_std(a):
    """Row-wise _ts_view_update_statistics (_cdistance.pyx:167-185): sequential sums (cumsum accumulates in order); the std
    is handed on as is (0 when the variance is <= 1e-13) by _DistanceProfile (_cdistance.pyx:1686-1697)."""
    a = np.atleast_2d(a)
    n = a.shape[1]
    ex = np.cumsum(a, axis=1)[:, -1]
    ex2 = np.cumsum(a * a, axis=1)[:, -1]
    mean = ex / n
    var = ex2 / n - mean * mean
    return mean, np.where(var > _EPSILON, np.sqrt(np.where(var > 0, var, 0.0)), 0.0)


def subsequence_match(y, x, threshold=None, *, dim=0, metric="dtw", metric_params=None, scale=False, max_matches=None,
                      exclude=None, return_distance=False, n_jobs=None):
    """Start indices (and distances) of the windows of every sample that match the subsequence (_distance.py:732-924).

    Same threshold forms (None -> the 10 best, float, "auto", callable, per-sample array), ``exclude`` and ``max_matches``
    post-filters and return shapes as the reference; the per-window distances come from ONE device launch over all
    windows of all samples (wb_cuda_subsequence_profile).  ``dim`` may select any dimension (the reference only accepts 0).
    """
    y = _validate_subsequence(y)
    if len(y) > 1:
        raise ValueError("A single subsequence expected, got %d" % len(y))
    y = y[0]
    x = check_array(x, allow_3d=True, ensure_2d=False, dtype=np.double)
    if y.shape[0] > x.shape[-1]:
        raise ValueError("Invalid subsequnce shape (%d > %d)" % (y.shape[0], x.shape[-1]))
    metric, scaled = _check_subsequence_metric(metric, scale)
    m = _make_metric(metric, metric_params)
    x_ = _check_ts_array(x)
    if isinstance(dim, bool) or not isinstance(dim, numbers.Integral) or not 0 <= dim < x_.shape[1]:
        raise ValueError("The parameter dim must be 0 <= dim < n_dims")
    if n_jobs is not None:
        warnings.warn("n_jobs is not yet supported.", UserWarning)
    n_samples = x.shape[0] if x.ndim > 1 else 1
    threshold, max_matches, max_dist = _resolve_threshold(threshold, max_matches, n_samples, allow_array=True)
    if exclude is not None and (isinstance(exclude, bool) or not isinstance(exclude, (numbers.Integral, numbers.Real))):
        raise TypeError("exclude must be an int or a float, got %s" % type(exclude).__qualname__)
    if exclude is not None and exclude < 0:
        raise ValueError("exclude == %r, must be >= 0." % (exclude,))
    if exclude is not None and not isinstance(exclude, numbers.Integral):
        exclude = math.ceil(y.size * exclude)
    dense = _profile(y, x_[:, int(dim), :], metric, m, scaled, threshold, _from_array_mean_std)
    indices, distances = _rows_to_matches(dense)
    if max_dist is not None:
        indices, distances = _filter(indices, distances, lambda i, _, d: max_dist(i, d))
    if exclude:
        indices, distances = _filter(indices, distances, _keep_nontrivial(exclude))
    if max_matches:
        indices, distances = _filter(indices, distances, lambda _, __, d: np.argsort(d)[:max_matches])
    indices = _format_return(_jagged(indices), len(y), x.ndim)
    if indices.size == 1:
        indices = indices.item()
    if return_distance:
        distances = _format_return(_jagged(distances), len(y), x.ndim)
        if distances.size == 1:
            distances = distances.item()
        return indices, distances
    return indices


def paired_subsequence_match(y, x, threshold=None, *, dim=0, metric="dtw", metric_params=None, scale=False, max_matches=None,
                             return_distance=False, n_jobs=None):
    """Matches of the i:th subsequence in the i:th sample (_distance.py:927-1080); subsequences of equal length share one
    device launch."""
    y = _validate_subsequence(y)
    x = check_array(x, allow_3d=True, ensure_2d=False, dtype=np.double)
    n_samples = x.shape[0] if x.ndim > 1 else 1
    if len(y) != n_samples:
        raise ValueError("The number of subsequences and samples must be the same, got %d subsequences and %d samples."
                         % (len(y), n_samples))
    for s in y:
        if s.shape[0] > x.shape[-1]:
            raise ValueError("invalid subsequnce shape (%d > %d)" % (s.shape[0], x.shape[-1]))
    metric, scaled = _check_subsequence_metric(metric, scale)
    m = _make_metric(metric, metric_params)
    x_ = _check_ts_array(x)
    if isinstance(dim, bool) or not isinstance(dim, numbers.Integral) or not 0 <= dim < x_.shape[1]:
        raise ValueError("The parameter dim must be 0 <= dim < n_dims")
    if n_jobs is not None:
        warnings.warn("n_jobs is not yet supported.", UserWarning)
    threshold, max_matches, max_dist = _resolve_threshold(threshold, max_matches, n_samples, allow_array=False)
    xd = x_[:, int(dim), :]
    indices, distances = [None] * n_samples, [None] * n_samples
    lengths = np.array([s.shape[0] for s in y])
    for length in np.unique(lengths):
        sel = np.flatnonzero(lengths == length)
        dense = _profile(np.array([y[q] for q in sel]), np.ascontiguousarray(xd[sel]), metric, m, scaled, threshold,
                         _from_array_mean_std)
        gi, gd = _rows_to_matches(dense)
        for q, a, b in zip(sel, gi, gd):
            indices[q], distances[q] = a, b
    if max_dist is not None:
        indices, distances = _filter(indices, distances, lambda i, _, d: max_dist(i, d))
    if max_matches:
        indices, distances = _filter(indices, distances, lambda _, __, d: np.argsort(d)[:max_matches])
    indices = _format_return(_jagged(indices), len(y), x.ndim)
    if indices.size == 1:
        indices = indices.reshape(1)
    if return_distance:
        distances = _format_return(_jagged(distances), len(y), x.ndim)
        if distances.size == 1:
            distances = distances.reshape(1)
        return indices, distances
    return indices


def distance_profile(y, x, *, dilation=1, padding=0, dim=0, metric="dtw", metric_params=None, scale=False, n_jobs=None):
    """Distance of the i:th subsequence to every window of the i:th sample (_distance.py:1477-1600).

    ``dilation=1, padding=0`` ends in ``_distance_profile`` (_cdistance.pyx:1655-1725, the subsequence metrics); any other
    setting in ``_dilated_distance_profile`` (_cdistance.pyx:804-935, 1728-1862: dilated, zero-padded windows through
    ``Metric._eadistance``; wdtw / wddtw keep the weight table of the full series, as ``metric.reset(X, X)`` sizes it)."""
    if isinstance(dilation, bool) or not isinstance(dilation, numbers.Integral) or dilation < 1:
        raise ValueError("dilation must be an int >= 1")
    y = np.squeeze(check_array(y, dtype=np.double, ensure_2d=False))
    x = np.squeeze(check_array(x, allow_3d=True, ensure_2d=False, dtype=np.double))
    if x.ndim == 1 and y.ndim != 1:
        x = np.broadcast_to(x, shape=(y.shape[0], x.shape[0]))
    if y.ndim == 1 and x.ndim != 1:
        y = np.broadcast_to(y, shape=(x.shape[0], y.shape[0]))
    x_ = _check_ts_array(x)
    y_ = _check_ts_array(y)
    shapelet_size = (y_.shape[2] - 1) * dilation + 1
    if isinstance(padding, str):
        if padding != "same":
            raise ValueError("padding must be an int >= 0 or 'same'")
        if y_.shape[2] % 2 == 0:
            raise ValueError("padding='same' is only supported for odd subsequence length")
        padding = shapelet_size // 2
    elif isinstance(padding, bool) or not isinstance(padding, numbers.Integral) or padding < 0:
        raise ValueError("padding must be an int >= 0 or 'same'")
    if shapelet_size > x_.shape[2] + 2 * padding:
        raise ValueError("subsequence in y is larger than input in x.")
    if y_.shape[0] != x_.shape[0]:
        raise ValueError("y and x must have the same number of samples.")
    if isinstance(dim, bool) or not isinstance(dim, numbers.Integral) or dim < 0 or dim >= x_.shape[1]:
        raise ValueError(f"The parameter dim must be dim ({dim}) < n_dims ({x_.shape[1]})")
    metric, scaled = _check_subsequence_metric(metric, scale)
    m = _make_metric(metric, metric_params)
    if dilation != 1 or padding != 0:
        # _DilatedDistanceProfile reads dimension `dim` of BOTH operands (&S[i, dim, 0], &X[i, dim, 0], CD:1778-1796); with a
        # univariate y and dim > 0 the reference reads past the end of y, which is refused here.  Its metric check on this
        # branch is check_metric (the full registry): every elastic metric of that registry is accepted except
        # wlcss / scaled_wlcss, which the device scan does not carry (ValueError from _check_subsequence_metric above).
        if dim >= y_.shape[1]:
            raise ValueError(f"dim ({dim}) must be < the number of dimensions of y ({y_.shape[1]}) when dilation != 1 or padding != 0")
        return np.squeeze(_dilated_profile(np.ascontiguousarray(y_[:, int(dim), :]), np.ascontiguousarray(x_[:, int(dim), :]), metric, m,
                                           scaled, int(dilation), int(padding)))
    # _DistanceProfile ignores `dim` and always reads dimension 0 of both operands (&self.y[i, 0, 0], &self.x[i, 0, 0],
    # CD:1690-1699); reproduced, so a multivariate x with dim > 0 gives the reference's answer
    dp = _profile(np.ascontiguousarray(y_[:, 0, :]), x_[:, 0, :], metric, m, scaled, np.inf, _view_mean_std)
    return np.squeeze(dp)


def _dilated_geometry(k_len, x_len, dilation, padding):
    """Per output position of `dilated_distance_profile` (_cdistance.pyx:804-858, stride 1): the sample indices and the
    kernel indices that enter the comparison (both truncated where the dilated kernel hangs over the zero padding)."""
    kernel_size = (k_len - 1) * dilation + 1
    output_size = x_len + 2 * padding - kernel_size + 1
    geo = []
    for o in range(output_size):
        padding_offset = padding - o
        if padding_offset > 0:
            kernel_offset = padding_offset if padding_offset % dilation == 0 else padding_offset + dilation - (padding_offset % dilation)
            input_offset = kernel_offset - padding_offset
        else:
            kernel_offset = 0
            input_offset = abs(padding_offset)
        convolution_size = min(x_len, input_offset + kernel_size - max(0, padding_offset)) - input_offset
        js = np.arange(0, max(convolution_size, 0), dilation)
        geo.append((input_offset + js, (js + kernel_offset) // dilation))
    return geo


def _dilated_profile(y, x, metric, m, scaled, dilation, padding):
    """`_dilated_distance_profile` (_cdistance.pyx:1809-1862) on the device: every (sample, output position) pair is one
    equal-length comparison `Metric._eadistance(x window, kernel part)` -- the window FIRST -- which is what
    wb_cuda_subsequence_argmin evaluates for a "subsequence" = the window against a "sample" = the kernel part with one
    window and k = 1.  Output positions are grouped by the number of points that enter (borders are truncated); the weight
    tables of wdtw / wddtw span the full series (`weight_len`), as the reference's `metric.reset(X, X)` leaves them."""
    n, k_len = y.shape
    x_len = x.shape[1]
    if scaled:
        # _distance.py:1581-1585: the subsequences are z-normalised with numpy before the driver is called
        std = np.std(y, axis=-1, keepdims=True)
        mean = np.mean(y, axis=-1, keepdims=True)
        std[std < _EPSILON] = 1
        y = (y - mean) / std
    geo = _dilated_geometry(k_len, x_len, dilation, padding)
    out = np.empty((n, len(geo)), dtype=np.double)
    by_k = {}
    for o, (xi, ki) in enumerate(geo):
        by_k.setdefault(len(xi), []).append(o)
    for k, positions in by_k.items():
        if k < 1 or (metric in ("ddtw", "wddtw") and k < 3):
            raise ValueError("a border window of this dilated profile keeps %d point(s); the reference writes no value there "
                             "and shifts the rest of the row -- use a smaller padding" % k)
        xi = np.stack([geo[o][0] for o in positions])            # (P, k) sample indices
        ki = np.stack([geo[o][1] for o in positions])            # (P, k) kernel indices
        step = max(1, (1 << 24) // max(len(positions) * k, 1))   # samples per call: bounded host / device buffers
        for a in range(0, n, step):
            win = x[a:a + step][:, xi]                            # (s, P, k)
            ker = y[a:a + step][:, ki]
            if scaled:
                # scaled_dilated_distance_profile (_cdistance.pyx:899-921): sequential sums over the k points that enter,
                # divided by the FULL kernel length
                mean = np.cumsum(win, axis=-1)[..., -1] / k_len
                var = np.cumsum(win * win, axis=-1)[..., -1] / k_len - mean * mean
                std = np.where(var > _EPSILON, np.sqrt(np.where(var > 0, var, 0.0)), 1.0)
                win = (win - mean[..., None]) / std[..., None]
            s_, P = win.shape[0], win.shape[1]
            dist = _shim.subsequence_argmin(m.metric_id, m._params(), np.ascontiguousarray(win.reshape(s_ * P, k)),
                                            np.ascontiguousarray(ker.reshape(s_ * P, k)), 1, scaled=False,
                                            weight_len=x_len)[1][:, 0]  # wdtw / wddtw: metric.reset(X, X) sized the weights
            # `tmp_dist / (<float>k / k_len)`: a C float division, widened
            out[a:a + step, positions] = dist.reshape(s_, P) / np.float64(np.float32(k) / np.float32(k_len))
    return out


def _seq_mean_std(a):
    """Row-wise fast_mean_std (utils/_stats.pyx:22-42) as `_ScaledArgminSubsequenceDistance` uses it (_cdistance.pyx:1504-1506):
    sequential sums, std 0 (variance <= 1e-13) replaced by 1."""
    mean, std = _view_mean_std(a)
    return mean, np.where(std != 0.0, std, 1.0)


def argmin_subsequence_distance(y, x, *, dim=0, k=1, metric="dtw", metric_params=None, scale=False, return_distance=False,
                                n_jobs=None):
    """Start indices (and distances) of the ``k`` windows of the i:th sample closest to the i:th subsequence
    (_distance.py:1636-1790): (n_samples, k) arrays in the reference's heap order.  ``metric`` is any elastic entry of
    ``_METRICS`` (``scaled_`` prefix or ``scale=True`` for the z-normalised scan); subsequences of equal length share one
    device launch."""
    if isinstance(k, bool) or not isinstance(k, numbers.Integral) or k < 1:
        raise ValueError("k must be an int >= 1")
    x = np.asarray(x)
    if isinstance(y, np.ndarray) and y.dtype == float:
        y = check_array(y, allow_3d=False, dtype=np.double)
        subs = [np.ascontiguousarray(s) for s in np.atleast_2d(y)]
    else:
        subs = []
        for s in y:
            s = np.asarray(s, dtype=np.double)
            if s.ndim > 1:
                raise ValueError("shapelet must be 1d-array")
            subs.append(s)
    if x.ndim == 1:
        x = np.broadcast_to(x, shape=(len(subs), x.shape[0]))
    x = check_array(x, allow_3d=True, dtype=np.double)
    x_ = _check_ts_array(x)
    max_len = max(s.shape[0] for s in subs)
    if not max_len <= x_.shape[2]:
        raise ValueError("the longest subsequence must be shorter than samples.")
    if len(subs) != x_.shape[0]:
        raise ValueError("both arrays must have the same number of samples.")
    if isinstance(dim, bool) or not isinstance(dim, numbers.Integral) or dim < 0 or dim >= x_.shape[1]:
        raise ValueError(f"The parameter dim must be dim ({dim}) < n_dims ({x_.shape[1]})")
    if not k <= (x_.shape[2] - max_len + 1):
        raise ValueError("k must be less x.shape[-1] - y.shape[-1] + 1.")
    if callable(metric):
        raise ValueError("callable metrics are not accelerated; use wildboar.distance for them")
    scaled = (isinstance(metric, str) and metric.startswith("scaled_")) or bool(scale)
    if isinstance(metric, str) and metric.startswith("scaled_"):
        metric = metric[7:]
    if metric not in _SUBSEQUENCE_METRICS:
        raise ValueError("unsupported metric '{}', 'metric' must be a str among {}".format(
            metric, set(_SUBSEQUENCE_METRICS) | {"scaled_" + b for b in _SUBSEQUENCE_METRICS}))
    m = _make_metric(metric, metric_params)
    xd = x_[:, int(dim), :]
    indices = np.empty((len(subs), k), dtype=np.intp)
    distances = np.empty((len(subs), k), dtype=np.double)
    lengths = np.array([s.shape[0] for s in subs])
    for length in np.unique(lengths):
        sel = np.flatnonzero(lengths == length)
        group = np.ascontiguousarray(np.array([subs[q] for q in sel]), dtype=np.double)
        if scaled:
            mean, std = _seq_mean_std(group)
            group = (group - mean[:, None]) / std[:, None]
        gi, gd = _shim.subsequence_argmin(m.metric_id, m._params(), group, np.ascontiguousarray(xd[sel]), k, scaled=scaled)
        indices[sel], distances[sel] = gi, gd
    if return_distance:
        return indices, distances
    return indices
