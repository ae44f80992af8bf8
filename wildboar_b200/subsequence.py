"""Subsequence search with the DTW-family metrics on the CUDA path.

Mirrors ``wildboar.distance.pairwise_subsequence_distance`` and ``paired_subsequence_distance``
(reference: src/wildboar/distance/_distance.py:543-636, 639-729) for ``metric`` in {dtw, wdtw, adtw, ddtw, wddtw}
with ``scale=False`` and for ``scaled_dtw`` (= ``metric="dtw", scale=True``: the z-normalised UCR-suite search):
same arguments, return shapes (``_format_return``), index of the FIRST best window.

B200-first: instead of one early-abandoning scan per (sample, subsequence) pair, all sliding windows of all
samples are the second operand of one pairwise DP launch per subsequence (windows addressed with stride 1, so a
warp's 32 consecutive windows read consecutive addresses), followed by a first-minimum reduction per sample.
``scaled_dtw`` additionally replays the reference's scan exactly (its LB_Kim prefilter is not a valid bound and decides
which window wins).  The other scaled and the non-DTW subsequence metrics are not covered; there is no CPU fallback.
"""
import numbers

import numpy as np

from . import _shim
from .distance import _check_ts_array, _format_return, _make_metric, check_array

__all__ = ["pairwise_subsequence_distance", "paired_subsequence_distance"]

_SUBSEQUENCE_METRICS = ("dtw", "wdtw", "adtw", "ddtw", "wddtw")


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
    if metric == "scaled_dtw":
        return "dtw", True
    if isinstance(metric, str) and metric.startswith("scaled_"):
        raise ValueError(f"the scaled subsequence metric {metric!r} is not accelerated (only scaled_dtw); use wildboar.distance for it")
    if metric not in _SUBSEQUENCE_METRICS:
        raise ValueError(
            "unsupported metric '{}', 'metric' must be a str among {}".format(metric, set(_SUBSEQUENCE_METRICS) | {"scaled_dtw"})
        )
    return metric, False


def _z_normalise(s):
    """ScaledSubsequenceMetric.from_array (_cdistance.pyx:453-467) + the std passed on (:283-298)."""
    mean, std = np.mean(s), np.std(s)
    if std <= _EPSILON:
        std = 0.0
    return (s - mean) / (std if std != 0 else 1.0)


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
    if scaled:
        if any(s.shape[0] < 3 for s in y):
            raise ValueError("scaled_dtw needs subsequences of at least 3 samples (the reference reads S[1], S[2] unconditionally)")
        y = [_z_normalise(s) for s in y]
    return y, x, x_[:, int(dim), :], m, scaled


def pairwise_subsequence_distance(y, x, *, dim=0, metric="dtw", metric_params=None, scale=False, return_index=False,
                                  n_jobs=None):
    """Minimum distance between every subsequence of ``y`` and every sample of ``x`` (_distance.py:543-636).

    Returns an array of shape (n_samples, n_subsequences) (squeezed like the reference) and, with
    ``return_index``, the start of the first best-matching window.
    """
    y, x, xd, m, scaled = _prepare(y, x, dim, metric, metric_params, scale)
    min_dist, min_ind = _shim.subsequence(m.metric_id, m._params(), y, xd, paired=False, scaled=scaled)
    if return_index:
        return _format_return(min_dist, len(y), x.ndim), _format_return(min_ind, len(y), x.ndim)
    return _format_return(min_dist, len(y), x.ndim)


def paired_subsequence_distance(y, x, *, dim=0, metric="dtw", metric_params=None, scale=False, return_index=False,
                                n_jobs=None):
    """Minimum distance between the i:th subsequence and the i:th sample (_distance.py:639-729)."""
    y, x, xd, m, scaled = _prepare(y, x, dim, metric, metric_params, scale)
    n_samples = x.shape[0] if x.ndim > 1 else 1
    if len(y) != n_samples:
        raise ValueError(
            "The number of subsequences and samples must be the same, got %d subsequences and %d samples." % (len(y), n_samples)
        )
    min_dist, min_ind = _shim.subsequence(m.metric_id, m._params(), y, xd, paired=True, scaled=scaled)
    if return_index:
        return _format_return(min_dist, len(y), x.ndim), _format_return(min_ind, len(y), x.ndim)
    return _format_return(min_dist, len(y), x.ndim)
