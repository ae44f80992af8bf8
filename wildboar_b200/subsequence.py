"""Subsequence search with the elastic metrics on the CUDA path.

Mirrors ``wildboar.distance.pairwise_subsequence_distance`` and ``paired_subsequence_distance``
(reference: src/wildboar/distance/_distance.py:543-636, 639-729) for every ELASTIC entry of the reference's
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
import numbers

import numpy as np

from . import _shim
from .distance import _check_ts_array, _format_return, _make_metric, check_array

__all__ = ["pairwise_subsequence_distance", "paired_subsequence_distance"]

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
