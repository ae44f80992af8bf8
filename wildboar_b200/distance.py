"""Host-side mirror of ``wildboar.distance`` for the elastic metrics.

Same entry points, ``metric=`` strings, ``metric_params``, return shapes and exceptions as the
reference (src/wildboar/distance/_distance.py: ``pairwise_distance`` :1178, ``paired_distance``
:1082, ``argmin_distance`` :1320, ``check_metric`` :231, ``_format_return`` :514); the bodies
hand float64 buffers to the CUDA library through the C ABI (``_shim``) instead of the Cython
batch drivers of ``_cdistance.pyx``.  Parameter validation follows the metric constructors of
``_elastic.pyx`` (:3133, 3326, 3233, 3347, 3441, 3546, 3574, 3681, 3896, 3993).
"""
import math
import numbers
import warnings

import numpy as np

from . import _shim

__all__ = ["pairwise_distance", "paired_distance", "argmin_distance", "check_metric"]

# ---------------------------------------------------------------------------------------------
# validation helpers (utils/validation.py:404-620)
# ---------------------------------------------------------------------------------------------
with np.errstate(invalid="ignore"):
    _EOS_BITS = np.array([0x7F800009], dtype=np.uint32).view(np.float32).astype(np.float64).view(np.uint64)[0]


def _is_end_of_series(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64) == _EOS_BITS


_POOL = None


def _sum_is_finite(array):
    """isfinite(sum(array)); large contiguous arrays are summed in slices on a few threads (numpy releases the GIL)."""
    global _POOL
    with np.errstate(over="ignore", invalid="ignore"):
        if array.size < (1 << 22) or not array.flags.c_contiguous:
            return bool(np.isfinite(array.sum()))
        if _POOL is None:
            import os
            from concurrent.futures import ThreadPoolExecutor
            _POOL = ThreadPoolExecutor(max_workers=max(1, min(8, os.cpu_count() or 1)))
        flat = array.reshape(-1)
        n = _POOL._max_workers
        step = -(-flat.shape[0] // n)
        parts = list(_POOL.map(lambda k: flat[k * step:(k + 1) * step].sum(), range(n)))
        return bool(np.isfinite(np.sum(parts)))


def check_array(array, *, allow_3d=False, ensure_2d=True, ensure_ts_array=False, dtype=float, input_name=""):
    """Numeric, finite, equal-length time series array (subset of utils/validation.py:404)."""
    try:
        array = np.asarray(array, dtype=np.float64 if dtype in (float, np.double, np.float64) else dtype)
    except (TypeError, ValueError) as e:
        raise ValueError(f"could not convert input to a float array: {e}") from e
    if array.ndim == 0:
        raise ValueError(
            "Expected 2D array, got scalar array instead:\narray={}.".format(array)
        )
    if ensure_2d and array.ndim == 1:
        raise ValueError("Expected 2D array, got 1D array instead:\narray={}.".format(array))
    if not allow_3d and array.ndim >= 3:
        raise ValueError("Found array with dim %d. None expected <= 2." % array.ndim)
    if allow_3d and array.ndim >= 4:
        raise ValueError("Found array with dim %d. None expected <= 3." % array.ndim)
    if array.ndim >= 2 and array.shape[0] < 1:
        raise ValueError(
            "Found array with %d sample(s) (shape=%s) while a minimum of 1 is required." % (array.shape[0], array.shape)
        )
    if array.ndim in (2, 3) and array.shape[-1] < 1:
        raise ValueError(
            "Found array with %d feature(s) (shape=%s) while a minimum of 1 is required." % (array.shape[-1], array.shape)
        )
    # one pass over the data in the common case (as sklearn's _assert_all_finite does): a finite sum means that
    # every element is finite; only otherwise look for what is wrong.  (Two boolean passes over a 400 MB
    # reference set cost more than the whole device side of a nearest-neighbour query.)
    if not _sum_is_finite(array):
        padded = input_name + " " if input_name else ""
        if np.isnan(array).any():
            if _is_end_of_series(array).any():
                raise ValueError(f"Input {padded}expected time series of equal length.")
            raise ValueError(f"Input {padded}contains NaN.")
        if np.isinf(array).any():
            raise ValueError(f"Input {padded}contains infinity.")
    return _check_ts_array(array) if ensure_ts_array else array


def _check_ts_array(array):
    """(n_samples, n_dims, n_timestep) float64 with a contiguous last axis (validation.py:594)."""
    if array.ndim == 1:
        array = array.reshape(1, 1, array.shape[0])
    elif array.ndim == 2:
        array = array.reshape(array.shape[0], 1, array.shape[1])
    last_stride = array.strides[2] // array.itemsize
    if last_stride != 1:
        array = np.ascontiguousarray(array)
    return array.astype(float, copy=False)


def _check_scalar(x, name, *, min_val=None, max_val=None, include_min=True, include_max=True):
    """sklearn.utils.check_scalar(x, name, float, ...) semantics (TypeError / ValueError)."""
    if isinstance(x, bool) or not isinstance(x, numbers.Real):
        raise TypeError(f"{name} must be an instance of float, not {type(x).__qualname__}.")
    x = float(x)
    if min_val is not None and (x < min_val if include_min else x <= min_val):
        op = ">=" if include_min else ">"
        raise ValueError(f"{name} == {x}, must be {op} {min_val}.")
    if max_val is not None and (x > max_val if include_max else x >= max_val):
        op = "<=" if include_max else "<"
        raise ValueError(f"{name} == {x}, must be {op} {max_val}.")
    return x


# ---------------------------------------------------------------------------------------------
# Metric objects: parameter holders mirroring the reference constructors
# ---------------------------------------------------------------------------------------------
class Metric:
    """Base of the elastic metric parameter objects (reference: _cdistance.pxd:268-316)."""

    name = None
    is_elastic = True
    _fields = ("r",)

    def _params(self):
        d = dict(r=1.0, g=0.0, p=1.0, c=1.0, epsilon=1.0, penalty=1.0, stiffness=0.001)
        for f in self._fields:
            d[f] = float(getattr(self, f))
        return _shim.WbParams(d["r"], d["g"], d["p"], d["c"], d["epsilon"], d["penalty"], d["stiffness"], 0, 0)

    @property
    def metric_id(self):
        return _shim.METRIC_IDS[self.name]

    def __reduce__(self):
        return self.__class__, tuple(getattr(self, f) for f in self._fields)

    def __repr__(self):
        return "%s(%s)" % (type(self).__name__, ", ".join(f"{f}={getattr(self, f)!r}" for f in self._fields))


class DtwMetric(Metric):  # EL:3126
    name = "dtw"

    def __init__(self, r=1.0):
        self.r = _check_scalar(r, "r", min_val=0.0, max_val=1.0)


class DerivativeDtwMetric(DtwMetric):  # EL:3228
    name = "ddtw"


class WeightedDtwMetric(DtwMetric):  # EL:3322
    name = "wdtw"
    _fields = ("r", "g")

    def __init__(self, r=1.0, g=0.05):
        super().__init__(r=r)
        self.g = _check_scalar(g, "g", min_val=0.0)


class AmercingDtwMetric(DtwMetric):  # EL:3344
    name = "adtw"
    _fields = ("r", "p")

    def __init__(self, r=1.0, p=1.0):
        super().__init__(r=r)
        self.p = _check_scalar(p, "p")


class WeightedDerivativeDtwMetric(DerivativeDtwMetric):  # EL:3404 (g is not validated there)
    name = "wddtw"
    _fields = ("r", "g")

    def __init__(self, r=1.0, g=0.05):
        super().__init__(r=r)
        self.g = float(g)


def _deprecated_threshold(epsilon, threshold):
    # TODO(1.4) in the reference: `threshold` was renamed to `epsilon` in 1.2 (EL:3443-3449)
    if not (isinstance(threshold, float) and math.isnan(threshold)):
        warnings.warn(
            "The parameter threshold has been renamed to epsilon in 1.2 and will be removed in 1.4.",
            FutureWarning,
        )
        return threshold
    return epsilon


class LcssMetric(Metric):  # EL:3433
    name = "lcss"
    _fields = ("r", "epsilon")

    def __init__(self, r=1.0, epsilon=1.0, threshold=float("nan")):
        epsilon = _deprecated_threshold(epsilon, threshold)
        self.r = _check_scalar(r, "r", min_val=0.0, max_val=1.0)
        self.epsilon = _check_scalar(epsilon, "epsilon", min_val=0, include_min=False)


class WeightedLcssMetric(LcssMetric):  # EL:3542
    name = "wlcss"
    _fields = ("r", "epsilon", "g")

    def __init__(self, r=1.0, epsilon=1.0, g=0.05, threshold=float("nan")):
        super().__init__(r=r, epsilon=epsilon, threshold=threshold)
        self.g = _check_scalar(g, "g", min_val=0.0)


class ErpMetric(Metric):  # EL:3564
    name = "erp"
    _fields = ("r", "g")

    def __init__(self, r=1.0, g=0.0):
        self.r = _check_scalar(r, "r", min_val=0.0, max_val=1.0)
        self.g = _check_scalar(g, "g", min_val=0)


class EdrMetric(Metric):  # EL:3671; epsilon=NaN => max(std_x, std_y) / 4 per pair
    name = "edr"
    _fields = ("r", "epsilon")

    def __init__(self, r=1.0, epsilon=float("nan"), threshold=float("nan")):
        self.r = _check_scalar(r, "r", min_val=0.0, max_val=1.0)
        epsilon = _deprecated_threshold(epsilon, threshold)
        if isinstance(epsilon, bool) or not isinstance(epsilon, numbers.Real):
            raise TypeError(f"epsilon must be an instance of float, not {type(epsilon).__qualname__}.")
        if not math.isnan(epsilon):
            epsilon = _check_scalar(epsilon, "epsilon", min_val=0, include_min=False)
        self.epsilon = float(epsilon)


class MsmMetric(Metric):  # EL:3887
    name = "msm"
    _fields = ("r", "c")

    def __init__(self, r=1.0, c=1.0):
        self.r = _check_scalar(r, "r", min_val=0.0, max_val=1.0)
        self.c = _check_scalar(c, "c", min_val=0)


class TweMetric(Metric):  # EL:3985
    name = "twe"
    _fields = ("r", "penalty", "stiffness")

    def __init__(self, r=1.0, penalty=1.0, stiffness=0.001):
        self.r = _check_scalar(r, "r", min_val=0.0, max_val=1.0)
        self.penalty = _check_scalar(penalty, "penalty", min_val=0.0)
        self.stiffness = _check_scalar(stiffness, "stiffness", min_val=0.0, include_min=False)


# the elastic subset of _METRICS (_distance.py:186-205)
_METRICS = {
    "adtw": AmercingDtwMetric,
    "dtw": DtwMetric,
    "ddtw": DerivativeDtwMetric,
    "wdtw": WeightedDtwMetric,
    "wddtw": WeightedDerivativeDtwMetric,
    "lcss": LcssMetric,
    "wlcss": WeightedLcssMetric,
    "erp": ErpMetric,
    "edr": EdrMetric,
    "msm": MsmMetric,
    "twe": TweMetric,
}


def check_metric(metric):
    """Metric class for `metric` (reference: _distance.py:231-260)."""
    if isinstance(metric, str) and metric in _METRICS:
        return _METRICS[metric]
    if callable(metric):
        raise ValueError(
            "callable metrics are not elastic metrics and are not accelerated; use wildboar.distance for them"
        )
    raise ValueError(
        "unsupported metric {}, 'metric' must be callable or a str among {}".format(metric, set(_METRICS.keys()))
    )


def _format_return(x, y_dims, x_dims):
    """_distance.py:514-540."""
    if x_dims == 1 and y_dims == 1 and x.size == 1:
        return x.item()
    elif x_dims == 1 or y_dims == 1:
        return np.squeeze(x)
    else:
        return x


def _make_metric(metric, metric_params):
    Metric = check_metric(metric)
    metric_params = metric_params if metric_params is not None else {}
    return Metric(**metric_params)


def _dim_slices(x_, dim, n_dims):
    """Yield the per-dimension 2-D views the reference iterates over (_distance.py:1289-1302)."""
    if dim in ["mean", "full"]:
        return list(range(n_dims)), dim
    elif isinstance(dim, numbers.Integral) and not isinstance(dim, bool) and 0 <= dim < n_dims:
        return [int(dim)], None
    raise ValueError("The parameter dim must be 0 <= dim < n_dims")


def _nd_call(call, how, n_out, n_dims):
    """Multivariate call with the dimensions combined on the device -- except in the one corner where the device's
    strictly sequential sum over the dimensions is not numpy's order: a SINGLE output element with 8 or more dimensions
    makes ``np.mean(list_of_matrices, axis=0)`` (_distance.py:1250-1253, 1289-1292) a contiguous 1-D reduction, which
    numpy sums pairwise (8 accumulators).  There the per-dimension values are fetched and numpy does the mean."""
    if how == "mean" and n_out == 1 and n_dims >= 8:
        return np.mean(call("full"), axis=0)
    return call(how)


def _combine(distances, how):
    if how == "mean":
        return np.mean(distances, axis=0)
    if how == "full":
        return np.stack(distances, axis=0)
    return distances[0]


# ---------------------------------------------------------------------------------------------
# public API
# ---------------------------------------------------------------------------------------------
def pairwise_distance(x, y=None, *, dim="mean", metric="euclidean", metric_params=None, n_jobs=None):
    """Distance between all pairs of x and y (reference: _distance.py:1178-1304).

    Elastic metrics only: ``metric`` in {dtw, ddtw, wdtw, wddtw, adtw, lcss, wlcss, erp, edr,
    msm, twe}.  ``n_jobs`` is accepted for signature compatibility and ignored (device
    selection: ``set_devices`` / ``WILDBOAR_CUDA_DEVICES``).
    """
    m = _make_metric(metric, metric_params)
    params = m._params()
    if y is None:
        y = x
    if x is y:
        x = check_array(x, allow_3d=True, ensure_2d=False, dtype=float)
        if x.ndim == 1:
            return 0.0
        x_ = _check_ts_array(x)
        n_dims = x.shape[1] if x.ndim == 3 else 1
        if n_dims == 1 and dim == "mean":
            dim = 0
        dims, how = _dim_slices(x_, dim, n_dims)
        if how is not None and n_dims > 1:  # all dimensions in ONE library call (combined on the device)
            return _nd_call(lambda h: _shim.pairwise_nd(m.metric_id, params, x_, None, h), how, x_.shape[0] ** 2, n_dims)
        return _combine([_shim.pairwise(m.metric_id, params, x_[:, d, :], None) for d in dims], how)
    x = check_array(x, allow_3d=True, ensure_2d=False, dtype=np.double)
    y = check_array(y, allow_3d=True, ensure_2d=False, dtype=np.double)
    if x.ndim != 1 and y.ndim != 1 and x.ndim != y.ndim:
        raise ValueError("x (%dD-array) and y (%dD-array) are not compatible" % (x.ndim, y.ndim))
    if x.ndim == 3 and x.shape[1] != y.shape[1]:
        raise ValueError("x and y must have the same number of dimensions.")
    x_ = _check_ts_array(x)
    y_ = _check_ts_array(y)
    n_dims = x.shape[1] if x.ndim == 3 else 1
    if n_dims == 1 and dim == "mean":
        dim = 0
    dims, how = _dim_slices(x_, dim, n_dims)
    if how is not None and n_dims > 1:
        distances = _nd_call(lambda h: _shim.pairwise_nd(m.metric_id, params, x_, y_, h), how, x_.shape[0] * y_.shape[0], n_dims)
    else:
        distances = _combine([_shim.pairwise(m.metric_id, params, x_[:, d, :], y_[:, d, :]) for d in dims], how)
    return _format_return(distances, y.ndim, x.ndim)


def paired_distance(x, y, *, dim="mean", metric="euclidean", metric_params=None, n_jobs=None):
    """Distance between the i:th sample of x and the i:th sample of y (_distance.py:1082-1175).

    As in the reference the operands reach the metric swapped, ``out[i] = metric(y[i], x[i])``
    (``_cdistance.pyx:1632-1647``), which matters for the asymmetric implementations.
    """
    x = check_array(x, allow_3d=True, ensure_2d=False, dtype=float)
    y = check_array(y, allow_3d=True, ensure_2d=False, dtype=float)
    x, y = np.broadcast_arrays(x, y)
    if x.ndim != y.ndim:
        raise ValueError("x (%dD-array) and y (%dD-array) are not compatible." % (x.ndim, y.ndim))
    if x.ndim == 3 and x.shape[1] != y.shape[1]:
        raise ValueError("x and y must have the same number of dimensions.")
    if x.ndim > 1 and y.ndim > 1 and x.shape[0] != y.shape[0]:
        raise ValueError("x and y must have the same number of samples.")
    if n_jobs is not None:
        warnings.warn("n_jobs is not yet supported.", UserWarning)
    m = _make_metric(metric, metric_params)
    params = m._params()
    n_dims = x.shape[1] if x.ndim == 3 else 1
    if n_dims == 1 and dim == "mean":
        dim = 0
    x_ = _check_ts_array(x)
    y_ = _check_ts_array(y)
    dims, how = _dim_slices(x_, dim, n_dims)
    if how is not None and n_dims > 1:
        distances = _nd_call(lambda h: _shim.paired_nd(m.metric_id, params, x_, y_, h), how, x_.shape[0], n_dims)
    else:
        distances = _combine([_shim.paired(m.metric_id, params, x_[:, d, :], y_[:, d, :]) for d in dims], how)
    return _format_return(distances, y.ndim, x.ndim)


def argmin_distance(x, y=None, *, dim=0, k=1, metric="euclidean", metric_params=None, sorted=False,  # noqa: A002
                    return_distance=False, lower_bound=None, n_jobs=None, device_lower_bound=True):
    """Indices of the k samples of y closest to each sample of x (_distance.py:1320-1457).

    Results (indices, distances and their order) equal the reference's sequential
    early-abandoning scan.  ``device_lower_bound`` (extension, dtw only) enables the on-device
    LB_Kim/LB_Keogh pruning cascade; it never changes the result.
    """
    if isinstance(k, bool) or not isinstance(k, numbers.Integral) or k < 1:
        raise ValueError(f"The 'k' parameter of argmin_distance must be an int in the range [1, inf). Got {k!r} instead.")
    if metric_params is not None and not isinstance(metric_params, dict):
        raise ValueError("The 'metric_params' parameter of argmin_distance must be an instance of 'dict' or None.")
    m = _make_metric(metric, metric_params)
    params = m._params()
    x = check_array(x, allow_3d=True, ensure_2d=False, ensure_ts_array=True, dtype=float)
    if y is None:
        y = x
    else:
        y = check_array(y, allow_3d=True, ensure_2d=False, ensure_ts_array=True, dtype=float)
    if x.ndim not in (1, y.ndim):
        raise ValueError(f"x ({x.ndim}d-array) and y ({y.ndim}d-array) are not compatible.")
    if lower_bound is not None:
        lower_bound = check_array(lower_bound, ensure_2d=False)
        lower_bound = np.atleast_2d(lower_bound)
        if x.shape[0] != lower_bound.shape[0] or y.shape[0] != lower_bound.shape[1]:
            raise ValueError(
                "The lower bound must be of shape (x.shape[0], y.shape[0]), got ({}, {})".format(*lower_bound.shape)
            )
    n_dims = x.shape[1] if x.ndim == 3 else 1
    k = min(k, y.shape[0])
    if 0 <= dim < 1:
        # sorted=True: the neighbours are returned by distance, not in heap order, so the library may use its neighbour-set
        # mode (seeded thresholds for 1 < k <= 8; _shim.argmin falls back to the exact scan when the result is not unique)
        indices, distances = _shim.argmin(m.metric_id, params, x[:, dim, :], y[:, dim, :], k, lower_bound,
                                          use_device_lb=bool(device_lower_bound) and m.name in ("dtw", "ddtw", "adtw"),
                                          neighbour_set=bool(sorted), ordered=True)
        if sorted:
            sort = np.argsort(distances, axis=1, kind="stable")
            indices = np.take_along_axis(indices, sort, axis=1)
            if return_distance:
                distances = np.take_along_axis(distances, sort, axis=1)
        if return_distance:
            return indices, distances
        else:
            return indices
    else:
        raise ValueError(f"The parameter dim must be dim ({dim}) < n_dims ({n_dims})")
