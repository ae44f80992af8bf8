"""Host-side mirror of the DTW lower bounds of ``wildboar.distance.lb``.

``DtwKeoghLowerBound`` (reference: src/wildboar/distance/lb.py:314-432) and ``DtwKimLowerBound``
(lb.py:198-311) with the reference's constructor parameters, ``fit`` / ``transform`` /
``fit_transform`` protocol, attribute names (``X_``, ``lower_``, ``upper_``) and errors; the
``(n_queries, n_samples)`` matrices come from the CUDA library (``wb_cuda_lb_keogh`` /
``wb_cuda_lb_kim``, include/wb_cuda.h) instead of the reference's Python double loop and are
bit-equal to it.  The result feeds ``argmin_distance(..., lower_bound=...)`` exactly as in the
reference (lb.py:341-351).  No CPU fallback.
"""
import numbers

import numpy as np

from . import _shim
from .distance import _check_ts_array, check_array

__all__ = ["DtwKeoghLowerBound", "DtwKimLowerBound"]

try:  # same exception type as the reference when scikit-learn is importable
    from sklearn.exceptions import NotFittedError
except Exception:  # pragma: no cover
    class NotFittedError(ValueError, AttributeError):
        pass

_KINDS = {"both": 0, "left": 1, "right": 2}


class _Params:
    """get_params / set_params / fit_transform of sklearn's BaseEstimator + TransformerMixin."""

    _param_names = ()

    def get_params(self, deep=True):
        return {k: getattr(self, k) for k in self._param_names}

    def set_params(self, **params):
        for k, v in params.items():
            if k not in self._param_names:
                raise ValueError(f"Invalid parameter {k!r} for estimator {type(self).__name__}().")
            setattr(self, k, v)
        return self

    def fit_transform(self, X, y=None):
        return self.fit(X, y).transform(X)

    def __repr__(self):
        return "%s(%s)" % (type(self).__name__, ", ".join(f"{k}={getattr(self, k)!r}" for k in self._param_names))


def _warp_size(n_timesteps, r):
    """dtw.py:38-40 `_compute_warp_size` + the `== n_timesteps -> -1` rule of lb.py:367-369."""
    w = max(int(np.floor(n_timesteps * r)), 1)
    return w - 1 if w == n_timesteps else w


class DtwKimLowerBound(_Params):
    """Constant-time DTW lower bound over the first / last three points (lb.py:198-311).

    As in the reference the value is the SUM of squared terms (no square root), i.e. it bounds the
    squared DTW cost."""

    def fit(self, X, y=None):
        self.X_ = check_array(X)
        self.n_features_in_ = self.X_.shape[1]
        return self

    def transform(self, X):
        if not hasattr(self, "X_"):
            raise NotFittedError("This DtwKimLowerBound instance is not fitted yet. Call 'fit' first.")
        Y = check_array(X)
        if Y.shape[1] != self.n_features_in_:
            raise ValueError(f"X has {Y.shape[1]} features, but DtwKimLowerBound is expecting {self.n_features_in_} features as input.")
        return _shim.lb_kim(np.ascontiguousarray(Y), np.ascontiguousarray(self.X_))


class DtwKeoghLowerBound(_Params):
    """LB_Keogh for DTW (lb.py:314-432).

    Parameters
    ----------
    r : float
        The warp window for DTW, in [0, 1].
    kind : {"both", "left", "right"}
        "both": maximum of the two directions; "left": the query against the fitted samples'
        envelopes; "right": the fitted samples against the query's envelope.
    """

    _param_names = ("r", "kind")

    def __init__(self, r=1.0, *, kind="both"):
        self.r = r
        self.kind = kind

    def _validate_params(self):
        if isinstance(self.r, bool) or not isinstance(self.r, numbers.Real) or not (0 <= self.r <= 1):
            raise ValueError(f"The 'r' parameter of DtwKeoghLowerBound must be a float in the range [0.0, 1.0]. Got {self.r!r} instead.")
        if self.kind not in _KINDS:
            raise ValueError(f"The 'kind' parameter of DtwKeoghLowerBound must be a str among {set(_KINDS)}. Got {self.kind!r} instead.")

    def fit(self, X, y=None):
        self._validate_params()
        self.X_ = check_array(X, allow_3d=True, ensure_ts_array=True)
        n_samples, _, n_timesteps = self.X_.shape
        self.n_timesteps_in_ = n_timesteps
        self._env = None
        return self

    def _envelopes(self):
        # lower_/upper_[k] = min/max X_[k-w .. k+w] (clipped), EL:1076-1092; kept for attribute parity
        x = self.X_[:, 0, :]
        w = _warp_size(self.n_timesteps_in_, self.r)
        T = x.shape[1]
        idx = np.arange(T)
        lo = np.empty_like(x); hi = np.empty_like(x)
        for k in idx:
            a, b = max(0, k - w), min(T - 1, k + w)
            lo[:, k] = x[:, a:b + 1].min(axis=1); hi[:, k] = x[:, a:b + 1].max(axis=1)
        return lo, hi

    @property
    def lower_(self):
        """(n_samples, n_timesteps) lower envelope of the fitted samples (lb.py:365-374), on demand."""
        if "X_" not in self.__dict__:
            raise AttributeError("lower_")
        if self._env is None:
            self._env = self._envelopes()
        return self._env[0]

    @property
    def upper_(self):
        if "X_" not in self.__dict__:
            raise AttributeError("upper_")
        if self._env is None:
            self._env = self._envelopes()
        return self._env[1]

    def transform(self, X):
        if "X_" not in self.__dict__:
            raise NotFittedError("This DtwKeoghLowerBound instance is not fitted yet. Call 'fit' first.")
        Y = check_array(X, allow_3d=True, ensure_ts_array=True)
        if Y.shape[2] != self.n_timesteps_in_:
            raise ValueError(f"X has {Y.shape[2]} timesteps, but DtwKeoghLowerBound is expecting {self.n_timesteps_in_} timesteps as input.")
        return _shim.lb_keogh(Y[:, 0, :], self.X_[:, 0, :], self.r, _KINDS[self.kind])
