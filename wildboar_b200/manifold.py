"""Multidimensional scaling on an elastic dissimilarity (reference: src/wildboar/distance/_manifold.py:10-132).

The reference's ``MDS`` is scikit-learn's SMACOF on a precomputed matrix; the only expensive step that belongs to this path is
that matrix, ``pairwise_distance(x, dim="mean", metric=..., metric_params=...)`` -- one self join on the device here (lower
triangle mirrored by the kernel, result in page-locked memory).  Same constructor, same ``fit`` / ``fit_transform``, the same
``mds_`` attribute; with a bit-equal dissimilarity matrix and the same ``random_state`` the embedding is the reference's.
"""
import numbers

from .distance import _METRICS, check_array, pairwise_distance

try:
    from sklearn.base import BaseEstimator as _SkBase
except Exception:  # pragma: no cover
    _SkBase = object

__all__ = ["MDS"]


class MDS(_SkBase):
    def __init__(self, n_components=2, *, metric_mds=True, n_init=4, max_iter=300, verbose=0, eps=1e-3, n_jobs=None,
                 random_state=None, metric="dtw", metric_params=None, normalized_stress="auto"):
        self.n_components = n_components
        self.metric_mds = metric_mds
        self.n_init = n_init
        self.max_iter = max_iter
        self.verbose = verbose
        self.eps = eps
        self.n_jobs = n_jobs
        self.random_state = random_state
        self.normalized_stress = normalized_stress
        self.metric = metric
        self.metric_params = metric_params

    def _validate_params(self):
        name = type(self).__name__

        def bad(param, what):
            return ValueError(f"The {param!r} parameter of {name} must be {what}. Got {getattr(self, param)!r} instead.")

        def is_int(v):
            return isinstance(v, numbers.Integral) and not isinstance(v, bool)

        for p in ("n_components", "n_init", "max_iter"):
            if not is_int(getattr(self, p)) or getattr(self, p) < 1:
                raise bad(p, "an int in the range [1, inf)")
        if not isinstance(self.metric_mds, bool):
            raise bad("metric_mds", "an instance of 'bool'")
        if isinstance(self.eps, bool) or not isinstance(self.eps, numbers.Real) or self.eps < 0:
            raise bad("eps", "a float in the range [0.0, inf)")
        if not (isinstance(self.metric, str) and self.metric in _METRICS):
            raise bad("metric", f"a str among {set(_METRICS)} (the elastic metrics; others are not accelerated)")
        if self.metric_params is not None and not isinstance(self.metric_params, dict):
            raise bad("metric_params", "an instance of 'dict' or None")
        if not (isinstance(self.normalized_stress, bool) or self.normalized_stress == "auto"):
            raise bad("normalized_stress", "an instance of 'bool' or a str among {'auto'}")

    def fit(self, x, y=None):
        self.fit_transform(x)
        return self

    def fit_transform(self, x, y=None):
        from sklearn.manifold import MDS as Sklearn_MDS  # noqa: N811
        self._validate_params()
        x = check_array(x, allow_3d=True, dtype=float, input_name="x")
        self.n_timesteps_in_ = x.shape[-1]
        self.mds_ = Sklearn_MDS(n_components=self.n_components, metric_mds=self.metric_mds, n_init=self.n_init,
                                max_iter=self.max_iter, verbose=self.verbose, eps=self.eps, n_jobs=self.n_jobs,
                                random_state=self.random_state, metric="precomputed", normalized_stress=self.normalized_stress)
        return self.mds_.fit_transform(pairwise_distance(x, dim="mean", metric=self.metric, metric_params=self.metric_params))
