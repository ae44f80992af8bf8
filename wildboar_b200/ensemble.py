"""Elastic ensemble classifier on the CUDA path.

Mirrors ``wildboar.ensemble.ElasticEnsembleClassifier`` (reference: src/wildboar/ensemble/_elastic.py): one
``KNeighborsClassifier`` per elastic metric, its ``metric_params`` chosen by leave-one-out cross-validation over a
parameter grid (``GridSearchCV(..., cv=LeaveOneOut())``), predictions weighted by the cross-validation scores.
Same grids (``make_parameter_grid``, distance/_multi_metric.py:26-51), same scores, same chosen parameters, same
probabilities.

B200-first: the reference fits and queries ``n_samples x n_candidates`` one-sample folds.  Here every candidate is ONE
library call over all folds at once -- leave-one-out is the ``argmin`` scan of every sample against the whole training
set with the sample itself masked by the ``lower_bound`` argument (``+inf`` on the diagonal: the scan skips it exactly
where the fold's training set lacks it, so thresholds, abandoning and neighbours are those of the fold) -- and the
final estimators keep their training set resident on the device.  Elastic metrics only; no CPU fallback.
"""
import itertools
import numbers
import re
from collections import defaultdict

import numpy as np

from . import _shim
from .distance import _METRICS, _check_ts_array, _make_metric, check_array
from .neighbors import KNeighborsClassifier, _NotFitted, _SkBase

__all__ = ["ElasticEnsembleClassifier", "make_parameter_grid"]


def parse_metric_spec(kwargs):
    """distance/_multi_metric.py:11-23."""
    specs = defaultdict(dict)
    for key, value in kwargs.items():
        m = re.match(r"^(num|max|min)_([a-zA-Z_$]\w*)$", key)
        if m:
            specs[m.group(2)][m.group(1)] = value
        else:
            raise ValueError(f"The parameter {key} must be prefixed with 'min_', 'max_' or 'num_', got {key} ")
    return specs


def make_parameter_grid(metric_spec, default_n=10):
    """distance/_multi_metric.py:26-51."""
    if metric_spec is None:
        return [{}]
    specs = parse_metric_spec(metric_spec)
    params, grids = [], []
    for param, spec in specs.items():
        if "max" not in spec:
            raise ValueError(f"The maximum value is missing for {param}.")
        if "min" not in spec:
            raise ValueError(f"The minimum value is missing for {param}.")
        num = spec["num"] if "num" in spec else default_n
        params.append(param)
        grids.append(np.linspace(spec["min"], spec["max"], num))
    return [{param: value for param, value in zip(params, grid)} for grid in itertools.product(*grids)]


def _make_elastic_parameter_grid(std):
    """ensemble/_elastic.py:15-49."""
    return {
        "dtw": {"min_r": 0.01, "max_r": 0.3, "num_r": 10},
        "adtw": {"min_r": 0.01, "max_r": 0.3, "num_r": 3, "min_p": 1, "max_p": 4, "num_p": 3},
        "ddtw": {"min_r": 0.01, "max_r": 0.3, "num_r": 10},
        "wdtw": {"min_g": 0.01, "max_g": 0.5, "num_g": 10},
        "wddtw": {"min_g": 0.01, "max_g": 0.5, "num_g": 10},
        "lcss": {"min_r": 0.0, "max_r": 0.25, "num_r": 3, "min_epsilon": 0.2 * std, "max_epsilon": std, "num_epsilon": 3},
        "erp": {"min_g": 0, "max_g": 1.0, "num_g": 10},
        "msm": {"min_c": 0.01, "max_c": 100, "num_c": 10},
        "twe": {"min_penalty": 1e-5, "max_penalty": 1.0, "num_penalty": 3,
                "min_stiffness": 1e-6, "max_stiffness": 0.1, "num_stiffness": 3},
    }


def _loo_neighbors(x, metric, metric_params, k):
    """Neighbours of every sample among the OTHER samples, exactly as the leave-one-out folds of the reference
    compute them (KNeighborsClassifier.predict_proba, distance/_neighbors.py:262-283), for all folds in one call."""
    n = x.shape[0]
    m = _make_metric(metric, metric_params)
    if x.ndim == 3 and x.shape[1] > 1:
        # multivariate folds: pairwise(dim="mean") of the sample against the fold's training set, then
        # argpartition(dists, n_neighbors)[:n_neighbors] -- the row of the full matrix without its own column
        dists = _shim.pairwise_nd(m.metric_id, m._params(), _check_ts_array(x), _check_ts_array(x), "mean")
        out = np.empty((n, k), dtype=np.intp)
        others = np.arange(n)
        for i in range(n):
            keep = others != i
            closest = np.argpartition(dists[i, keep][None, :], k, axis=1)[:, :k]
            out[i] = others[keep][closest[0]]
        return out
    xd = _check_ts_array(x)[:, 0, :]
    mask = np.full((n, n), -np.inf)
    np.fill_diagonal(mask, np.inf)
    idx, _ = _shim.argmin(m.metric_id, m._params(), xd, xd, min(k, n - 1), lower_bound=mask,
                          use_device_lb=m.name in ("dtw", "ddtw", "adtw"))
    return idx


def _loo_score(y, closest):
    """Mean leave-one-out accuracy (GridSearchCV's mean_test_score over the n one-sample folds)."""
    n = y.shape[0]
    hits = 0
    for i in range(n):
        fold_classes = np.unique(np.delete(y, i))                 # classes_ of the fold's classifier
        votes = y[closest[i]]
        counts = np.array([np.sum(votes == c) for c in fold_classes])
        hits += int(fold_classes[np.argmax(counts)] == y[i])     # argmax: first (smallest) class wins ties
    return np.float64(hits) / n  # == np.mean of the n 0/1 fold scores (the sum of small integers is exact)


class ElasticEnsembleClassifier(_SkBase):
    """Ensemble of nearest-neighbour classifiers over the elastic metrics (ensemble/_elastic.py:63-224).

    ``metric``: "auto" / "elastic" (the reference's grid over dtw, adtw, ddtw, wdtw, wddtw, lcss, erp, msm, twe) or a
    dict ``{metric: grid spec}``; the reference's "non_elastic" / "all" contain metrics that are not on this path.
    """

    _param_names = ("n_neighbors", "metric", "n_jobs")

    def __init__(self, n_neighbors=1, *, metric="auto", n_jobs=None):
        self.n_neighbors = n_neighbors
        self.metric = metric
        self.n_jobs = n_jobs

    def fit(self, x, y):
        k = self.n_neighbors
        if isinstance(k, bool) or not isinstance(k, numbers.Integral) or k < 1:
            raise ValueError(f"The 'n_neighbors' parameter of ElasticEnsembleClassifier must be an int in the range [1, inf). Got {k!r} instead.")
        x = check_array(x, allow_3d=True, dtype=float, input_name="x")
        y = np.asarray(y)
        if y.ndim != 1 or y.shape[0] != x.shape[0]:
            raise ValueError(f"Found input variables with inconsistent numbers of samples: [{x.shape[0]}, {y.shape[0] if y.ndim else 0}]")
        self.classes_ = np.unique(y)
        if len(self.classes_) < 2:
            raise ValueError("too few labels")
        if isinstance(self.metric, str):
            if self.metric in ("elastic", "auto"):
                metric = _make_elastic_parameter_grid(x.std())
            elif self.metric in ("non_elastic", "all"):
                raise ValueError(f"metric={self.metric!r} contains non-elastic metrics, which are not accelerated; use wildboar.ensemble for them")
            else:
                raise ValueError(f"The 'metric' parameter of ElasticEnsembleClassifier must be a dict or a str among {{'auto', 'elastic'}}. Got {self.metric!r} instead.")
        elif isinstance(self.metric, dict):
            metric = self.metric
        else:
            raise ValueError(f"The 'metric' parameter of ElasticEnsembleClassifier must be a dict or a str. Got {self.metric!r} instead.")

        metric_param_grid = {}
        for metric_name, param_grid in metric.items():
            if metric_name not in _METRICS:
                raise ValueError(f"{metric_name} is not supported")
            metric_param_grid[metric_name] = make_parameter_grid(param_grid)

        self.estimators_ = []
        self.scores_ = []
        self.cv_results_ = {}
        for metric_name, candidates in metric_param_grid.items():
            scores = np.array([_loo_score(y, _loo_neighbors(x, metric_name, params, k)) for params in candidates])
            best = int(np.argmax(scores))  # GridSearchCV: rank 1, first candidate among ties
            estimator = KNeighborsClassifier(n_neighbors=k, metric=metric_name, metric_params=candidates[best],
                                             n_jobs=self.n_jobs).fit(x, y)
            self.estimators_.append(estimator)
            self.scores_.append((metric_name, scores[best]))
            self.cv_results_[metric_name] = {"params": candidates, "mean_test_score": scores}
        return self

    def predict_proba(self, x):
        if not hasattr(self, "estimators_"):
            raise _NotFitted("This ElasticEnsembleClassifier instance is not fitted yet. Call 'fit' with appropriate arguments before using this estimator.")
        x = check_array(x, allow_3d=True, dtype=float, input_name="x")
        proba = np.zeros((x.shape[0], len(self.estimators_), len(self.classes_)))
        score_sum = 0
        for i, ((_, score), estimator) in enumerate(zip(self.scores_, self.estimators_)):
            proba[:, i] = estimator.predict_proba(x) * score
            score_sum += score
        return proba.sum(axis=1) / score_sum

    def predict(self, x):
        return np.take(self.classes_, np.argmax(self.predict_proba(x), axis=1))

    def score(self, x, y):
        return float(np.mean(self.predict(x) == np.asarray(y)))
