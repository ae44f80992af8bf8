"""Nearest-neighbour estimators on the CUDA path, with the training set resident on the device.

Mirrors ``wildboar.distance.NearestNeighbors`` / ``KNeighborsClassifier`` (reference:
src/wildboar/distance/_neighbors.py:19-160 and :163-300): same constructor parameters, ``fit`` /
``kneighbors`` / ``predict_proba`` / ``predict`` semantics, same neighbour sets, order and
probabilities -- they come from the same ``argmin_distance`` (heap order, strict ``<`` ties) and
``pairwise_distance(dim="mean")`` + ``np.argpartition`` calls as in the reference.

B200-first difference: ``fit`` uploads ``_fit_X`` once (``wb_cuda_fit``); every later query moves only
the queries and the result.  Elastic metrics only; there is no CPU fallback.
"""
import numbers

import numpy as np

from . import _shim
from .distance import _METRICS, _check_ts_array, _make_metric, check_array

try:  # sklearn is optional: it only contributes get_params / set_params / clone support
    from sklearn.base import BaseEstimator as _SkBase
except Exception:  # pragma: no cover
    class _SkBase:  # minimal stand-in
        def get_params(self, deep=True):
            return {k: getattr(self, k) for k in self._param_names}

        def set_params(self, **params):
            for k, v in params.items():
                if k not in self._param_names:
                    raise ValueError(f"Invalid parameter {k!r} for estimator {type(self).__name__}.")
                setattr(self, k, v)
            return self

try:
    from sklearn.exceptions import NotFittedError as _NotFitted
except Exception:  # pragma: no cover
    class _NotFitted(ValueError, AttributeError):
        pass

__all__ = ["NearestNeighbors", "KNeighborsClassifier", "KMeans", "KMedoids"]


class _NeighborsBase(_SkBase):
    _param_names = ("n_neighbors", "metric", "metric_params", "n_jobs")

    def __init__(self, n_neighbors=5, *, metric="dtw", metric_params=None, n_jobs=None):
        self.n_neighbors = n_neighbors
        self.metric = metric
        self.metric_params = metric_params
        self.n_jobs = n_jobs

    # _parameter_constraints of the reference (_neighbors.py:37-42, 188-193)
    def _validate_params(self):
        k = self.n_neighbors
        if isinstance(k, bool) or not isinstance(k, numbers.Integral) or k < 1:
            raise ValueError(
                f"The 'n_neighbors' parameter of {type(self).__name__} must be an int in the range [1, inf). Got {k!r} instead."
            )
        if not (isinstance(self.metric, str) and self.metric in _METRICS):
            raise ValueError(
                f"The 'metric' parameter of {type(self).__name__} must be a str among {set(_METRICS)} "
                f"(the elastic metrics; others are not accelerated). Got {self.metric!r} instead."
            )
        if self.metric_params is not None and not isinstance(self.metric_params, dict):
            raise ValueError(
                f"The 'metric_params' parameter of {type(self).__name__} must be an instance of 'dict' or None. "
                f"Got {self.metric_params!r} instead."
            )
        if self.n_jobs is not None and (isinstance(self.n_jobs, bool) or not isinstance(self.n_jobs, numbers.Integral)):
            raise ValueError(f"The 'n_jobs' parameter of {type(self).__name__} must be an int or None. Got {self.n_jobs!r} instead.")

    def _fit_x(self, x):
        self._validate_params()
        x = check_array(x, allow_3d=True, dtype=float, input_name="x")
        self.n_timesteps_in_ = x.shape[-1]
        self.n_dims_in_ = x.shape[1] if x.ndim == 3 else 1
        self._fit_X = x.copy()
        self._release()
        self._fitted = _shim.FittedSet(_check_ts_array(self._fit_X))  # resident until refit / release / gc
        return x

    def _release(self):
        f = self.__dict__.pop("_fitted", None)
        if f is not None:
            f.close()

    def release(self):
        """Free the device copy of the training set (it is re-uploaded on the next query)."""
        self._release()

    def __getstate__(self):  # device handles do not pickle; the copy is re-created lazily
        state = dict(self.__dict__)
        state.pop("_fitted", None)
        return state

    def _check_is_fitted(self):
        if not hasattr(self, "_fit_X"):
            raise _NotFitted(
                f"This {type(self).__name__} instance is not fitted yet. Call 'fit' with appropriate arguments before using this estimator."
            )
        if "_fitted" not in self.__dict__:
            self._fitted = _shim.FittedSet(_check_ts_array(self._fit_X))

    def _check_query(self, x):
        x = check_array(x, allow_3d=True, dtype=float, input_name="x")
        n_dims = x.shape[1] if x.ndim == 3 else 1
        if n_dims != self.n_dims_in_:
            raise ValueError(f"X has {n_dims} dimensions, but {type(self).__name__} is expecting {self.n_dims_in_} dimensions as input.")
        if x.shape[-1] != self.n_timesteps_in_:
            raise ValueError(f"X has {x.shape[-1]} timesteps, but {type(self).__name__} is expecting {self.n_timesteps_in_} timesteps as input.")
        # "Treat a multivariate time series with a single dimension as a univariate time series" (_neighbors.py:112-115)
        if x.ndim == 3 and x.shape[1] == 1:
            x = x.reshape(x.shape[0], -1)
        return x

    # the two query forms of the reference, against the resident training set
    def _pairwise_mean(self, x):
        m = _make_metric(self.metric, self.metric_params)
        return _shim.pairwise_fitted(m.metric_id, m._params(), _check_ts_array(x), self._fitted, "mean")

    def _argmin(self, x, k, sorted_, neighbour_set=False):
        # neighbour_set: the caller counts the k nearest (class votes) and does not look at their order, so the library may
        # seed its pruning thresholds for k > 1 as well (wb_cuda.h, use_device_lb bit 1; exact set, repeated with the plain
        # scan when a tie at the kth distance makes the reference's set depend on its scan history)
        m = _make_metric(self.metric, self.metric_params)
        k = min(k, self._fit_X.shape[0])
        idx, dist = _shim.argmin_fitted(m.metric_id, m._params(), _check_ts_array(x)[:, 0, :], self._fitted, k,
                                        use_device_lb=m.name in ("dtw", "ddtw", "adtw"), neighbour_set=neighbour_set or bool(sorted_),
                                        ordered=bool(sorted_))
        if sorted_:
            order = np.argsort(dist, axis=1, kind="stable")
            idx = np.take_along_axis(idx, order, axis=1)
            dist = np.take_along_axis(dist, order, axis=1)
        return idx, dist


class NearestNeighbors(_NeighborsBase):
    """Unsupervised neighbour searches (reference: _neighbors.py:19-160)."""

    def fit(self, x, y=None):
        self._fit_x(x)
        return self

    def kneighbors(self, x=None, n_neighbors=None, return_distance=True):
        self._check_is_fitted()
        if n_neighbors is None:
            n_neighbors = self.n_neighbors
        if x is not None:
            query_is_train = False
            x = self._check_query(x)
        else:
            query_is_train = True
            x = self._fit_X
            if x.ndim == 3 and x.shape[1] == 1:
                x = x.reshape(x.shape[0], -1)

        if x.ndim == 3:
            if query_is_train:
                # `pairwise_distance(x, self._fit_X)` with x IS _fit_X takes the reference's singleton form
                # (upper triangle mirrored, _distance.py:1236-1264) -- it differs for the asymmetric metrics
                m = _make_metric(self.metric, self.metric_params)
                dists = _shim.pairwise_nd(m.metric_id, m._params(), _check_ts_array(x), None, "mean")
                np.fill_diagonal(dists, np.inf)
            else:
                dists = self._pairwise_mean(x)
            sample_range = np.arange(x.shape[0])[:, None]
            neigh_ind = np.argpartition(dists, n_neighbors - 1, axis=1)[:, :n_neighbors]
            neigh_dist = dists[sample_range, neigh_ind]
            sort_inds = np.argsort(neigh_dist, axis=1)
            neigh_ind = neigh_ind[sample_range, sort_inds]
            if return_distance:
                return neigh_dist[sample_range, sort_inds], neigh_ind
            return neigh_ind

        k = n_neighbors + 1 if query_is_train else n_neighbors
        indices, distances = self._argmin(x, k, True)
        if query_is_train:
            mask = indices != np.arange(x.shape[0])[:, None]
            indices = indices[mask].reshape(x.shape[0], -1)
            distances = distances[mask].reshape(x.shape[0], -1)
        if return_distance:
            return distances, indices
        return indices


class KNeighborsClassifier(_NeighborsBase):
    """k-nearest-neighbour classifier (reference: _neighbors.py:163-300)."""

    def fit(self, x, y):
        y = np.asarray(y)
        if y.ndim != 1:
            raise ValueError(f"y should be a 1d array, got an array of shape {y.shape} instead.")
        if y.dtype.kind == "f" and not np.all(y == np.floor(y)):
            raise ValueError("Unknown label type: continuous. Maybe you are trying to fit a classifier, which expects discrete classes on a regression target with continuous values.")
        x = check_array(x, allow_3d=True, dtype=float, input_name="x")
        if x.shape[0] != y.shape[0]:
            raise ValueError(f"Found input variables with inconsistent numbers of samples: [{x.shape[0]}, {y.shape[0]}]")
        self._fit_x(x)
        self.classes_, self._y = np.unique(y, return_inverse=True)
        return self

    def predict_proba(self, x):
        self._check_is_fitted()
        x = self._check_query(x)
        if x.ndim == 3:
            dists = self._pairwise_mean(x)
            closest = np.argpartition(dists, self.n_neighbors, axis=1)[:, : self.n_neighbors]
        else:
            closest, _ = self._argmin(x, self.n_neighbors, False, neighbour_set=True)  # only counted below
        preds = self._y[closest]
        probs = np.empty((x.shape[0], len(self.classes_)), dtype=float)
        for i in range(len(self.classes_)):
            probs[:, i] = np.sum(preds == i, axis=1) / self.n_neighbors
        return probs

    def predict(self, x):
        proba = np.argmax(self.predict_proba(x), axis=1)
        return np.take(self.classes_, proba)

    def score(self, x, y):
        """Mean accuracy (sklearn.base.ClassifierMixin.score)."""
        return float(np.mean(self.predict(x) == np.asarray(y)))


class KMeans(_SkBase):
    """KMeans clustering with DTW / weighted-DTW barycentres (reference: _neighbors.py:303-640, metric="dtw").

    Same parameters (except that only ``metric="dtw"`` is accelerated -- euclidean k-means is not an elastic
    path), same random-number consumption, same ``cluster_centers_`` / ``labels_`` / ``inertia_`` /
    ``n_iter_`` as the reference.  B200-first: assignment and cost are one ``argmin`` / ``paired`` call each,
    and the barycentre update of ALL clusters runs as batched DBA epochs on the device
    (``wildboar_b200.dtw.dtw_average_many``) against a sample set that is uploaded once per ``fit``.
    """

    _param_names = ("n_clusters", "metric", "r", "g", "init", "n_init", "max_iter", "tol", "verbose", "random_state")

    def __init__(self, n_clusters=8, *, metric="dtw", r=1.0, g=None, init="random", n_init="auto", max_iter=300,
                 tol=1e-3, verbose=0, random_state=None):
        self.metric = metric
        self.r = r
        self.g = g
        self.n_clusters = n_clusters
        self.init = init
        self.n_init = n_init
        self.max_iter = max_iter
        self.tol = tol
        self.verbose = verbose
        self.random_state = random_state

    def _validate_params(self):
        name = type(self).__name__

        def bad(param, what):
            return ValueError(f"The {param!r} parameter of {name} must be {what}. Got {getattr(self, param)!r} instead.")

        def is_int(v):
            return isinstance(v, numbers.Integral) and not isinstance(v, bool)

        if not is_int(self.n_clusters) or self.n_clusters < 1:
            raise bad("n_clusters", "an int in the range [1, inf)")
        if self.metric != "dtw":
            raise bad("metric", "'dtw' (euclidean k-means is not an elastic path and is not accelerated)")
        if isinstance(self.r, bool) or not isinstance(self.r, numbers.Real) or not 0 <= self.r <= 1:
            raise bad("r", "a float in the range [0, 1]")
        if self.g is not None and (isinstance(self.g, bool) or not isinstance(self.g, numbers.Real) or not self.g > 0):
            raise bad("g", "None or a float in the range (0, inf)")
        if self.init != "random":
            raise bad("init", "a str among {'random'}")
        if not (self.n_init == "auto" or (is_int(self.n_init) and self.n_init >= 1)):
            raise bad("n_init", "a str among {'auto'} or an int in the range [1, inf)")
        if not is_int(self.max_iter) or self.max_iter < 1:
            raise bad("max_iter", "an int in the range [1, inf)")
        if not isinstance(self.tol, float):
            raise bad("tol", "an instance of 'float'")

    def _metric(self):
        if self.g is None:
            return "dtw", {"r": self.r}
        return "wdtw", {"r": self.r, "g": self.g}

    def fit(self, x, y=None):
        from .dtw import _check_random_state
        self._validate_params()
        x = check_array(x, allow_3d=False, dtype=float, input_name="x")
        self.n_timesteps_in_ = x.shape[-1]
        n_init = 1 if self.n_init == "auto" else self.n_init
        random_state = _check_random_state(self.random_state)
        fitted = _shim.FittedSet(x.reshape(x.shape[0], 1, x.shape[1]), devices=[_shim._first_device()])
        try:
            best = None
            best_cost = np.inf
            for _ in range(n_init):
                run = self._fit_one_init(x, fitted, random_state)
                if run["cost"] < best_cost:
                    best_cost, best = run["cost"], run
        finally:
            fitted.close()
        self.n_iter_ = best["iter"]
        self.inertia_ = best_cost
        self.cluster_centers_ = best["centroids"]
        if best["reassign"]:
            best["assigned"], best["distance"] = self._assign(x, best["centroids"])
        self.labels_ = best["assigned"]
        return self

    # _KMeansCluster.assign / cost (_neighbors.py:303-347)
    def _assign(self, x, centroids):
        from .distance import argmin_distance
        metric, mp = self._metric()
        assigned, distance = argmin_distance(x, centroids, k=1, metric=metric, metric_params=mp, return_distance=True)
        return np.ravel(assigned), distance

    def _cost(self, x, centroids, assigned):
        from .distance import paired_distance
        metric, mp = self._metric()
        return paired_distance(x, centroids[assigned], dim="mean", metric=metric, metric_params=mp).sum() / x.shape[0]

    def _update(self, x, fitted, centroids, assigned, distance, random_state):
        """`_KMeansCluster.update` (_neighbors.py:349-361): the cluster loop only does the bookkeeping (membership
        snapshots, empty / singleton clusters, one `randint` per barycentre as in `_DtwCluster._update_centroid`);
        the barycentres themselves are computed afterwards, all clusters per device step."""
        from .dtw import dtw_average_many
        jobs = []
        for c in range(centroids.shape[0]):
            members = np.flatnonzero(assigned == c)
            if members.shape[0] == 0:
                far_from_center = distance.min(axis=1).argmax()
                assigned[far_from_center] = c
                centroids[c] = x[far_from_center]
            elif members.shape[0] == 1:
                centroids[c] = x[members[0]]
            else:
                random_state.randint(np.iinfo(np.int32).max)  # consumed by the reference, unused by method="mm"
                jobs.append((c, members))
        if jobs:
            means, _ = dtw_average_many(x, [m for _, m in jobs], [centroids[c] for c, _ in jobs], r=self.r, g=self.g,
                                        fitted=fitted)
            for (c, _), mean in zip(jobs, means):
                centroids[c] = mean

    def _fit_one_init(self, x, fitted, random_state):
        import math
        centroids = x[random_state.choice(x.shape[0], size=self.n_clusters, replace=False)]
        prev_cost, cost, reassign = np.inf, -np.inf, True
        assigned = distance = None
        for it in range(self.max_iter):
            assigned, distance = self._assign(x, centroids)
            prev_cost, cost = cost, self._cost(x, centroids, assigned)
            if self.verbose > 0:
                print(f"Iteration {it}, {cost} (prev_cost = {prev_cost})")
            if math.isclose(cost, prev_cost, rel_tol=self.tol):
                reassign = False
                break
            self._update(x, fitted, centroids, assigned, distance, random_state)
        return dict(iter=it, cost=cost, centroids=centroids, assigned=assigned, distance=distance, reassign=reassign)

    def transform(self, x):
        from .distance import pairwise_distance
        if not hasattr(self, "cluster_centers_"):
            raise _NotFitted(f"This {type(self).__name__} instance is not fitted yet. Call 'fit' with appropriate arguments before using this estimator.")
        x = check_array(x, allow_3d=False, dtype=float, input_name="x")
        metric, mp = self._metric()
        return pairwise_distance(x, self.cluster_centers_, dim="mean", metric=metric, metric_params=mp)

    def predict(self, x):
        return self.transform(x).argmin(axis=1)

    def fit_predict(self, x, y=None):
        return self.fit(x).labels_


# ---------------------------------------------------------------------------------------------
# KMedoids (reference: _neighbors.py:615-930 + the PAM helpers of _cneighbors.pyx)
# ---------------------------------------------------------------------------------------------
def _seq_sum(v):
    """Sum in index order (one rounding per addition), as the reference's C loops accumulate."""
    return float(np.cumsum(v)[-1]) if len(v) else 0.0


def _pam_build(D, n_clusters):
    """BUILD initialisation, `_pam_build` of _cneighbors.pyx:70-108: greedy, cost changes accumulated in index order,
    `>=` keeps the LAST best candidate."""
    n = len(D)
    medoids = np.zeros(n_clusters, dtype=np.intp)
    medoids[0] = np.argmin(np.sum(D, axis=0))
    not_medoids = np.delete(np.arange(n, dtype=np.intp), medoids[0])
    Dj = D[medoids[0]].copy()
    for cur in range(1, n_clusters):
        best, best_val = (0, 0), 0.0
        djn = Dj[not_medoids]
        for i, id_i in enumerate(not_medoids):
            cc = _seq_sum(np.maximum(0.0, djn - D[id_i, not_medoids]))
            if cc >= best_val:
                best_val, best = cc, (id_i, i)
        medoids[cur] = best[0]
        not_medoids = np.delete(not_medoids, best[1])
        Dj = np.minimum(Dj, D[:, best[0]])
    return medoids


def _pam_optimal_swap(D, medoids, not_medoids, Djs, Ejs, n_clusters):
    """SWAP step, `_pam_optimal_swap` of _cneighbors.pyx:13-67: the same four cases per non-medoid point, summed in index
    order; strict `<` keeps the FIRST best swap (h outer, i inner)."""
    best = (1, 1, 0.0)
    djn, ejn = Djs[not_medoids], Ejs[not_medoids]
    for id_h in not_medoids:
        dh_row = D[id_h, not_medoids]   # D[id_h, id_j]
        dh_col = D[not_medoids, id_h]   # D[id_j, id_h]
        second = dh_row < ejn
        gain = dh_col - djn
        for id_i in medoids:
            in_i = D[id_i, not_medoids] == djn
            terms = np.where(in_i & second, gain, np.where(in_i & ~second, ejn - djn, np.where(~in_i & (dh_col < djn), gain, 0.0)))
            cc = _seq_sum(terms)
            cc += D[id_i, id_h] if D[id_h, id_i] < Ejs[id_i] else Ejs[id_i]
            if cc < best[2]:
                best = (int(id_i), int(id_h), float(cc))
    return best if best[2] < 0 else None


class KMedoids(_SkBase):
    """KMedoids on an elastic metric (reference: _neighbors.py:677-930).

    Same parameters, random-number consumption, ``medoid_indices_`` / ``labels_`` / ``inertia_`` / ``n_iter_`` /
    ``cluster_centers_`` as the reference.  The one expensive step -- the n x n distance matrix
    ``pairwise_distance(x, dim="mean")`` -- is one self join on the device (lower triangle mirrored by the kernel); the
    medoid updates are the reference's matrix bookkeeping ("fast": per-cluster row sums; "pam": the BUILD / SWAP loops of
    _cneighbors.pyx restated with their summation order).  ``transform`` / ``predict`` compare against the medoids only.
    """

    _param_names = ("n_clusters", "metric", "metric_params", "init", "n_init", "algorithm", "max_iter", "tol", "verbose",
                    "n_jobs", "random_state")

    def __init__(self, n_clusters=8, metric="dtw", metric_params=None, init="random", n_init="auto", algorithm="fast",
                 max_iter=30, tol=1e-4, verbose=0, n_jobs=None, random_state=None):
        self.n_clusters = n_clusters
        self.metric = metric
        self.metric_params = metric_params
        self.init = init
        self.n_init = n_init
        self.algorithm = algorithm
        self.max_iter = max_iter
        self.tol = tol
        self.verbose = verbose
        self.n_jobs = n_jobs
        self.random_state = random_state

    def _validate_params(self):
        name = type(self).__name__

        def bad(param, what):
            return ValueError(f"The {param!r} parameter of {name} must be {what}. Got {getattr(self, param)!r} instead.")

        def is_int(v):
            return isinstance(v, numbers.Integral) and not isinstance(v, bool)

        if not is_int(self.n_clusters) or self.n_clusters < 1:
            raise bad("n_clusters", "an int in the range [1, inf)")
        if not (isinstance(self.metric, str) and (self.metric in _METRICS or self.metric == "precomputed")):
            raise bad("metric", f"a str among {set(_METRICS) | {'precomputed'}} (the elastic metrics; others are not accelerated)")
        if self.metric_params is not None and not isinstance(self.metric_params, dict):
            raise bad("metric_params", "an instance of 'dict' or None")
        if self.init not in ("random", "auto", "min"):
            raise bad("init", "a str among {'auto', 'min', 'random'}")
        if not (self.n_init == "auto" or (is_int(self.n_init) and self.n_init >= 1)):
            raise bad("n_init", "a str among {'auto'} or an int in the range [1, inf)")
        if self.algorithm not in ("fast", "pam"):
            raise bad("algorithm", "a str among {'fast', 'pam'}")
        if not is_int(self.max_iter) or self.max_iter < 1:
            raise bad("max_iter", "an int in the range [1, inf)")
        if isinstance(self.tol, bool) or not isinstance(self.tol, numbers.Real) or self.tol < 0:
            raise bad("tol", "a float in the range [0.0, inf)")

    def fit(self, x, y=None):
        import warnings

        from .distance import pairwise_distance
        from .dtw import _check_random_state
        self._validate_params()
        x = check_array(x, allow_3d=True, dtype=float, input_name="x")
        self.n_timesteps_in_ = x.shape[-1]
        self.n_dims_in_ = x.shape[1] if x.ndim == 3 else 1
        n_init = (1 if self.init == "auto" else 10) if self.n_init == "auto" else self.n_init
        if n_init > 1 and self.init != "random":  # initial medoids are deterministic
            n_init = 1
        max_iter = self.max_iter
        if self.algorithm == "pam" and self.n_clusters == 1 and self.max_iter != 0:
            warnings.warn("n_clusters must be larger than 1 if max_iter larger than 0")
            max_iter = 0
        random_state = _check_random_state(self.random_state)
        if self.metric == "precomputed":
            dist = x
        else:
            dist = pairwise_distance(x, dim="mean", metric=self.metric, metric_params=self.metric_params)
        best_iter, best_cost, best, best_reassign, last = 0, np.inf, None, None, None
        for _ in range(n_init):
            it, last, reassign = self._fit_one_init(dist, max_iter=max_iter, random_state=random_state)
            if last["cost"] < best_cost:
                best_cost, best, best_iter, best_reassign = last["cost"], last, it, reassign
        if best_reassign:
            self._assign(dist, last)  # (the reference re-assigns the LAST clusterer here, _neighbors.py:834-835)
        self.cluster_centers_ = None if self.metric == "precomputed" else x[best["medoids"]]
        self.inertia_ = best["cost"]
        self.medoid_indices_ = best["medoids"]
        self.n_iter_ = best_iter
        self.labels_ = best["labels"]
        return self

    # _KMedoidsCluster (reference :615-633)
    @staticmethod
    def _assign(dist, c):
        c["labels"] = dist[c["medoids"]].argmin(axis=0)

    @staticmethod
    def _cost(dist, c):
        c["cost"] = np.take(dist, c["medoids"][c["labels"]]).sum() / dist.shape[0]

    def _update(self, dist, c):
        med = c["medoids"]
        if self.algorithm == "fast":  # _FastKMedoidsCluster.update, :635-647
            for idx in range(med.shape[0]):
                members = np.where(c["labels"] == idx)[0]
                if members.shape[0] == 0:
                    continue
                cost = dist[members, members.reshape(-1, 1)].sum(axis=1)
                min_idx = cost.argmin()
                if cost[min_idx] < cost[(members == med[idx]).argmax()]:
                    med[idx] = members[min_idx]
        else:  # _PamKMedoidsCluster.update, :655-672
            not_med = np.delete(np.arange(dist.shape[0]), med)
            swap = _pam_optimal_swap(dist, med, not_med, c["djs"], c["ejs"], med.shape[0])
            if swap is not None:
                i, j, _ = swap
                med[med == i] = j
                c["djs"], c["ejs"] = np.sort(dist[med], axis=0)[[0, 1]]

    def _fit_one_init(self, dist, *, max_iter, random_state):
        import math
        import warnings
        c = {"medoids": np.asarray(self._init_centers(dist, self.n_clusters, random_state)).copy()}
        if self.algorithm == "pam":
            c["djs"], c["ejs"] = np.sort(dist[c["medoids"]], axis=0)[[0, 1]]
        self._assign(dist, c)
        reassign = False
        prev_cost = np.inf
        it = 0
        for it in range(max_iter):
            self._update(dist, c)
            self._cost(dist, c)
            if self.verbose:
                print(f"Iteration {it}/{self.max_iter}: current_cost={c['cost']}")
            if math.isclose(c["cost"], prev_cost, rel_tol=self.tol):
                reassign = True
                break
            prev_cost = c["cost"]
            self._assign(dist, c)
        if max_iter == 0:
            self._cost(dist, c)
        if it + 1 == self.max_iter:
            try:
                from sklearn.exceptions import ConvergenceWarning
            except Exception:  # pragma: no cover
                ConvergenceWarning = UserWarning
            warnings.warn("Maximum number of iterations reached before convergence. Consider increasing max_iter to improve the fit",
                          ConvergenceWarning)
        return it, c, reassign

    def _init_centers(self, dist, n_clusters, random_state):
        if self.init == "random":
            return random_state.choice(dist.shape[0], n_clusters, replace=False)
        if self.init == "min" or (self.algorithm == "fast" and self.init == "auto"):
            return np.argpartition(dist.sum(axis=1), n_clusters)[:n_clusters]
        return _pam_build(dist, n_clusters)

    def transform(self, x):
        from .distance import pairwise_distance
        if not hasattr(self, "medoid_indices_"):
            raise _NotFitted(f"This {type(self).__name__} instance is not fitted yet. Call 'fit' with appropriate arguments before using this estimator.")
        x = check_array(x, allow_3d=True, dtype=float, input_name="x")
        if self.metric == "precomputed":
            return x[:, self.medoid_indices_]
        if x.shape[-1] != self.n_timesteps_in_:
            raise ValueError(f"X has {x.shape[-1]} timesteps, but {type(self).__name__} is expecting {self.n_timesteps_in_} timesteps as input.")
        return pairwise_distance(x, self.cluster_centers_, dim="mean", metric=self.metric, metric_params=self.metric_params)

    def predict(self, x):
        return self.transform(x).argmin(axis=1)

    def fit_predict(self, x, y=None):
        return self.fit(x).labels_
