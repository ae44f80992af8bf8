"""Batched pivot distances for proximity trees (last item).

The reference's ``ProximityTreeClassifier`` (src/wildboar/tree/_cptree.pyx) spends its time in two scalar loops over an
elastic metric:

* ``_partition_pivots`` (_cptree.pyx:887-910), at every candidate split of every node: each sample of the node goes to the
  branch of the nearest pivot, ``distance(metric, X, pivots[p], X, j)`` -- the PIVOT is the first operand;
* ``find_min_branch`` (_cptree.pyx:273-289), at prediction time for every node on a sample's path:
  ``_distance(metric, sample, pivot.data[b])`` -- the SAMPLE is the first operand.

Both are "nearest of a few series under one metric, first minimum wins" (strict ``<`` from +inf), which is one pairwise launch
on the device for ALL samples of the node at once (``n_branches x n_samples`` pairs; a node with few samples lands on the
cooperative engine, a large one on the strip engine).  The operand order matters for the asymmetric implementations (msm,
wdtw's row-0 weights, adtw, erp with a window): each helper keeps the order of the loop it replaces.  Only these two hot loops
are provided -- the tree itself (random metric / pivot sampling, impurity, recursion) is host-side bookkeeping that a
wildboar build keeps as is.
"""
import numpy as np

from .distance import pairwise_distance

__all__ = ["partition_pivots", "find_min_branch"]


def _first_min(d, axis):
    # strict `<` against a running minimum that starts at +inf: the first minimum wins; a row of +inf / NaN distances
    # assigns no branch (-1)
    idx = np.argmin(np.where(np.isnan(d), np.inf, d), axis=axis)
    best = np.take_along_axis(d, np.expand_dims(idx, axis), axis).squeeze(axis)
    return np.where(best < np.inf, idx, -1).astype(np.intp)


def partition_pivots(x, samples, pivots, *, metric, metric_params=None, return_distance=False):
    """Branch of every sample of a node: ``argmin_p metric(x[pivots[p]], x[samples[i]])`` (_cptree.pyx:887-910).

    x : (n_samples_total, n_timestep) array; samples : indices of the node's samples; pivots : indices of the branch
    exemplars (one per branch).  Returns the branch index per sample (and the ``(n_samples, n_branches)`` distances)."""
    x = np.asarray(x, dtype=float)
    samples = np.asarray(samples, dtype=np.intp)
    pivots = np.asarray(pivots, dtype=np.intp)
    d = np.atleast_2d(pairwise_distance(np.ascontiguousarray(x[pivots]), np.ascontiguousarray(x[samples]), dim=0, metric=metric,
                                        metric_params=metric_params))
    d = d.reshape(len(pivots), len(samples)).T
    branch = _first_min(d, axis=1)
    return (branch, d) if return_distance else branch


def find_min_branch(pivot_data, samples, *, metric, metric_params=None, return_distance=False):
    """Branch of every sample at a fitted node: ``argmin_b metric(samples[i], pivot_data[b])`` (_cptree.pyx:273-289)."""
    pivot_data = np.atleast_2d(np.asarray(pivot_data, dtype=float))
    samples = np.atleast_2d(np.asarray(samples, dtype=float))
    d = np.atleast_2d(pairwise_distance(samples, pivot_data, dim=0, metric=metric, metric_params=metric_params))
    d = d.reshape(len(samples), len(pivot_data))
    branch = _first_min(d, axis=1)
    return (branch, d) if return_distance else branch
