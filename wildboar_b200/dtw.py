"""DTW alignments, warping paths and DTW barycentre averaging (DBA) on the CUDA path.

Mirrors the part of ``wildboar.distance.dtw`` (reference: src/wildboar/distance/dtw.py) that sits on
the elastic hot path: ``dtw_distance`` / ``wdtw_distance`` / ``ddtw_distance`` / ``wddtw_distance``
(:43-152), ``dtw_alignment`` / ``wdtw_alignment`` (:246-345), ``jeong_weight`` (:347-374),
``dtw_mapping`` (:377-418) and ``dtw_average`` (:421-690).  Same signatures, same results.

B200-first differences: the alignment matrix is never shipped around -- the device
records the back-walk's moves and returns the path as one column range per row
(``wb_cuda_dtw_paths``); ``dtw_average(method="mm")`` runs every epoch as ONE library call over all
samples (``wb_cuda_dba_epoch``) against a sample set that stays resident on the device, instead of a
Python loop over samples and path cells.  ``dtw_paths`` / ``dtw_average_many`` are the batched forms
the estimators use.  There is no CPU fallback.
"""
import math
import numbers

import numpy as np

from . import _shim
from .distance import DtwMetric, WeightedDtwMetric, _check_scalar, check_array, pairwise_distance

__all__ = [
    "dtw_alignment", "wdtw_alignment", "dtw_distance", "wdtw_distance", "ddtw_distance", "wddtw_distance",
    "dtw_mapping", "jeong_weight", "dtw_average", "dtw_paths", "dtw_average_many", "dtw_envelop", "dtw_lb_keogh",
]


def _compute_warp_size(x_size, r, *, y_size=0):
    """dtw.py:38-40."""
    _check_scalar(r, "r", min_val=0, max_val=1)
    return max(math.floor(max(x_size, y_size) * r), 1)


def _series(x, name):
    x = check_array(x, ensure_2d=False, dtype=float, input_name=name)
    return x.ravel() if x.ndim != 1 else x


def dtw_distance(x, y, *, r=1.0):
    """dtw.py:43-67."""
    return pairwise_distance(_series(x, "x"), _series(y, "y"), metric="dtw", metric_params={"r": r})


def ddtw_distance(x, y, *, r=1.0):
    """dtw.py:70-94."""
    return pairwise_distance(_series(x, "x"), _series(y, "y"), metric="ddtw", metric_params={"r": r})


def wdtw_distance(x, y, *, r=1.0, g=0.05):
    """dtw.py:97-123."""
    return pairwise_distance(_series(x, "x"), _series(y, "y"), metric="wdtw", metric_params={"r": r, "g": g})


def wddtw_distance(x, y, *, r=1.0, g=0.05):
    """dtw.py:126-152."""
    return pairwise_distance(_series(x, "x"), _series(y, "y"), metric="wddtw", metric_params={"r": r, "g": g})


def dtw_envelop(x, *, r=1.0):
    """Envelope for LB_Keogh: ``lower[k], upper[k]`` = min / max of ``x[k-w .. k+w]`` (distance/dtw.py:155-190;
    ``w = max(floor(T r), 1)``, ``T - 1`` when that equals ``T``; ``_dtw_envelop`` EL:1076-1092 on the device kernel that
    also feeds the LB transformers)."""
    x = _series(x, "x")
    warp_size = _compute_warp_size(x.shape[0], r)
    if warp_size == x.shape[0]:
        warp_size -= 1
    if not 0 <= warp_size < x.shape[0]:
        raise ValueError("invalid r")
    lower, upper = _shim.dtw_envelope(x.reshape(1, -1), warp_size)
    return lower[0], upper[0]


def dtw_lb_keogh(x, y=None, *, lower=None, upper=None, r=1.0):
    """LB_Keogh of x against the envelope of y (or a given envelope): ``(min_dist, per-time-step terms)``
    (distance/dtw.py:193-243 -> ``_dtw_lb_keogh`` EL:1095-1115: squared excess over the envelope per step, summed in
    time order, square root of the sum).  If y is given, lower and upper are ignored; otherwise both are required and r
    is ignored."""
    x = _series(x, "x")
    if y is not None:
        y = _series(y, "y")
        if y.shape[0] != x.shape[0]:
            raise ValueError("x (%d) and y (%d) must have the same number of timesteps" % (x.shape[0], y.shape[0]))
        lower, upper = dtw_envelop(y, r=r)
    elif lower is None or upper is None:
        raise ValueError("both y, lower and upper can't be None")
    lower = _series(lower, "lower")
    upper = _series(upper, "upper")
    if lower.shape[0] != upper.shape[0] or lower.shape[0] != x.shape[0]:
        raise ValueError("lower (%d), upper (%d) and x (%d) have the same number of timesteps"
                         % (lower.shape[0], upper.shape[0], x.shape[0]))
    md, cb = _shim.dtw_lb_keogh_terms(x.reshape(1, -1), lower.reshape(1, -1), upper.reshape(1, -1))
    return float(md[0]), cb[0]


def jeong_weight(n, g=0.05):
    """Weights of Jeong et al. (2011), dtw.py:347-374 (numpy's exp, exactly as the reference)."""
    return 1.0 / (1.0 + np.exp(-g * (np.arange(n, dtype=float) - n / 2.0)))


def _check_weight(weight, n):
    if weight is None:
        return None
    weight = _series(weight, "weight")
    if weight.shape[0] != n:
        raise ValueError("weight must have the same size as max(x.size, y.size) %d, got %d" % (n, weight.shape[0]))
    return weight


def dtw_alignment(x, y, *, r=1.0, weight=None, out=None):
    """DTW alignment matrix (dtw.py:246-290, `_dtw_alignment` _elastic.pyx:1011-1073).

    The in-band cells equal the reference's bit for bit; cells outside the Sakoe-Chiba band, which the
    reference leaves uninitialised (``np.empty``), are ``+inf`` here.
    """
    x, y = _series(x, "x"), _series(y, "y")
    _compute_warp_size(x.shape[0], r, y_size=y.shape[0])  # validates r
    weight = _check_weight(weight, max(x.shape[0], y.shape[0]))
    if out is not None and (out.shape[0] < x.shape[0] or out.shape[1] < y.shape[0]):
        raise ValueError("out has wrong shape, got [%d, %d]" % out.shape)
    _, _, mat = _shim.dtw_paths(x.reshape(1, -1), y.reshape(1, -1), r, weights=weight, want_matrix=True)
    if out is not None:
        out[: x.shape[0], : y.shape[0]] = mat[0]
        return out
    return mat[0]


def wdtw_alignment(x, y, *, r=1.0, g=0.5, out=None):
    """dtw.py:293-344."""
    x, y = _series(x, "x"), _series(y, "y")
    return dtw_alignment(x, y, r=r, weight=jeong_weight(max(x.shape[0], y.shape[0]), g), out=out)


def _indicator(lo, hi, n_cols):
    cols = np.arange(n_cols)[None, :]
    return (cols >= lo[:, None]) & (cols <= hi[:, None])


def dtw_mapping(x=None, y=None, *, alignment=None, r=1, return_index=False):
    """Optimal warping path (dtw.py:377-418).

    With ``x`` and ``y`` the path comes straight from the device (no matrix is transferred); a
    precomputed ``alignment`` is walked back on the host exactly like the reference does.
    """
    if alignment is None:
        if x is None or y is None:
            raise ValueError("if alignment=None, neither x or y can be None")
        x, y = _series(x, "x"), _series(y, "y")
        lo, hi = _shim.dtw_paths(x.reshape(1, -1), y.reshape(1, -1), r)
        indicator = _indicator(lo[0], hi[0], y.shape[0])
    else:
        alignment = np.asarray(alignment, dtype=float)
        if alignment.ndim != 2:
            raise ValueError("Expected 2D array, got %dD array instead" % alignment.ndim)
        indicator = np.zeros(alignment.shape, dtype=bool)
        i, j = alignment.shape[0] - 1, alignment.shape[1] - 1
        while i > 0 or j > 0:  # host bookkeeping over a matrix the caller already holds (no DP)
            indicator[i, j] = True
            option_diag = alignment[i - 1, j - 1] if i > 0 and j > 0 else np.inf
            option_up = alignment[i - 1, j] if i > 0 else np.inf
            option_left = alignment[i, j - 1] if j > 0 else np.inf
            move = np.argmin([option_diag, option_up, option_left])
            if move == 0:
                i -= 1
                j -= 1
            elif move == 1:
                i -= 1
            else:
                j -= 1
        indicator[0, 0] = True
    if return_index:
        return indicator, indicator.nonzero()
    return indicator


def dtw_paths(a, b, *, r=1.0, weight=None, ia=None, ib=None, return_cost=False):
    """Batched warping paths (extension): pair p aligns ``a[ia[p]]`` (rows) with ``b[ib[p]]`` (columns).

    Returns ``(lo, hi)`` -- int32 arrays of shape (n_pairs, a_timestep): the path covers columns
    ``lo[p, m] .. hi[p, m]`` of row m -- and, with ``return_cost``, the squared-cost DTW value of each pair.
    """
    a = check_array(a, dtype=float, input_name="a")
    b = check_array(b, dtype=float, input_name="b")
    _compute_warp_size(a.shape[1], r, y_size=b.shape[1])
    weight = _check_weight(weight, max(a.shape[1], b.shape[1]))
    if ia is None and ib is None and a.shape[0] != b.shape[0]:
        raise ValueError("a and b must have the same number of samples when no index arrays are given")
    return _shim.dtw_paths(a, b, r, weights=weight, ia=ia, ib=ib, want_cost=return_cost)


# ---------------------------------------------------------------------------------------------
# DBA
# ---------------------------------------------------------------------------------------------
def _metric_for(r, g):
    return DtwMetric(r=r) if g is None else WeightedDtwMetric(r=r, g=g)


def _mm_many(fitted, n_timestep, means, groups, *, r, g, sample_weight, max_epoch, tol, verbose=False):
    """`_mm_dtw_average` (dtw.py:655-690) for several (mean, group) problems at once.

    groups[c]: ascending sample indices of problem c.  Every epoch is one device step over all still
    active problems; the per-problem cost (np.mean / np.average of the member distances, dtw.py:597-603)
    and the convergence test stay on the host, in numpy, exactly as in the reference.
    """
    metric = _metric_for(r, g)
    means = [np.array(m, dtype=float, copy=True) for m in means]
    K = len(means)
    costs = [None] * K

    def step(active, update):
        offsets = np.zeros(len(active) + 1, dtype=np.int64)
        for k, c in enumerate(active):
            offsets[k + 1] = offsets[k] + len(groups[c])
        members = np.concatenate([np.asarray(groups[c], dtype=np.int64) for c in active])
        tm = means[active[0]].shape[0]
        weights = None if g is None else jeong_weight(max(tm, n_timestep), g)
        new_means, dist = _shim.dba_epoch(fitted, metric.metric_id, metric._params(), np.stack([means[c] for c in active]),
                                          offsets, members, sample_weight=sample_weight, weights=weights, update=update)
        out = []
        for k, c in enumerate(active):
            d = dist[offsets[k]:offsets[k + 1]]
            if sample_weight is None:
                cost = np.mean(d)
            else:
                cost = np.average(d, weights=np.asarray(sample_weight)[np.asarray(groups[c])])
            out.append((new_means[k], cost))
        return out

    active = list(range(K))
    for c, (_, cost) in zip(active, step(active, False)):
        costs[c] = cost
    for epoch in range(max_epoch):
        if not active:
            break
        still = []
        for c, (mean, cost) in zip(active, step(active, True)):
            means[c] = mean
            prev_cost, costs[c] = costs[c], cost
            if abs(prev_cost - cost) < tol:
                if verbose:
                    print(f"Complete at epoch={epoch} with cost={cost}.")
            else:
                still.append(c)
        active = still
    return means, costs


def dtw_average_many(X, groups, inits, *, r=1.0, g=None, sample_weight=None, tol=1e-5, max_epoch=50, fitted=None):
    """DBA (method="mm") of several groups of ``X`` at once (extension used by KMeans).

    ``groups[c]``: ascending sample indices, ``inits[c]``: initial barycentre.  Returns ``(means, costs)``,
    each entry equal to ``dtw_average(X[groups[c]], init=inits[c], method="mm", return_cost=True)``.
    """
    X = check_array(X, dtype=float, input_name="X")
    _check_scalar(r, "r", min_val=0.0, max_val=1.0)
    if g is not None:
        g = _check_scalar(g, "g", min_val=0, include_min=False)
    own = fitted is None
    if own:
        fitted = _shim.FittedSet(X.reshape(X.shape[0], 1, X.shape[1]), devices=[_shim._first_device()])
    try:
        return _mm_many(fitted, X.shape[1], inits, groups, r=r, g=g, sample_weight=sample_weight, max_epoch=max_epoch, tol=tol)
    finally:
        if own:
            fitted.close()


def _check_random_state(seed):
    if seed is None or seed is np.random:
        return np.random.mtrand._rand
    if isinstance(seed, numbers.Integral):
        return np.random.RandomState(seed)
    if isinstance(seed, np.random.RandomState):
        return seed
    raise ValueError("%r cannot be used to seed a numpy.random.RandomState instance" % seed)


def dtw_average(X, *, r=1.0, g=None, sample_weight=None, init="random", method="mm", max_stable=5, learning_rate=0.1,
                decay=0.9, tol=1e-5, max_epoch=50, return_cost=False, verbose=False, random_state=None):
    """DTW barycentre average (dtw.py:421-650); same parameters, same result as the reference."""
    X = check_array(X, dtype=float, input_name="X")
    if X.shape[0] < 2:
        raise ValueError("Found array with %d sample(s) (shape=%s) while a minimum of 2 is required." % (X.shape[0], X.shape))
    r = _check_scalar(r, "r", min_val=0.0, max_val=1.0)
    random_state = _check_random_state(random_state)
    if isinstance(init, str) and init == "random":
        mean = X[random_state.randint(X.shape[0])].copy()
    elif hasattr(init, "__len__") or hasattr(init, "shape") or hasattr(init, "__array__"):
        mean = np.array(init, dtype=float, copy=True)
    else:
        raise ValueError("init must be array-like or 'random', not %r" % type(init).__qualname__)
    if sample_weight is not None:
        sample_weight = np.asarray(sample_weight, dtype=float)
        if sample_weight.shape != (X.shape[0],):
            raise ValueError("sample_weight.shape == {}, expected {}!".format(sample_weight.shape, (X.shape[0],)))
    if g is not None:
        g = _check_scalar(g, "g", min_val=0, include_min=False)

    fitted = _shim.FittedSet(X.reshape(X.shape[0], 1, X.shape[1]), devices=[_shim._first_device()])
    try:
        if method == "mm":
            means, costs = _mm_many(fitted, X.shape[1], [mean], [np.arange(X.shape[0])], r=r, g=g,
                                    sample_weight=sample_weight, max_epoch=max_epoch, tol=tol, verbose=verbose)
            mean, cost = means[0], costs[0]
        elif method == "ssg":
            mean, cost = _ssg(fitted, X, mean, r=r, g=g, sample_weight=sample_weight, max_epoch=max_epoch,
                              max_stable=max_stable, learning_rate=learning_rate, decay=decay, verbose=verbose,
                              random_state=random_state)
        else:
            raise ValueError("method must be 'mm' or 'ssg', got %r" % method)
    finally:
        fitted.close()
    return (mean, cost) if return_cost else mean


def _ssg(fitted, X, mean, *, r, g, sample_weight, max_epoch, learning_rate, decay, max_stable, verbose, random_state):
    """Stochastic subgradient mean (dtw.py:610-652).  The mean changes after every sample, so the alignments
    are inherently sequential: one single-pair path call per sample; the epoch cost is one device step."""
    metric = _metric_for(r, g)
    n = X.shape[0]
    weights = None if g is None else jeong_weight(max(mean.shape[0], X.shape[1]), g)
    all_members = np.arange(n, dtype=np.int64)
    offsets = np.array([0, n], dtype=np.int64)

    def costfn(mean):
        _, dist = _shim.dba_epoch(fitted, metric.metric_id, metric._params(), mean.reshape(1, -1), offsets, all_members, update=False)
        return np.mean(dist) if sample_weight is None else np.average(dist, weights=sample_weight)

    best_mean, min_cost, n_stable = None, np.inf, 0
    order = np.arange(n)
    z = np.empty(mean.shape[0], dtype=float)
    for epoch in range(max_epoch):
        if n_stable > max_stable:
            if verbose:
                print(f"Completed at epoch={epoch} with cost={min_cost}.")
            break
        random_state.shuffle(order)
        for i, o in enumerate(order):
            z.fill(0)
            lo, hi = _shim.dtw_paths(mean.reshape(1, -1), X[o:o + 1], r, weights=weights)
            w = 1.0 if sample_weight is None else sample_weight[o]
            for m in range(mean.shape[0]):  # z[m] += mean[m] - X[o, x] * w over the path cells of row m, ascending
                for x in range(lo[0, m], hi[0, m] + 1):
                    z[m] += mean[m] - X[o, x] * w
            mean -= learning_rate * z
            if epoch == 0:
                learning_rate = decay ** i * learning_rate
        cost = costfn(mean)
        if cost < min_cost:
            if verbose:
                print(f"New min cost={cost} at epoch={epoch} with learning_rate={learning_rate}.")
            min_cost, n_stable, best_mean = cost, 0, mean.copy()
        else:
            n_stable += 1
    return best_mean, min_cost
