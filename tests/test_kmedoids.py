"""KMedoids mirror (wildboar_b200.neighbors.KMedoids; reference: src/wildboar/distance/_neighbors.py:615-930 and the PAM
helpers of _cneighbors.pyx) against golden vectors generated from the reference (tests/golden/make_golden_kmedoids.py).

CPU: the clustering logic alone on the reference's own distance matrices (metric="precomputed": no kernel involved).
GPU: the whole estimator -- distance matrix from the device self join, then the same bookkeeping."""
import ast
import warnings

import numpy as np
import pytest


@pytest.fixture(scope="module")
def km_golden():
    import os
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "kmedoids_golden.npz")
    with np.load(path) as z:
        return {k: z[k] for k in z.files}


def _cases(g):
    return list(enumerate(ast.literal_eval(str(g["meta_cases"]))))


def _check(est, g, pre):
    assert np.array_equal(est.medoid_indices_, g[pre + "|medoids"]), pre
    assert np.array_equal(est.labels_, g[pre + "|labels"]), pre
    assert est.inertia_ == g[pre + "|inertia"], pre
    assert est.n_iter_ == g[pre + "|n_iter"], pre


def test_kmedoids_bookkeeping_matches_reference_on_precomputed_matrices(wb, km_golden):
    from wildboar_b200.neighbors import KMedoids
    g = km_golden
    for c, (metric, mp, k, alg, init, n_init, seed) in _cases(g):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            est = KMedoids(n_clusters=k, metric="precomputed", algorithm=alg, init=init, n_init=n_init, random_state=seed).fit(g[f"{c}|dist"])
        _check(est, g, f"{c}|pre")
        assert est.cluster_centers_ is None
        assert np.array_equal(est.transform(g[f"{c}|dist"][:5]), g[f"{c}|dist"][:5][:, est.medoid_indices_])


def test_kmedoids_pam_helpers_match_live_reference(wb, oracle):
    from oracle import ref
    if ref.load() is None:
        pytest.skip("oracle/_ref not built")
    from wildboar.distance._cneighbors import _pam_build, _pam_optimal_swap
    from wildboar_b200.neighbors import _pam_build as my_build, _pam_optimal_swap as my_swap
    rng = np.random.default_rng(3)
    for trial in range(6):
        n, k = int(rng.integers(12, 40)), int(rng.integers(2, 6))
        X = np.cumsum(rng.standard_normal((n, 20)), axis=1)
        D = oracle.pairwise("msm" if trial % 2 else "dtw", X, None, r=0.3)   # msm: asymmetric values mirrored, like the reference's matrix
        med = _pam_build(D, k)
        assert np.array_equal(my_build(D, k), med)
        not_med = np.delete(np.arange(n), med)
        djs, ejs = np.sort(D[med], axis=0)[[0, 1]]
        want = _pam_optimal_swap(D, med.astype(np.intp), not_med.astype(np.intp), djs, ejs, k)
        got = my_swap(D, med, not_med, djs, ejs, k)
        assert (want is None and got is None) or (tuple(want)[:2] == got[:2] and want[2] == got[2]), (trial, want, got)


def test_kmedoids_validation(wb):
    from wildboar_b200.neighbors import KMedoids
    x = np.zeros((6, 8))
    for kw in (dict(n_clusters=0), dict(metric="euclidean"), dict(algorithm="slow"), dict(init="best"), dict(n_init=0), dict(max_iter=0),
               dict(tol=-1.0), dict(metric_params=3)):
        with pytest.raises(ValueError):
            KMedoids(**kw).fit(x)
    with pytest.raises(Exception):
        KMedoids().transform(x)


@pytest.mark.gpu
def test_kmedoids_matches_reference_golden(wb, km_golden):
    from wildboar_b200.neighbors import KMedoids
    wb.set_devices([0])
    g = km_golden
    for c, (metric, mp, k, alg, init, n_init, seed) in _cases(g):
        X, Q = g[f"{c}|X"], g[f"{c}|Q"]
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            est = KMedoids(n_clusters=k, metric=metric, metric_params=mp, algorithm=alg, init=init, n_init=n_init, random_state=seed).fit(X)
        _check(est, g, f"{c}|fit")
        assert np.array_equal(est.cluster_centers_, g[f"{c}|fit|centers"])
        assert np.array_equal(est.transform(Q), g[f"{c}|fit|transform"]), (c, metric)
        assert np.array_equal(est.predict(Q), g[f"{c}|fit|predict"])
        assert np.array_equal(KMedoids(n_clusters=k, metric=metric, metric_params=mp, algorithm=alg, init=init, n_init=n_init,
                                       random_state=seed).fit_predict(X), g[f"{c}|fit|labels"])


# ---- proximity-tree pivot loops (wildboar_b200.tree; reference: src/wildboar/tree/_cptree.pyx:273-289, 887-910) ----
@pytest.mark.gpu
@pytest.mark.parametrize("metric,mp", [("dtw", {"r": 0.1}), ("msm", {"r": 0.2}), ("wdtw", {"r": 0.3, "g": 0.1}), ("erp", {"r": 0.2}),
                                        ("twe", {"r": 0.15}), ("lcss", {"r": 0.5, "epsilon": 0.7})])
def test_pivot_partition_keeps_the_operand_order_of_the_loop_it_replaces(wb, oracle, metric, mp):
    from wildboar_b200.tree import find_min_branch, partition_pivots
    wb.set_devices([0])
    rng = np.random.default_rng(17)
    X = np.cumsum(rng.standard_normal((90, 64)), axis=1)
    for n_node, n_branch in ((90, 3), (17, 2), (40, 5)):
        samples = rng.choice(90, n_node, replace=False)
        pivots = rng.choice(samples, n_branch, replace=False)
        # _partition_pivots: distance(metric, X, pivots[p], X, j) -- pivot first
        want_d = oracle.pairwise(metric, X[pivots], X[samples], n_jobs=0, **mp).T
        branch, d = partition_pivots(X, samples, pivots, metric=metric, metric_params=mp, return_distance=True)
        assert np.array_equal(d, want_d) and np.array_equal(branch, want_d.argmin(axis=1)), (metric, n_node)
        # find_min_branch: _distance(metric, sample, pivot) -- sample first
        want_d = oracle.pairwise(metric, X[samples], X[pivots], n_jobs=0, **mp)
        branch, d = find_min_branch(X[pivots], X[samples], metric=metric, metric_params=mp, return_distance=True)
        assert np.array_equal(d, want_d) and np.array_equal(branch, want_d.argmin(axis=1)), (metric, n_node)
    # ties: the first pivot wins (strict `<`)
    Xt = np.vstack([X[:1], X[:1], X[1:5]])
    assert np.array_equal(partition_pivots(Xt, [2, 3, 4], [0, 1], metric=metric, metric_params=mp), np.zeros(3, dtype=np.intp))


# ---- MDS (wildboar_b200.manifold.MDS; reference: src/wildboar/distance/_manifold.py) ----
# golden vectors: reference embeddings for seeded random walks (generated in the build container with oracle/_ref:
# wd.MDS(n_components, metric_mds, n_init=2, max_iter=60, random_state, metric, metric_params).fit_transform(X))
@pytest.mark.gpu
def test_mds_matches_reference_golden(wb):
    import os
    from wildboar_b200.manifold import MDS
    wb.set_devices([0])
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mds_golden.npz")
    with np.load(path) as z:
        g = {k: z[k] for k in z.files}
    for c, (metric, mp, nc, mm, seed) in enumerate(ast.literal_eval(str(g["meta_cases"]))):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            est = MDS(n_components=nc, metric_mds=mm, n_init=2, max_iter=60, random_state=seed, metric=metric, metric_params=mp)
            emb = est.fit_transform(g[f"{c}|X"])
        assert np.array_equal(emb, g[f"{c}|emb"]), (c, metric)
        assert est.mds_.stress_ == g[f"{c}|stress"]


def test_mds_validation(wb):
    from wildboar_b200.manifold import MDS
    for kw in (dict(n_components=0), dict(metric="euclidean"), dict(metric_mds=1), dict(eps=-1.0), dict(n_init=0)):
        with pytest.raises(ValueError):
            MDS(**kw).fit(np.zeros((5, 8)))
