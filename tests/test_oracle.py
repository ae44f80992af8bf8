"""CPU: the oracle (oracle/elastic_oracle.c) against the reference's golden vectors and, when
oracle/_ref is built, against the live reference.  This is what pins the oracle."""
import numpy as np
import pytest

from util import METRICS, golden_cases, random_walks


def test_oracle_matches_golden(oracle, golden):
    n = 0
    for case, metric, extra, r, pre in golden_cases(golden):
        x, y = golden[f"x{case}"], golden[f"y{case}"]
        assert np.array_equal(oracle.pairwise(metric, x, y, r=r, **extra), golden[pre + "|pairwise"]), pre
        if pre + "|self" in golden:
            assert np.array_equal(oracle.pairwise(metric, x, None, r=r, **extra), golden[pre + "|self"]), pre
            m = min(len(x), len(y))
            assert np.array_equal(oracle.paired(metric, x[:m], y[:m], r=r, **extra), golden[pre + "|paired"]), pre
        for k in (1, 2):
            if pre + f"|argmin{k}|idx" in golden:
                idx, dist = oracle.argmin(metric, x, y, k=k, r=r, **extra)
                assert np.array_equal(idx, golden[pre + f"|argmin{k}|idx"]), pre
                assert np.array_equal(dist, golden[pre + f"|argmin{k}|dist"]), pre
        n += 1
    assert n > 500


def test_oracle_matches_live_reference(oracle):
    from oracle import ref
    wd = ref.load()
    if wd is None:
        pytest.skip("oracle/_ref not built (oracle/build_ref.sh needs /root/reference)")
    rng = np.random.default_rng(42)
    for metric in METRICS:
        for (nx, ny, Tx, Ty) in [(6, 7, 40, 40), (4, 5, 33, 47), (5, 4, 47, 33), (3, 3, 5, 5)]:
            if metric == "wddtw" and Tx > Ty:
                continue
            x = np.cumsum(rng.standard_normal((nx, Tx)), axis=1)
            y = np.cumsum(rng.standard_normal((ny, Ty)), axis=1)
            for r in (0.0, 0.1, 0.35, 1.0):
                a = wd.pairwise_distance(x, y, metric=metric, metric_params={"r": r})
                assert np.array_equal(a, oracle.pairwise(metric, x, y, r=r)), (metric, Tx, Ty, r)
                ai, ad = wd.argmin_distance(x, y, k=2, metric=metric, metric_params={"r": r}, return_distance=True)
                bi, bd = oracle.argmin(metric, x, y, k=2, r=r)
                assert np.array_equal(ai, bi) and np.array_equal(ad, bd), (metric, Tx, Ty, r)


def test_oracle_threads_agree(oracle):
    x, y = random_walks(13, 50, 1), random_walks(9, 50, 2)
    for metric in ("dtw", "msm", "edr"):
        a = oracle.pairwise(metric, x, y, r=0.2, n_jobs=1)
        b = oracle.pairwise(metric, x, y, r=0.2, n_jobs=4)
        assert np.array_equal(a, b)
        ai, ad = oracle.argmin(metric, x, y, k=3, r=0.2, n_jobs=1)
        bi, bd = oracle.argmin(metric, x, y, k=3, r=0.2, n_jobs=3)
        assert np.array_equal(ai, bi) and np.array_equal(ad, bd)


def test_lb_keogh_is_a_lower_bound(oracle):
    # property test of the reference (tests/wildboar/distance/test_lb.py:94-105) on synthetic data
    x, y = random_walks(8, 64, 3), random_walks(8, 64, 4)
    for r in (0.05, 0.1, 0.3, 1.0):
        R = oracle.compute_r(64, r)
        d = oracle.pairwise("dtw", x, y, r=r)
        for i in range(8):
            for j in range(8):
                lo, hi = oracle.envelope(y[j], R - 1)
                assert oracle.lb_keogh_one(x[i], lo, hi) <= d[i, j] + 1e-12
