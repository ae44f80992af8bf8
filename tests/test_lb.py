"""DTW lower-bound transformers: oracle pinned on the reference's golden vectors (CPU);
device matrices bit-equal to the oracle / golden vectors and usable as `lower_bound=` (GPU)."""
import numpy as np
import pytest

from util import random_walks


@pytest.fixture(scope="module")
def lb_golden():
    import os
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "lb_golden.npz")
    with np.load(path) as z:
        return {k: z[k] for k in z.files}


def _cases(g):
    for k in g:
        if "|" in k:
            parts = k.split("|")
            yield k, g["q" + parts[0]], g["x" + parts[0]], parts


def test_oracle_lb_matches_golden(oracle, lb_golden):
    n = 0
    for k, q, x, parts in _cases(lb_golden):
        got = oracle.lb_kim(q, x) if parts[1] == "kim" else oracle.lb_keogh(q, x, r=float(parts[2]), kind=parts[3])
        assert np.array_equal(got, lb_golden[k]), k
        n += 1
    assert n >= 90


def test_oracle_lb_matches_live_reference(oracle):
    from oracle import ref
    if ref.load() is None:
        pytest.skip("oracle/_ref not built")
    from wildboar.distance.lb import DtwKeoghLowerBound, DtwKimLowerBound
    q, x = random_walks(6, 50, 31), random_walks(9, 50, 32)
    for r in (0.0, 0.07, 0.3, 1.0):
        for kind in ("both", "left", "right"):
            assert np.array_equal(oracle.lb_keogh(q, x, r=r, kind=kind), DtwKeoghLowerBound(r=r, kind=kind).fit(x).transform(q))
    assert np.array_equal(oracle.lb_kim(q, x), DtwKimLowerBound().fit(x).transform(q))


def test_keogh_lower_bounds_dtw_property(oracle):
    """The reference's own property test (tests/wildboar/distance/test_lb.py:94-105) on synthetic data."""
    X = random_walks(30, 60, 33)
    for r in np.linspace(0, 1, 10, endpoint=True):
        d = oracle.pairwise("dtw", X, None, r=float(r))
        lb = oracle.lb_keogh(X, X, r=float(r))
        assert (lb <= d + 1e-12).all()


def test_lb_host_api_validation(wb):
    from wildboar_b200.lb import DtwKeoghLowerBound, DtwKimLowerBound, NotFittedError
    with pytest.raises(NotFittedError):
        DtwKeoghLowerBound().transform(np.zeros((2, 5)))
    with pytest.raises(NotFittedError):
        DtwKimLowerBound().transform(np.zeros((2, 5)))
    with pytest.raises(ValueError):
        DtwKeoghLowerBound(r=1.5).fit(np.zeros((2, 5)))
    with pytest.raises(ValueError):
        DtwKeoghLowerBound(kind="up").fit(np.zeros((2, 5)))
    est = DtwKeoghLowerBound(r=0.2, kind="left")
    assert est.get_params() == {"r": 0.2, "kind": "left"} and est.set_params(r=0.5).r == 0.5
    # envelope attributes follow the reference's definition (no device needed)
    X = random_walks(4, 20, 34)
    est = DtwKeoghLowerBound(r=0.1).fit(X)
    w = 2
    for k in range(20):
        a, b = max(0, k - w), min(19, k + w)
        assert np.array_equal(est.lower_[:, k], X[:, a:b + 1].min(axis=1)) and np.array_equal(est.upper_[:, k], X[:, a:b + 1].max(axis=1))


@pytest.mark.gpu
def test_device_lb_matches_golden_and_oracle(wb, oracle, lb_golden):
    from wildboar_b200.lb import DtwKeoghLowerBound, DtwKimLowerBound
    wb.set_devices([0])
    for k, q, x, parts in _cases(lb_golden):
        if parts[1] == "kim":
            got = DtwKimLowerBound().fit(x).transform(q)
        else:
            got = DtwKeoghLowerBound(r=float(parts[2]), kind=parts[3]).fit(x).transform(q)
        assert np.array_equal(got, lb_golden[k]), k
    # larger shapes incl. ragged tiles (sample count not a multiple of 256, query count not of 8, T > chunk)
    for (nq, nx, T, r) in [(37, 700, 256, 0.05), (9, 300, 300, 0.1), (70, 33, 140, 1.0)]:
        q, x = random_walks(nq, T, 35), random_walks(nx, T, 36)
        for kind in ("both", "left", "right"):
            assert np.array_equal(DtwKeoghLowerBound(r=r, kind=kind).fit(x).transform(q), oracle.lb_keogh(q, x, r=r, kind=kind)), (nq, nx, T, kind)
        assert np.array_equal(DtwKimLowerBound().fit(x).transform(q), oracle.lb_kim(q, x))


@pytest.mark.gpu
def test_device_lb_feeds_argmin_like_the_reference(wb, oracle):
    """lb.py:341-351 usage: argmin_distance(..., lower_bound=lbkeogh.transform(query)) equals the oracle's scan."""
    from wildboar_b200.lb import DtwKeoghLowerBound
    wb.set_devices([0])
    q, refs = random_walks(40, 96, 37), random_walks(400, 96, 38)
    lb = DtwKeoghLowerBound(r=0.1).fit(refs).transform(q)
    idx, dist = wb.argmin_distance(q, refs, k=3, metric="dtw", metric_params={"r": 0.1}, lower_bound=lb, return_distance=True,
                                   device_lower_bound=False)
    oi, od = oracle.argmin("dtw", q, refs, k=3, lower_bound=lb, r=0.1)
    assert np.array_equal(idx, oi) and np.array_equal(dist, od)
    assert (lb <= oracle.pairwise("dtw", q, refs, r=0.1, n_jobs=0) + 1e-12).all()


# ---- dtw_envelop / dtw_lb_keogh (distance/dtw.py:155-243), second half of SURVEY 8f-2 ----
@pytest.fixture(scope="module")
def dtw_lb_golden():
    import os
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "dtw_lb_golden.npz")
    with np.load(path) as z:
        return {k: z[k] for k in z.files}


def _dtw_lb_cases(g):
    for k in g:
        if k.startswith("lower|"):
            _, T, r = k.split("|")
            yield int(T), float(r), r


def test_oracle_envelope_matches_dtw_envelop_golden(oracle, dtw_lb_golden):
    g = dtw_lb_golden
    n = 0
    for T, r, rs in _dtw_lb_cases(g):
        w = oracle.lb_warp_size(T, r)
        lo, hi = oracle.envelope(g[f"y|{T}"], w)
        assert np.array_equal(lo, g[f"lower|{T}|{rs}"]) and np.array_equal(hi, g[f"upper|{T}|{rs}"]), (T, r)
        # min_dist = sqrt of the sequential sum of the golden per-step terms; the terms are the squared excess
        x = g[f"x|{T}"]
        cb = np.where(x > hi, (x - hi) ** 2, np.where(x < lo, (x - lo) ** 2, 0.0))
        assert np.array_equal(cb, g[f"cb|{T}|{rs}"]), (T, r)
        s = 0.0
        for v in cb:
            s += v
        assert np.sqrt(s) == g[f"min_dist|{T}|{rs}"], (T, r)
        n += 1
    assert n == 35


def test_dtw_lb_host_validation(wb):
    from wildboar_b200.dtw import dtw_envelop, dtw_lb_keogh
    with pytest.raises(ValueError):
        dtw_envelop(np.arange(5.0), r=1.5)
    with pytest.raises(ValueError, match="can't be None"):
        dtw_lb_keogh(np.arange(5.0))
    with pytest.raises(ValueError, match="same number of timesteps"):
        dtw_lb_keogh(np.arange(5.0), np.arange(6.0))
    with pytest.raises(ValueError, match="same number of timesteps"):
        dtw_lb_keogh(np.arange(5.0), lower=np.arange(4.0), upper=np.arange(5.0))


@pytest.mark.gpu
def test_device_dtw_envelop_and_lb_keogh_match_golden(wb, oracle, dtw_lb_golden):
    from wildboar_b200 import _shim
    from wildboar_b200.dtw import dtw_envelop, dtw_lb_keogh
    wb.set_devices([0])
    g = dtw_lb_golden
    for T, r, rs in _dtw_lb_cases(g):
        x, y = g[f"x|{T}"], g[f"y|{T}"]
        lo, hi = dtw_envelop(y, r=r)
        assert np.array_equal(lo, g[f"lower|{T}|{rs}"]) and np.array_equal(hi, g[f"upper|{T}|{rs}"]), (T, r)
        md, cb = dtw_lb_keogh(x, y, r=r)
        assert md == g[f"min_dist|{T}|{rs}"] and np.array_equal(cb, g[f"cb|{T}|{rs}"]), (T, r)
        md2, cb2 = dtw_lb_keogh(x, lower=lo, upper=hi)
        assert md2 == md and np.array_equal(cb2, cb)
    # batched form behind the mirrors: many series in one call, against the oracle's envelope
    X = random_walks(300, 190, 39)
    lo, hi = _shim.dtw_envelope(X, 7)
    for i in (0, 17, 299):
        ol, oh = oracle.envelope(X[i], 7)
        assert np.array_equal(lo[i], ol) and np.array_equal(hi[i], oh)
    Q = random_walks(300, 190, 40)
    md, cb = _shim.dtw_lb_keogh_terms(Q, lo, hi)
    for i in (0, 17, 299):
        assert md[i] == oracle.lb_keogh_one(Q[i], lo[i], hi[i])
