"""CPU: the DEVICE engines (wildboar_b200/csrc/engine_*.cuh) compiled for the host and run one
emulated thread at a time must be bit-equal to the oracle for every metric and geometry."""
import numpy as np
import pytest

import sim
from util import METRICS


def _stable_hash(name):
    """Python's hash() of a str changes from process to process (PYTHONHASHSEED): the seeds of these sweeps must not."""
    import zlib
    return zlib.crc32(name.encode())


def _params(oracle, metric, **kw):
    op = oracle.make_params(metric, **kw)
    return sim.WbParams(op.r, op.g, op.p, op.c, op.epsilon, op.penalty, op.stiffness, 0, 0)


@pytest.mark.parametrize("metric", METRICS)
def test_engines_match_oracle(oracle, metric):
    rng = np.random.default_rng(_stable_hash(metric) % 1000)
    mid = oracle.METRIC_IDS[metric]
    checked = 0
    for trial in range(40):
        mode = trial % 5
        if mode == 0:
            Tx = Ty = int(rng.integers(2, 120))
        elif mode == 1:
            Tx, Ty = int(rng.integers(2, 20)), int(rng.integers(40, 120))
        elif mode == 2:
            Tx, Ty = int(rng.integers(40, 120)), int(rng.integers(2, 20))
        elif mode == 3:
            Tx = int(rng.integers(30, 100)); Ty = Tx + int(rng.integers(-4, 5))
        else:
            Tx, Ty = int(rng.integers(1, 10)), int(rng.integers(1, 10))
        if metric == "wddtw" and Tx > Ty:
            Tx, Ty = Ty, Tx
        r = float(rng.choice([0, 0.01, 0.05, 0.1, 0.2, 0.5, 1.0]))
        x = np.cumsum(rng.standard_normal(Tx)); y = np.cumsum(rng.standard_normal(Ty))
        ref = oracle.pairwise(metric, x, y.reshape(1, -1), r=r)[0, 0]
        p = _params(oracle, metric, r=r)
        for engine, W in [(1, 0), (2, 2), (2, 4), (2, 8), (2, 12), (2, 16)]:
            rc, v, _ = sim.pair(engine, W, mid, p, x, y, bs=int(rng.integers(1, 3)))
            if rc == 1:
                continue  # geometry routed to the row-scan engine
            assert rc == 0 and v == ref, (metric, engine, W, Tx, Ty, r, v, ref)
            checked += 1
    assert checked > 100


def test_rowscan_early_abandon_matches_reference_semantics(oracle):
    """Row minima / abandoning of the row-scan engine reproduce eadistance (CD:1322-1340)."""
    rng = np.random.default_rng(9)
    for metric in ("dtw", "erp", "msm", "twe", "lcss", "edr"):
        mid = oracle.METRIC_IDS[metric]
        x = np.cumsum(rng.standard_normal((6, 30)), axis=1)
        y = np.cumsum(rng.standard_normal((40, 30)), axis=1)
        p = _params(oracle, metric, r=0.2)
        k = 3
        ref_idx, ref_dist = oracle.argmin(metric, x, y, k=k, r=0.2)
        for i in range(len(x)):
            # replay the sequential scan from (d, M) pairs exactly as argmin.cuh does
            idxs, vals, t = [], [], np.inf
            for j in range(len(y)):
                rc, d, M = sim.pair(1, 0, mid, p, x[i], y[j], ea=1)
                assert rc == 0
                if metric == "dtw":
                    T = t * t
                elif metric == "lcss":
                    T = np.inf if np.isinf(t) else 30 - t * 30
                elif metric == "edr":
                    T = t * 30
                else:
                    T = t
                if metric != "dtw" and M > T:
                    continue
                if d < t:
                    idxs.append(j); vals.append(d)
                    hi, hv, n = oracle.heap_replay(k, idxs, vals)
                    t = hv[0] if n == k else np.inf
            hi, hv, n = oracle.heap_replay(k, idxs, vals)
            assert np.array_equal(hi, ref_idx[i]) and np.array_equal(hv, ref_dist[i]), (metric, i)


def test_strip_abandon_is_safe(oracle):
    """Column-minimum abandoning of the strip engine never drops a pair below the threshold."""
    rng = np.random.default_rng(3)
    mid = oracle.METRIC_IDS["dtw"]
    p = _params(oracle, "dtw", r=0.1)
    for _ in range(200):
        x = np.cumsum(rng.standard_normal(64)); y = np.cumsum(rng.standard_normal(64))
        d = oracle.pairwise("dtw", x, y.reshape(1, -1), r=0.1)[0, 0]
        thr = d * float(rng.uniform(0.5, 1.5))
        rc, v, _ = sim.pair(2, 8, mid, p, x, y, min_dist_raw=thr * thr)
        assert rc == 0
        if d < thr:
            assert v == d
        else:
            assert v == d or np.isinf(v)


def _replay_first(d, M, kind, scale):
    """k = 1 replay of k_replay (argmin.cuh): accept iff d < t and not M > T(t); returns (t, index)."""
    t, idx = np.inf, 0
    for w in range(len(d)):
        if not d[w] < t:
            continue
        if M is not None:
            if kind == "ident":
                T = t
            elif kind == "lcss":
                T = np.inf if np.isinf(t) else scale - t * scale
            elif kind == "scale":
                T = t * scale
            else:
                T = t * t
            if M[w] > T:
                continue
        t, idx = d[w], w
    return t, idx


_SCAN_CASES = [("lcss", {"r": 0.2, "epsilon": 0.6}), ("lcss", {}), ("erp", {"r": 0.1, "g": 0.3}), ("erp", {}),
               ("edr", {"r": 0.25, "epsilon": 0.5}), ("edr", {}), ("msm", {"r": 0.15, "c": 0.4}), ("msm", {}),
               ("twe", {"r": 0.2, "penalty": 0.4, "stiffness": 0.05}), ("twe", {})]


@pytest.mark.parametrize("metric,mp", _SCAN_CASES)
def test_subsequence_scan_scheme_matches_oracle(oracle, metric, mp):
    """The device scheme of subseq_scan_worker (wb_cuda.cu) -- every window's (d, M) from the row-scan engine without
    abandoning, then the k = 1 replay with the metric's threshold transform -- reproduces the reference's
    early-abandoning scan for lcss / erp / edr / msm / twe (the oracle is pinned to the reference)."""
    rng = np.random.default_rng(11)
    mid = oracle.METRIC_IDS[metric]
    T = 48
    X = np.cumsum(rng.standard_normal((4, T)), axis=1)
    subs = [np.cumsum(rng.standard_normal(m)) for m in (1, 2, 7, 16, 31, 48)]
    od, oi = oracle.pairwise_subsequence(metric, subs, X, **mp)
    for k, s in enumerate(subs):
        m = len(s)
        kw = dict(mp)
        if metric == "edr" and "epsilon" not in kw:
            kw["epsilon"] = oracle._subsequence_mean_std(s)[1] / 4.0
        p = _params(oracle, metric, **kw)
        kind, scale = {"lcss": ("lcss", float(m)), "edr": ("scale", float(T))}.get(metric, ("ident", 1.0))
        for i in range(len(X)):
            d, M = [], []
            for w in range(T - m + 1):
                rc, dv, mv = sim.pair(1, 0, mid, p, s, X[i, w:w + m])
                assert rc == 0
                d.append(dv); M.append(mv)
            t, idx = _replay_first(d, M, kind, scale)
            assert t == od[i, k] and idx == oi[i, k], (metric, m, i, t, od[i, k], idx, oi[i, k])


@pytest.mark.parametrize("metric,mp", _SCAN_CASES + [("adtw", {"r": 0.2, "p": 0.3}), ("ddtw", {"r": 0.3}), ("adtw", {})])
def test_scaled_subsequence_scan_scheme_matches_oracle(oracle, metric, mp):
    """Same for scaled_<metric> (ScaledSubsequenceMetricWrap): windows z-normalised with the device's IncStats loop."""
    rng = np.random.default_rng(12)
    mid = oracle.METRIC_IDS[metric]
    T = 40
    X = np.cumsum(rng.standard_normal((3, T)), axis=1)
    X[2, 10:22] = X[2, 10]  # constant stretch: windows with zero variance (std -> 1)
    subs = [np.cumsum(rng.standard_normal(m)) for m in (1, 3, 9, 20, 40)]
    od, oi = oracle.pairwise_scaled_subsequence(metric, subs, X, **mp)
    p = _params(oracle, metric, **mp)
    dtwfam = metric in ("adtw", "ddtw")
    for k, s in enumerate(subs):
        m = len(s)
        mean, std = oracle._subsequence_mean_std(s)
        sn = (s - mean) / std
        kind, scale = {"lcss": ("lcss", float(m)), "edr": ("scale", float(m))}.get(metric, ("ident", 1.0))
        for i in range(len(X)):
            wm, ws = sim.inc_window_stats(X[i], m)
            om, os_ = oracle.inc_window_stats(X[i], m)
            assert np.array_equal(wm, om) and np.array_equal(ws, os_)
            if metric == "ddtw" and m < 3:
                assert np.isinf(od[i, k])
                continue
            d, M = [], []
            for w in range(T - m + 1):
                rc, dv, mv = sim.pair(1, 0, mid, p, sn, (X[i, w:w + m] - wm[w]) / ws[w], ea=1)
                assert rc == 0
                d.append(dv); M.append(mv)
            t, idx = _replay_first(d, None if dtwfam else M, kind, scale)
            assert t == od[i, k] and idx == oi[i, k], (metric, m, i, t, od[i, k], idx, oi[i, k])


@pytest.mark.parametrize("metric", METRICS)
def test_band_engine_matches_rowscan_and_oracle(oracle, metric):
    """The band-register engine (engine_band.cuh: previous band row in registers) returns the row-scan engine's value AND
    its row-minimum maximum, bit for bit, with and without abandoning, for every metric and every narrow-band geometry."""
    rng = np.random.default_rng(100 + _stable_hash(metric) % 1000)
    mid = oracle.METRIC_IDS[metric]
    checked = abandoned = 0
    for trial in range(120):
        T = int(rng.integers(2, 70)) if trial % 3 else int(rng.integers(2, 12))
        r = float(rng.choice([0, 0.02, 0.05, 0.1, 0.15, 0.2, 0.3, 0.5, 1.0]))
        x = np.cumsum(rng.standard_normal(T)); y = np.cumsum(rng.standard_normal(T))
        if trial % 7 == 0:
            y = x + 0.05 * rng.standard_normal(T)
        p = _params(oracle, metric, r=r)
        for ea in (0, 1):
            rc1, v1, m1 = sim.pair(1, 0, mid, p, x, y, ea=ea)
            assert rc1 == 0
            for HB in (8, 16, 32):
                rc, v, mm = sim.pair(3, HB, mid, p, x, y, ea=ea)
                if rc == 1:
                    continue  # band too wide for HB (or MSM with R = 1): row-scan territory
                assert rc == 0 and v == v1 and mm == m1, (metric, T, r, HB, ea, v, v1, mm, m1)
                checked += 1
                if ea == 0:
                    assert v == oracle.pairwise(metric, x, y.reshape(1, -1), r=r)[0, 0]
                # abandoning against a bound around the row-minimum maximum: same decision, same M at the exit
                for f in (0.5, 0.999999, 1.0, 1.5):
                    thr = m1 * f if np.isfinite(m1) and m1 != 0 else f - 0.75
                    rca, va, ma = sim.pair(1, 0, mid, p, x, y, ea=ea, min_dist_raw=thr)
                    rcb, vb, mb = sim.pair(3, HB, mid, p, x, y, ea=ea, min_dist_raw=thr)
                    assert rcb == 0 and ((va == vb) or (np.isnan(va) and np.isnan(vb))) and ma == mb, (metric, T, r, HB, thr, va, vb, ma, mb)
                    abandoned += np.isinf(vb) and not np.isinf(v1)
    assert checked > 100 and abandoned > 10


def test_interleaved_layout_roundtrip():
    """k_interleave32's index map (incl. the padded last group) and the base + 32 t walk of the YS = 32 kernels."""
    rng = np.random.default_rng(2)
    for n, T in ((1, 1), (31, 5), (32, 5), (33, 5), (40, 7), (64, 3), (1000, 13), (97, 64)):
        assert sim.interleave_roundtrip(rng.standard_normal((n, T))) == 0, (n, T)


@pytest.mark.parametrize("metric,mp", [c for c in _SCAN_CASES if c[0] in ("erp", "msm", "twe", "edr")])
def test_subsequence_scan_with_head_then_abandon_matches_oracle(oracle, metric, mp):
    """The abandoning scheme of subseq_scan_worker: a head of windows evaluated fully and replayed, every later window
    abandoned against T(running minimum of the head) -- through the band engine's own abandon test -- and replayed with the
    carried-over state.  Must equal the reference's scan (whose bound is never larger)."""
    rng = np.random.default_rng(13)
    mid = oracle.METRIC_IDS[metric]
    T, head = 60, 5
    X = np.cumsum(rng.standard_normal((4, T)), axis=1)
    subs = [np.cumsum(rng.standard_normal(m)) for m in (6, 13, 24)] + [X[1, 20:32].copy()]
    od, oi = oracle.pairwise_subsequence(metric, subs, X, **mp)
    abandoned = 0
    for k, s in enumerate(subs):
        m = len(s)
        kw = dict(mp)
        if metric == "edr" and "epsilon" not in kw:
            kw["epsilon"] = oracle._subsequence_mean_std(s)[1] / 4.0
        p = _params(oracle, metric, **kw)
        scale = float(T) if metric == "edr" else 1.0
        for i in range(len(X)):
            nw = T - m + 1
            d, M = [], []
            for w in range(head):
                rc, dv, mv = sim.pair(1, 0, mid, p, s, X[i, w:w + m])
                d.append(dv); M.append(mv)
            t1, _ = _replay_first(d, M, "scale" if metric == "edr" else "ident", scale)
            for w in range(head, nw):
                rc, dv, mv = sim.pair(3, 32, mid, p, s, X[i, w:w + m], min_dist_raw=t1 * scale)
                if rc == 1:  # band too wide for the band engine: the row-scan engine abandons the same way
                    rc, dv, mv = sim.pair(1, 0, mid, p, s, X[i, w:w + m], min_dist_raw=t1 * scale)
                assert rc == 0
                abandoned += np.isinf(dv)
                d.append(dv); M.append(mv)
            t, idx = _replay_first(d, M, "scale" if metric == "edr" else "ident", scale)
            assert t == od[i, k] and idx == oi[i, k], (metric, m, i, t, od[i, k], idx, oi[i, k])
    assert abandoned > 0 or metric == "edr"


@pytest.mark.parametrize("metric", ["msm", "twe", "erp", "lcss", "edr", "adtw", "ddtw"])
def test_interleaved_engine_variants_match_plain(oracle, metric):
    """The YS = 32 instantiations of the band and row-scan engines (y read from an interleaved group, element stride 32)
    return exactly what the plain ones return, value and row-minimum maximum."""
    rng = np.random.default_rng(200 + _stable_hash(metric) % 1000)
    mid = oracle.METRIC_IDS[metric]
    checked = 0
    for trial in range(40):
        T = int(rng.integers(2, 60))
        r = float(rng.choice([0, 0.05, 0.1, 0.2, 0.5, 1.0]))
        x = np.cumsum(rng.standard_normal(T)); y = np.cumsum(rng.standard_normal(T))
        p = _params(oracle, metric, r=r)
        rc1, v1, m1 = sim.pair(1, 0, mid, p, x, y, ea=1)
        rc5, v5, m5 = sim.pair(5, 0, mid, p, x, y, ea=1)
        assert rc1 == 0 and rc5 == 0 and v5 == v1 and m5 == m1, (metric, T, r)
        for HB in (8, 16, 32):
            rc4, v4, m4 = sim.pair(4, HB, mid, p, x, y, ea=1)
            if rc4 == 1:
                continue
            assert rc4 == 0 and v4 == v1 and m4 == m1, (metric, T, r, HB)
            checked += 1
    assert checked > 20


# ---- cooperative engine (engine_coop.cuh): a group of lanes per pair, emulated in lockstep on the host ----
@pytest.mark.parametrize("metric", METRICS)
def test_coop_engine_matches_oracle(oracle, metric):
    """All 11 metrics x equal / unequal lengths x windows x (W, G) layouts, incl. the layouts the library ships (W = 8 and
    W = 13): masked rows (start-up, drain, band outside the matrix, row 0 / column 0 rules, MSM's stale left edge and extra
    cell) and the unrolled fast blocks with rotating column registers must reproduce the oracle bit for bit."""
    rng = np.random.default_rng(1000 + _stable_hash(metric) % 1000)
    mid = oracle.METRIC_IDS[metric]
    checked = 0
    for trial in range(36):
        mode = trial % 6
        if mode == 0:
            Tx = Ty = int(rng.integers(4, 140))
        elif mode == 1:
            Tx = int(rng.integers(30, 100)); Ty = Tx + int(rng.integers(-6, 7))
        elif mode == 2:
            Tx, Ty = int(rng.integers(5, 30)), int(rng.integers(40, 120))
        elif mode == 3:
            Tx, Ty = int(rng.integers(40, 120)), int(rng.integers(5, 30))
        elif mode == 4:
            Tx = Ty = int(rng.integers(300, 700))      # long interior: many fast blocks
        else:
            Tx = Ty = 150                               # cfg1's shape (r = 0.1: H = 29 -> W = 8, G = 4)
        if metric == "wddtw" and Tx > Ty:
            Tx, Ty = Ty, Tx
        r = 0.1 if mode == 5 else float(rng.choice([0.02, 0.05, 0.1, 0.2, 0.3, 0.5, 1.0]))
        if mode == 4:
            r = float(rng.choice([0.02, 0.05, 0.1]))
        x = np.cumsum(rng.standard_normal(Tx)); y = np.cumsum(rng.standard_normal(Ty))
        ref = oracle.pairwise(metric, x, y.reshape(1, -1), r=r)[0, 0]
        p = _params(oracle, metric, r=r)
        for W, G in [(3, 32), (4, 8), (4, 32), (8, 4), (8, 16), (8, 32), (13, 32), (108, 16), (104, 16)]:
            rc, v = sim.coop_pair(W, G, mid, p, x, y)
            if rc == 1:
                continue  # no exact tiling of this band with (W, G)
            assert rc == 0 and v == ref, (metric, W, G, Tx, Ty, r, v, ref)
            checked += 1
    assert checked > 60


def test_coop_engine_cfg5_band_shape(oracle):
    """The long-series configuration the engine exists for: H = 407 band coordinates over 32 lanes (23 wide, 9 narrow)."""
    rng = np.random.default_rng(77)
    T = 4096
    x = np.cumsum(rng.standard_normal(T)); y = np.cumsum(rng.standard_normal(T))
    for metric in ("msm", "twe"):
        ref = oracle.pairwise(metric, x, y.reshape(1, -1), r=0.05)[0, 0]
        rc, v = sim.coop_pair(13, 32, oracle.METRIC_IDS[metric], _params(oracle, metric, r=0.05), x, y)
        assert rc == 0 and v == ref, (metric, v, ref)
