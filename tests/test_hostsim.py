"""CPU: the DEVICE engines (wildboar_b200/csrc/engine_*.cuh) compiled for the host and run one
emulated thread at a time must be bit-equal to the oracle for every metric and geometry."""
import numpy as np
import pytest

import sim
from util import METRICS


def _params(oracle, metric, **kw):
    op = oracle.make_params(metric, **kw)
    return sim.WbParams(op.r, op.g, op.p, op.c, op.epsilon, op.penalty, op.stiffness, 0, 0)


@pytest.mark.parametrize("metric", METRICS)
def test_engines_match_oracle(oracle, metric):
    rng = np.random.default_rng(hash(metric) % 1000)
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
