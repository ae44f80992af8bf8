"""GPU (-m gpu): the CUDA path, called through the public API -> ctypes shim -> C ABI, against
the oracle on identical seeded inputs.  Bar: bit-exact (np.array_equal) in fp64, which implies
the north-star's 1e-12 relative tolerance; argmin indices exact."""
import os

import numpy as np
import pytest

from util import METRICS, NINE, golden_cases, random_walks

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def W(wb):
    assert wb.device_count() >= 1, "no CUDA device: the product has no CPU fallback"
    wb.set_devices([0])
    return wb


def _eq(a, b, what):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    if not np.array_equal(a, b):
        bad = np.argwhere(a != b)
        raise AssertionError(f"{what}: {len(bad)} mismatches, first {bad[0]}: got {a[tuple(bad[0])]!r} want {b[tuple(bad[0])]!r}")


def test_golden_vectors(W, golden):
    """Every committed golden vector of the reference (tests/golden/make_golden.py)."""
    n = 0
    for case, metric, extra, r, pre in golden_cases(golden):
        x, y = golden[f"x{case}"], golden[f"y{case}"]
        mp = dict(extra, r=r)
        _eq(W.pairwise_distance(x, y, metric=metric, metric_params=mp), golden[pre + "|pairwise"], pre + " pairwise")
        if pre + "|self" in golden:
            _eq(W.pairwise_distance(x, metric=metric, metric_params=mp), golden[pre + "|self"], pre + " self")
            m = min(len(x), len(y))
            _eq(W.paired_distance(x[:m], y[:m], metric=metric, metric_params=mp), golden[pre + "|paired"], pre + " paired")
        for k in (1, 2):
            if pre + f"|argmin{k}|idx" in golden:
                idx, dist = W.argmin_distance(x, y, k=k, metric=metric, metric_params=mp, return_distance=True)
                _eq(idx, golden[pre + f"|argmin{k}|idx"], pre + f" argmin{k} idx")
                _eq(dist, golden[pre + f"|argmin{k}|dist"], pre + f" argmin{k} dist")
        n += 1
    assert n > 500


def test_cfg1_gunpoint_shape_dtw(W, oracle):
    """BASELINE configs[0]: 200x150 vs 200x150, dtw r=0.1, full size."""
    x, y = random_walks(200, 150, 1), random_walks(200, 150, 2)
    _eq(W.pairwise_distance(x, y, metric="dtw", metric_params={"r": 0.1}), oracle.pairwise("dtw", x, y, r=0.1, n_jobs=0), "cfg1")


@pytest.mark.parametrize("metric", NINE + ["wddtw", "wlcss"])
def test_cfg2_ecg5000_shape_all_metrics(W, oracle, metric):
    """BASELINE configs[1] (5000x140, default params) on a 160-row slice: two-array and singleton forms."""
    X = random_walks(5000, 140, 1)[:160]
    Y = X[40:].copy()
    _eq(W.pairwise_distance(X[:96], Y, metric=metric), oracle.pairwise(metric, X[:96], Y, n_jobs=0), metric + " two-array")
    _eq(W.pairwise_distance(X, metric=metric), oracle.pairwise(metric, X, None, n_jobs=0), metric + " singleton")
    _eq(W.paired_distance(X[:100], X[60:160], metric=metric), oracle.paired(metric, X[:100], X[60:160], n_jobs=0), metric + " paired")


@pytest.mark.parametrize("metric", METRICS)
@pytest.mark.parametrize("r", [0.0, 0.05, 0.3])
def test_windows_and_unequal_lengths(W, oracle, metric, r):
    for (nx, ny, Tx, Ty) in [(33, 70, 64, 64), (20, 45, 50, 77), (45, 20, 77, 50), (5, 40, 9, 120)]:
        if metric == "wddtw" and Tx > Ty:
            continue
        x, y = random_walks(nx, Tx, 5), random_walks(ny, Ty, 6)
        _eq(W.pairwise_distance(x, y, metric=metric, metric_params={"r": r}), oracle.pairwise(metric, x, y, r=r, n_jobs=0),
            f"{metric} r={r} {Tx}x{Ty}")


def test_cfg3_row_subset(W, oracle):
    """BASELINE configs[2] shape (T=512, r=0.1): 48 x rows against 300 y rows."""
    x, y = random_walks(10000, 512, 1)[:48], random_walks(10000, 512, 2)[:300]
    _eq(W.pairwise_distance(x, y, metric="dtw", metric_params={"r": 0.1}), oracle.pairwise("dtw", x, y, r=0.1, n_jobs=0), "cfg3")


@pytest.mark.parametrize("metric", ["msm", "twe"])
def test_cfg5_long_series_subset(W, oracle, metric):
    """BASELINE configs[4] shape (T=4096, r=0.05): 6 x 40 pairs."""
    x, y = random_walks(2000, 4096, 1)[:6], random_walks(2000, 4096, 2)[:40]
    _eq(W.pairwise_distance(x, y, metric=metric, metric_params={"r": 0.05}), oracle.pairwise(metric, x, y, r=0.05, n_jobs=0), metric)


@pytest.mark.parametrize("metric", ["dtw", "adtw", "msm", "erp", "twe", "lcss"])
def test_engines_agree_on_device(W, oracle, metric, monkeypatch):
    """Strip engine and row-scan engine are independent formulations: both must equal the oracle."""
    x, y = random_walks(40, 90, 7), random_walks(70, 90, 8)
    want = oracle.pairwise(metric, x, y, r=0.15, n_jobs=0)
    for eng in ("rowscan", "strip"):
        monkeypatch.setenv("WILDBOAR_CUDA_ENGINE", eng)
        _eq(W.pairwise_distance(x, y, metric=metric, metric_params={"r": 0.15}), want, f"{metric} {eng}")
        from wildboar_b200 import last_stats
        assert last_stats()["engine"] == {"rowscan": 1, "strip": 2}[eng]


@pytest.mark.parametrize("metric", METRICS)
def test_band_engine_matches_oracle(W, oracle, metric, monkeypatch):
    """The band-register engine (narrow bands, equal lengths) forced for pairwise / paired, and selected automatically behind
    argmin (row minima): H = 7 (HB 8), H = 15 (HB 16), H = 27 (HB 32); all engines agree with the oracle."""
    from wildboar_b200 import last_stats
    for T, r in ((40, 0.1), (80, 0.1), (140, 0.1)):
        x, y = random_walks(33, T, 21), random_walks(70, T, 22)
        want = oracle.pairwise(metric, x, y, r=r, n_jobs=0)
        for eng in ("band", "rowscan"):
            monkeypatch.setenv("WILDBOAR_CUDA_ENGINE", eng)
            _eq(W.pairwise_distance(x, y, metric=metric, metric_params={"r": r}), want, f"{metric} {eng} T={T}")
            assert last_stats()["engine"] == {"band": 3, "rowscan": 1}[eng]
        monkeypatch.delenv("WILDBOAR_CUDA_ENGINE")
        oi, od = oracle.argmin(metric, x, y, k=3, r=r)
        gi, gd = W.argmin_distance(x, y, k=3, metric=metric, metric_params={"r": r}, return_distance=True)
        _eq(gi, oi, f"{metric} argmin idx T={T}")
        _eq(gd, od, f"{metric} argmin dist T={T}")
        if metric in ("lcss", "erp", "edr", "msm", "twe", "wlcss"):
            assert last_stats()["engine"] == 3


def test_edge_shapes(W, oracle):
    for metric in ("dtw", "msm", "edr", "lcss", "twe", "erp"):
        for (nx, ny, Tx, Ty) in [(1, 1, 1, 1), (3, 2, 1, 5), (2, 3, 5, 1), (1, 65, 2, 2), (31, 1, 3, 3), (2, 33, 17, 17)]:
            x, y = random_walks(nx, Tx, 11), random_walks(ny, Ty, 12)
            _eq(W.pairwise_distance(x, y, metric=metric, metric_params={"r": 0.2}), oracle.pairwise(metric, x, y, r=0.2),
                f"{metric} {nx}x{ny} T={Tx},{Ty}")
    x = random_walks(4, 20, 13)
    # 1-D operands and return shapes (DI:514-540)
    d = W.pairwise_distance(x[0], x[1], metric="dtw")
    assert isinstance(d, float) and d == oracle.pairwise("dtw", x[:1], x[1:2])[0, 0]
    assert W.pairwise_distance(x[0], x, metric="dtw").shape == (4,)
    assert W.pairwise_distance(x, x[0], metric="dtw").shape == (4,)
    # 3-D input: dim = "mean" / "full" / int (DI:1289-1302)
    x3, y3 = np.random.default_rng(1).standard_normal((5, 3, 30)), np.random.default_rng(2).standard_normal((6, 3, 30))
    per_dim = [oracle.pairwise("dtw", x3[:, d], y3[:, d], r=0.5) for d in range(3)]
    _eq(W.pairwise_distance(x3, y3, dim="full", metric="dtw", metric_params={"r": 0.5}), np.stack(per_dim), "full")
    _eq(W.pairwise_distance(x3, y3, dim="mean", metric="dtw", metric_params={"r": 0.5}), np.mean(per_dim, axis=0), "mean")
    _eq(W.pairwise_distance(x3, y3, dim=2, metric="dtw", metric_params={"r": 0.5}), per_dim[2], "dim=2")
    # non-contiguous rows (strided samples)
    big = random_walks(20, 40, 14)
    _eq(W.pairwise_distance(big[::2], big[1::2], metric="dtw"), oracle.pairwise("dtw", big[::2].copy(), big[1::2].copy()), "strided")
    # identical series -> exact zeros
    assert W.pairwise_distance(x, x.copy(), metric="dtw").diagonal().max() == 0.0


def test_paired_is_swapped_like_the_reference(W, oracle):
    x, y = random_walks(50, 60, 15), random_walks(50, 60, 16)
    for metric in ("wdtw", "adtw", "msm"):
        got = W.paired_distance(x, y, metric=metric, metric_params={"r": 0.3})
        _eq(got, oracle.paired(metric, x, y, r=0.3), metric)
        _eq(got, np.diag(oracle.pairwise(metric, y, x, r=0.3)), metric + " == diag(pairwise(y, x))")


@pytest.mark.parametrize("metric", METRICS)
@pytest.mark.parametrize("k", [1, 3, 7])
def test_argmin_matches_sequential_scan(W, oracle, metric, k):
    """Indices, distances AND heap order of the reference's scan (CD:1302-1345)."""
    x, y = random_walks(37, 48, 21), random_walks(150, 48, 22)
    y[60] = y[10]; y[61] = y[10]; x[5] = y[10]  # exact ties / zero distances
    for r in (0.1, 1.0):
        idx, dist = W.argmin_distance(x, y, k=k, metric=metric, metric_params={"r": r}, return_distance=True)
        oi, od = oracle.argmin(metric, x, y, k=k, r=r, n_jobs=0)
        _eq(idx, oi, f"{metric} k={k} r={r} idx")
        _eq(dist, od, f"{metric} k={k} r={r} dist")
    si, sd = W.argmin_distance(x, y, k=k, metric=metric, sorted=True, return_distance=True)
    assert np.all(np.diff(sd, axis=1) >= 0)


def test_argmin_self_join_and_lower_bound(W, oracle):
    x = random_walks(64, 40, 23)
    idx = W.argmin_distance(x, metric="dtw", metric_params={"r": 0.1})
    _eq(idx[:, 0], np.arange(64), "self join returns the self match")
    # user lower bound (LB_Keogh, reference lb.py:314-432) skips exactly like CD:1331
    q, refs = random_walks(20, 40, 24), random_walks(90, 40, 25)
    R = oracle.compute_r(40, 0.1)
    lb = np.zeros((20, 90))
    for j in range(90):
        lo, hi = oracle.envelope(refs[j], R)
        for i in range(20):
            lb[i, j] = oracle.lb_keogh_one(q[i], lo, hi)
    for k in (1, 4):
        idx, dist = W.argmin_distance(q, refs, k=k, metric="dtw", metric_params={"r": 0.1}, lower_bound=lb, return_distance=True)
        oi, od = oracle.argmin("dtw", q, refs, k=k, lower_bound=lb, r=0.1)
        _eq(idx, oi, "lb idx"); _eq(dist, od, "lb dist")


@pytest.mark.parametrize("k", [1, 5])
def test_argmin_device_cascade_is_exact(W, oracle, k, monkeypatch):
    """Many small chunks force the LB_Kim -> LB_Keogh -> list-mode DP cascade; results must not move."""
    monkeypatch.setenv("WILDBOAR_CUDA_ARGMIN_CHUNK", "64")
    q, refs = random_walks(70, 128, 41), random_walks(1500, 128, 42)
    refs[700] = q[3]; refs[20] = q[3]; refs[1499] = q[69]  # exact matches early, late and duplicated
    oi, od = oracle.argmin("dtw", q, refs, k=k, r=0.1, n_jobs=0)
    for lb in (True, False):
        idx, dist = W.argmin_distance(q, refs, k=k, metric="dtw", metric_params={"r": 0.1}, return_distance=True,
                                      device_lower_bound=lb)
        _eq(idx, oi, f"cascade={lb} idx"); _eq(dist, od, f"cascade={lb} dist")
        st = W.last_stats()
        if lb:
            assert st["lb_kim_pruned"] + st["lb_keogh_pruned"] > 0, st
            assert st["pairs"] < 70 * 1500
    # smooth series: LB_Keogh is tight, most pairs are pruned
    t = np.linspace(0, 6.28, 128)
    qs = np.sin(t)[None, :] * np.linspace(1, 2, 40)[:, None]
    rs = np.sin(t + 0.05)[None, :] * np.linspace(0.5, 3, 900)[:, None]
    idx, dist = W.argmin_distance(qs, rs, k=k, metric="dtw", metric_params={"r": 0.05}, return_distance=True)
    oi, od = oracle.argmin("dtw", qs, rs, k=k, r=0.05, n_jobs=0)
    _eq(idx, oi, "smooth idx"); _eq(dist, od, "smooth dist")


def test_argmin_cascade_odd_lengths_offsets_and_near_ties(W, oracle, monkeypatch):
    """The cascade's fp32 sums are rigorous lower bounds whatever the data looks like: odd lengths (tail of the 8-step
    blocks, time permutation with gcd handling), huge offsets (fp32 envelopes become loose, never wrong), tiny values,
    references that differ from the query by one ulp-sized step (distances right at the running threshold)."""
    monkeypatch.setenv("WILDBOAR_CUDA_ARGMIN_CHUNK", "32")
    rng = np.random.default_rng(77)
    for T, r in ((131, 0.1), (37, 0.3), (9, 0.5), (5, 1.0), (3, 0.1), (2, 1.0), (96, 0.02)):
        for scale, offset in ((1.0, 0.0), (1.0, 1e7), (1e-30, 0.0), (1e6, -3e9)):
            q = np.cumsum(rng.standard_normal((13, T)), axis=1) * scale + offset
            refs = np.cumsum(rng.standard_normal((330, T)), axis=1) * scale + offset
            refs[5] = q[2]
            refs[200] = np.nextafter(q[2], np.inf)          # almost the same series again, later in the scan
            refs[300] = q[7] + scale * 1e-9
            for k in (1, 4):
                idx, dist = W.argmin_distance(q, refs, k=k, metric="dtw", metric_params={"r": r}, return_distance=True)
                oi, od = oracle.argmin("dtw", q, refs, k=k, r=r, n_jobs=0)
                _eq(idx, oi, f"T={T} scale={scale} offset={offset} k={k} idx"); _eq(dist, od, f"T={T} scale={scale} offset={offset} k={k} dist")


def test_argmin_cfg4_shape_subset(W, oracle):
    """BASELINE configs[3] shape: T=256, r=0.05, k=1 -- 48 queries x 3000 references."""
    q, refs = random_walks(20000, 256, 3)[:48], random_walks(200000, 256, 4)[:3000]
    idx, dist = W.argmin_distance(q, refs, k=1, metric="dtw", metric_params={"r": 0.05}, return_distance=True)
    oi, od = oracle.argmin("dtw", q, refs, k=1, r=0.05, n_jobs=0)
    _eq(idx, oi, "cfg4 idx"); _eq(dist, od, "cfg4 dist")
    # size-independent property: argmin == argmin of the pairwise row
    full = W.pairwise_distance(q, refs, metric="dtw", metric_params={"r": 0.05})
    _eq(idx[:, 0], full.argmin(axis=1), "argmin == argmin(pairwise)")
    _eq(dist[:, 0], full.min(axis=1), "min == min(pairwise)")


def test_full_size_properties_cfg3(W, oracle):
    """Full BASELINE configs[2] rows are too slow for the oracle; check properties instead:
    a 1536 x 10000 slab, symmetry of equal-length DTW, and a random sample against the oracle."""
    x, y = random_walks(10000, 512, 1), random_walks(10000, 512, 2)
    xs = x[:1536]
    d = W.pairwise_distance(xs, y, metric="dtw", metric_params={"r": 0.1})
    assert d.shape == (1536, 10000) and np.isfinite(d).all() and (d >= 0).all()
    dt = W.pairwise_distance(y[:512], xs[:256], metric="dtw", metric_params={"r": 0.1})
    _eq(dt.T, d[:256, :512], "DTW(x,y) == DTW(y,x) for equal lengths")
    rng = np.random.default_rng(0)
    ii, jj = rng.integers(0, 1536, 300), rng.integers(0, 10000, 300)
    want = oracle.paired("dtw", y[jj], xs[ii], r=0.1, n_jobs=0)  # paired evaluates metric(second, first)
    _eq(d[ii, jj], want, "random sample vs oracle")


def test_multi_gpu_row_sharding_is_transparent(W, oracle):
    n = W.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    x, y = random_walks(203, 64, 31), random_walks(101, 64, 32)
    try:
        W.set_devices(list(range(n)))
        for metric in ("dtw", "msm"):
            _eq(W.pairwise_distance(x, y, metric=metric, metric_params={"r": 0.2}), oracle.pairwise(metric, x, y, r=0.2, n_jobs=0), metric)
            _eq(W.pairwise_distance(x, metric=metric, metric_params={"r": 0.2}), oracle.pairwise(metric, x, None, r=0.2, n_jobs=0), metric + " self")
            i, d = W.argmin_distance(x, y, k=3, metric=metric, metric_params={"r": 0.2}, return_distance=True)
            oi, od = oracle.argmin(metric, x, y, k=3, r=0.2, n_jobs=0)
            _eq(i, oi, metric + " argmin idx"); _eq(d, od, metric + " argmin dist")
        # a self join large enough for the multi-threaded host mirror (several devices: the lower triangle is transposed on
        # the host, destination block rows dealt out over threads), asymmetric metric
        X = random_walks(1500, 24, 33)
        got = W.pairwise_distance(X, metric="msm", metric_params={"r": 0.3})
        _eq(got, oracle.pairwise("msm", X, None, r=0.3, n_jobs=0), "msm self 1500, multi-device")
        assert np.array_equal(got, got.T)
    finally:
        W.set_devices([0])


# ---------------------------------------------------------------------------------------------
# optional fp32 mode (north star: <= 1e-4 relative error against the reference's fp64 values)
# ---------------------------------------------------------------------------------------------
FP32_RTOL = 1e-4
FP32_METRICS = ["dtw", "ddtw", "wdtw", "wddtw", "adtw", "erp", "msm", "twe"]


@pytest.fixture
def fp32(W):
    W.set_precision("fp32")
    yield W
    W.set_precision(None)


def _close32(got, want, what):
    got, want = np.asarray(got), np.asarray(want)
    assert got.shape == want.shape and got.dtype == np.float64, what
    err = np.abs(got - want) / np.maximum(np.abs(want), 1e-30)
    err = np.where(want == 0, np.abs(got), err)
    assert err.max() <= FP32_RTOL, (what, float(err.max()))
    return float(err.max())


@pytest.mark.parametrize("metric", FP32_METRICS)
def test_fp32_mode_within_tolerance(fp32, oracle, metric):
    """fp32 arithmetic, float64 in / float64 out: pairwise (both engines' geometries), singleton, paired."""
    for (nx, ny, Tx, Ty, r) in [(48, 100, 140, 140, 1.0), (40, 70, 256, 256, 0.05), (30, 45, 50, 77, 0.2), (6, 40, 3, 9, 0.5)]:
        if metric == "wddtw" and Tx > Ty:
            continue
        x, y = random_walks(nx, Tx, 11), random_walks(ny, Ty, 12)
        _close32(fp32.pairwise_distance(x, y, metric=metric, metric_params={"r": r}),
                 oracle.pairwise(metric, x, y, r=r, n_jobs=0), f"{metric} {Tx}x{Ty} r={r}")
    X = random_walks(120, 140, 13)
    _close32(fp32.pairwise_distance(X, metric=metric), oracle.pairwise(metric, X, None, n_jobs=0), metric + " singleton")
    _close32(fp32.paired_distance(X[:60], X[60:], metric=metric), oracle.paired(metric, X[:60], X[60:], n_jobs=0), metric + " paired")
    assert fp32.get_precision() == "fp32"


def test_fp32_mode_cfg3_and_cfg5_shapes(fp32, oracle):
    x, y = random_walks(10000, 512, 1)[:32], random_walks(10000, 512, 2)[:200]
    _close32(fp32.pairwise_distance(x, y, metric="dtw", metric_params={"r": 0.1}), oracle.pairwise("dtw", x, y, r=0.1, n_jobs=0), "cfg3 fp32")
    x, y = random_walks(2000, 4096, 1)[:4], random_walks(2000, 4096, 2)[:40]
    for metric in ("msm", "twe"):
        _close32(fp32.pairwise_distance(x, y, metric=metric, metric_params={"r": 0.05}),
                 oracle.pairwise(metric, x, y, r=0.05, n_jobs=0), f"cfg5 {metric} fp32")


@pytest.mark.parametrize("metric", ["lcss", "wlcss", "edr"])
def test_fp32_mode_keeps_threshold_metrics_exact(fp32, oracle, metric):
    """lcss / wlcss / edr are step functions of |x-y| vs eps: they stay in fp64 (bit-equal) in fp32 mode."""
    x, y = random_walks(40, 90, 7), random_walks(70, 90, 8)
    _eq(fp32.pairwise_distance(x, y, metric=metric, metric_params={"r": 0.3}), oracle.pairwise(metric, x, y, r=0.3, n_jobs=0), metric)


def test_fp32_argmin_indices_on_separated_data(fp32, oracle):
    """argmin in fp32 mode: distances within tolerance; indices equal wherever the fp64 margin between the
    best and the runner-up exceeds the tolerance."""
    q, refs = random_walks(64, 128, 21), random_walks(500, 128, 22)
    idx, dist = fp32.argmin_distance(q, refs, k=1, metric="dtw", metric_params={"r": 0.1}, return_distance=True)
    full = oracle.pairwise("dtw", q, refs, r=0.1, n_jobs=0)
    srt = np.sort(full, axis=1)
    clear = (srt[:, 1] - srt[:, 0]) > 4 * FP32_RTOL * srt[:, 1]
    assert clear.sum() > 32
    assert np.array_equal(idx[clear, 0], np.argmin(full, axis=1)[clear])
    _close32(dist[:, 0], full[np.arange(len(q)), idx[:, 0]], "argmin fp32 distances")


# ---- round 2: host pipeline (page-locked results, device-side mirror), fp64_fma mode, advisor corners ----
@pytest.mark.parametrize("metric", ["dtw", "msm", "erp", "ddtw"])
def test_self_join_mirrored_on_the_device(W, oracle, metric, monkeypatch):
    """Single-device self join: the kernel writes the lower triangle into the device-resident n x n matrix; with many
    small slabs (row0 != 0 in every chunk but the first) the result is still the reference's mirrored upper triangle --
    also for the asymmetric metrics, whose lower triangle is NOT d(x_j, x_i)."""
    X = random_walks(333, 40, 51)
    want = oracle.pairwise(metric, X, None, r=0.3, n_jobs=0)
    monkeypatch.setenv("WILDBOAR_CUDA_SLAB_KB", "64")
    got = W.pairwise_distance(X, metric=metric, metric_params={"r": 0.3})
    _eq(got, want, metric + " self, 64 KB slabs")
    assert np.array_equal(got, got.T) and not got.diagonal().any()
    monkeypatch.delenv("WILDBOAR_CUDA_SLAB_KB")
    _eq(W.pairwise_distance(X, metric=metric, metric_params={"r": 0.3}), want, metric + " self, one slab")
    # multivariate self join, dim="mean": the mirrored entries carry the combined value
    X3 = random_walks(90 * 3, 30, 52).reshape(90, 3, 30)
    per_dim = [oracle.pairwise(metric, np.ascontiguousarray(X3[:, d]), None, r=0.3, n_jobs=0) for d in range(3)]
    _eq(W.pairwise_distance(X3, dim="mean", metric=metric, metric_params={"r": 0.3}), np.mean(per_dim, axis=0), metric + " self mean")


def test_results_land_in_page_locked_memory_and_the_pool_recycles(W, oracle, monkeypatch):
    from wildboar_b200 import _shim
    x, y = random_walks(600, 32, 53), random_walks(700, 32, 54)
    want = oracle.pairwise("dtw", x, y, r=0.2, n_jobs=0)
    monkeypatch.setenv("WILDBOAR_CUDA_SLAB_KB", "256")  # > 2 slabs: exercises the event-ordered double buffering
    a = W.pairwise_distance(x, y, metric="dtw", metric_params={"r": 0.2})
    assert isinstance(a.base, _shim._PinnedBlock) or isinstance(getattr(a.base, "base", None), _shim._PinnedBlock)
    _eq(a, want, "pinned result")
    ptr = a.ctypes.data
    del a
    b = W.pairwise_distance(x, y, metric="dtw", metric_params={"r": 0.2})
    assert b.ctypes.data == ptr, "the released block was not recycled by the pool"
    _eq(b, want, "recycled pinned result")
    c = b.copy()  # ordinary memory; the pinned block goes back to the pool when b dies
    del b
    _eq(c, want, "copy")
    monkeypatch.setenv("WILDBOAR_CUDA_PINNED_RESULTS", "0")
    d = W.pairwise_distance(x, y, metric="dtw", metric_params={"r": 0.2})
    assert not isinstance(d.base, _shim._PinnedBlock)
    _eq(d, want, "pageable result")


@pytest.fixture
def fma(W):
    W.set_precision("fp64_fma")
    yield W
    W.set_precision(None)


@pytest.mark.parametrize("metric", ["dtw", "ddtw", "wdtw", "adtw", "wddtw"])
def test_fp64_fma_mode_within_1e12(fma, oracle, metric):
    """Optional fused-multiply-add mode (wb_params.precision = 2): <= 1e-12 relative against the reference (the north
    star's fp64 tolerance) -- written here as the test's tolerance; the default mode stays bit-equal."""
    worst = 0.0
    for (nx, ny, Tx, Ty, r) in [(48, 100, 140, 140, 1.0), (40, 70, 512, 512, 0.1), (30, 45, 50, 77, 0.2), (6, 40, 3, 9, 0.5)]:
        if metric == "wddtw" and Tx > Ty:
            continue
        x, y = random_walks(nx, Tx, 61), random_walks(ny, Ty, 62)
        got = fma.pairwise_distance(x, y, metric=metric, metric_params={"r": r})
        want = oracle.pairwise(metric, x, y, r=r, n_jobs=0)
        err = np.abs(got - want) / np.maximum(np.abs(want), 1e-300)
        assert err.max() <= 1e-12, (metric, Tx, Ty, r, float(err.max()))
        worst = max(worst, float(err.max()))
    assert fma.get_precision() == "fp64_fma"


def test_fp64_fma_mode_leaves_other_metrics_bit_equal_and_argmin_indices_exact(fma, oracle):
    x, y = random_walks(40, 90, 63), random_walks(70, 90, 64)
    for metric in ("msm", "twe", "erp", "lcss", "edr"):
        _eq(fma.pairwise_distance(x, y, metric=metric, metric_params={"r": 0.3}), oracle.pairwise(metric, x, y, r=0.3, n_jobs=0), metric)
    q, refs = random_walks(64, 128, 65), random_walks(500, 128, 66)
    idx, dist = fma.argmin_distance(q, refs, k=1, metric="dtw", metric_params={"r": 0.1}, return_distance=True)
    oi, od = oracle.argmin("dtw", q, refs, k=1, r=0.1)
    assert np.array_equal(idx, oi)
    assert (np.abs(dist - od) <= 1e-12 * np.abs(od)).all()


@pytest.mark.parametrize("k", [1, 3])
def test_argmin_adtw_negative_penalty_matches_the_scan(W, oracle, k):
    """adtw accepts a negative penalty (EL:3349); row minima are then not monotone and the reference's eadistance may
    abandon a pair whose final distance is below the threshold -- the replay has to use the row-minimum maxima."""
    q, refs = random_walks(24, 60, 71), random_walks(150, 60, 72)
    for p in (-0.5, -3.0):
        for r in (0.1, 1.0):
            idx, dist = W.argmin_distance(q, refs, k=k, metric="adtw", metric_params={"r": r, "p": p}, return_distance=True)
            oi, od = oracle.argmin("adtw", q, refs, k=k, r=r, p=p)
            _eq(idx, oi, f"adtw p={p} r={r} idx")
            _eq(dist, od, f"adtw p={p} r={r} dist")


@pytest.mark.parametrize("n_dims", [8, 9, 17, 130])
def test_dim_mean_single_output_follows_numpy_pairwise_summation(W, oracle, n_dims):
    """np.mean(list_of_matrices, axis=0) on ONE output element with >= 8 dimensions is a contiguous reduction that numpy
    sums pairwise; the mirror must return numpy's value there (advisor finding, round 1)."""
    x = random_walks(n_dims, 25, 81).reshape(1, n_dims, 25)
    y = random_walks(n_dims, 25, 82).reshape(1, n_dims, 25)
    per_dim = [oracle.pairwise("dtw", x[:, d], y[:, d], r=0.2, n_jobs=0) for d in range(n_dims)]
    got = W.pairwise_distance(x, y, dim="mean", metric="dtw", metric_params={"r": 0.2})
    assert np.array_equal(np.asarray(got).reshape(1, 1), np.mean(per_dim, axis=0))
    per_dim_p = [oracle.paired("dtw", x[:, d], y[:, d], r=0.2) for d in range(n_dims)]
    got = W.paired_distance(x, y, dim="mean", metric="dtw", metric_params={"r": 0.2})
    assert np.array_equal(np.asarray(got).reshape(1), np.mean(per_dim_p, axis=0))
    # more than one output element: sequential over the dimensions (device-combined) == numpy
    x3 = random_walks(3 * n_dims, 25, 83).reshape(3, n_dims, 25)
    per_dim = [oracle.pairwise("dtw", np.ascontiguousarray(x3[:, d]), y[:, d], r=0.2, n_jobs=0) for d in range(n_dims)]
    got = W.pairwise_distance(x3, y, dim="mean", metric="dtw", metric_params={"r": 0.2})
    assert np.array_equal(np.asarray(got).reshape(3, 1), np.mean(per_dim, axis=0))


# ---- round 2: cooperative engine (a group of lanes per pair, band row in registers, warp shuffles) ----
COOP_SHAPES = [(37, 53, 150, 150, 0.1), (20, 70, 140, 140, 1.0), (9, 40, 512, 512, 0.1), (30, 45, 50, 77, 0.2), (12, 33, 90, 61, 0.3),
               (5, 11, 600, 600, 0.02), (14, 25, 96, 96, 0.1), (8, 19, 200, 200, 0.05)]   # the last two: H = 17 / 19 -> the W = 4 layout


@pytest.mark.parametrize("metric", METRICS)
def test_coop_engine_matches_oracle(W, oracle, metric, monkeypatch):
    """WILDBOAR_CUDA_ENGINE=coop forces k_coop for every metric: pairwise (equal and unequal lengths, W = 8 and W = 13
    layouts, 4 to 32 lanes per pair), paired and the self join must equal the oracle bit for bit."""
    monkeypatch.setenv("WILDBOAR_CUDA_ENGINE", "coop")
    for (nx, ny, Tx, Ty, r) in COOP_SHAPES:
        if metric in ("wddtw",) and Tx > Ty:
            continue
        x, y = random_walks(nx, Tx, 91), random_walks(ny, Ty, 92)
        got = W.pairwise_distance(x, y, metric=metric, metric_params={"r": r})
        assert W.last_stats()["engine"] == 4, W.last_stats()
        _eq(got, oracle.pairwise(metric, x, y, r=r, n_jobs=0), f"coop {metric} {Tx}x{Ty} r={r}")
    X = random_walks(75, 150, 93)
    _eq(W.pairwise_distance(X, metric=metric, metric_params={"r": 0.1}), oracle.pairwise(metric, X, None, r=0.1, n_jobs=0), metric + " coop self")
    assert W.last_stats()["engine"] == 4
    _eq(W.paired_distance(X[:30], X[30:60], metric=metric, metric_params={"r": 0.1}),
        oracle.paired(metric, X[:30], X[30:60], r=0.1, n_jobs=0), metric + " coop paired")
    assert W.last_stats()["engine"] == 4


def test_coop_engine_is_selected_for_small_problems_and_tall_bands(W, oracle):
    """Automatic dispatch (measured, profiles/r02d_engines_coop_v3.jsonl): problems with too few pairs to give every SM a
    few thread-per-pair warps (one query against a set, DBA-sized lists) and the DTW family on tall bands (T = 4096,
    r = 0.05: per-pair boundary buffers would spill out of L2) run on the cooperative engine; cfg1 / cfg3-like shapes
    stay on the strip engine."""
    x, y = random_walks(1, 512, 1), random_walks(2000, 512, 2)
    _eq(W.pairwise_distance(x, y, metric="dtw", metric_params={"r": 0.1}), oracle.pairwise("dtw", x, y, r=0.1, n_jobs=0), "1 x 2000")
    st = W.last_stats()
    assert st["engine"] == 4 and st["strip_w"] == 8 and st["strip_nr"] == 16, st
    x, y = random_walks(2000, 4096, 1)[:6], random_walks(2000, 4096, 2)[:40]
    for metric in ("msm", "twe", "dtw"):
        _eq(W.pairwise_distance(x, y, metric=metric, metric_params={"r": 0.05}), oracle.pairwise(metric, x, y, r=0.05, n_jobs=0), "cfg5 " + metric)
        st = W.last_stats()
        assert st["engine"] == 4 and st["strip_w"] == 13 and st["strip_nr"] == 32, st
    x, y = random_walks(2000, 4096, 1)[:40], random_walks(2000, 4096, 2)[:1000]
    got = W.pairwise_distance(x, y, metric="dtw", metric_params={"r": 0.05})
    assert W.last_stats()["engine"] == 4        # tall band, DTW family: cooperative at any pair count
    ii, jj = np.arange(0, 40, 7), np.arange(3, 1000, 171)
    _eq(got[np.ix_(ii, jj)], oracle.pairwise("dtw", x[ii], y[jj], r=0.05, n_jobs=0), "cfg5 dtw sample")
    W.pairwise_distance(x, y, metric="msm", metric_params={"r": 0.05})
    assert W.last_stats()["engine"] == 2        # tall band, msm, enough warps: strip engine
    x, y = random_walks(200, 150, 1), random_walks(200, 150, 2)
    _eq(W.pairwise_distance(x, y, metric="dtw", metric_params={"r": 0.1}), oracle.pairwise("dtw", x, y, r=0.1, n_jobs=0), "cfg1")
    assert W.last_stats()["engine"] == 2
    x, y = random_walks(10000, 512, 1)[:64], random_walks(10000, 512, 2)[:3000]
    W.pairwise_distance(x, y, metric="dtw", metric_params={"r": 0.1})
    assert W.last_stats()["engine"] == 2


def test_argmin_cfg4_all_references_many_chunks(W, oracle):
    """BASELINE configs[3] at its full reference count: 64 of the 20 000 queries against ALL 200 000 references (dtw,
    r = 0.05, k = 1 and 3).  49 chunks of 4096 columns: the thresholds, the heap and the LB cascade's state are carried
    across every chunk boundary; indices, distances and heap order must equal the oracle's sequential scan."""
    q = random_walks(20000, 256, 3)[np.random.default_rng(7).choice(20000, 64, replace=False)]
    refs = random_walks(200000, 256, 4)
    for k in (1, 3):
        idx, dist = W.argmin_distance(q, refs, k=k, metric="dtw", metric_params={"r": 0.05}, return_distance=True)
        st = W.last_stats()
        oi, od = oracle.argmin("dtw", q, refs, k=k, r=0.05, n_jobs=0)
        _eq(idx, oi, f"cfg4 all refs k={k} idx")
        _eq(dist, od, f"cfg4 all refs k={k} dist")
        assert st["lb_kim_pruned"] + st["lb_keogh_pruned"] > 0.9 * 64 * 200000
    # a non-DTW metric over many chunks (row-minimum maxima replayed): 16 queries x 20 000 references
    q2, r2 = random_walks(16, 128, 5), random_walks(20000, 128, 6)
    for metric in ("msm", "erp"):
        idx, dist = W.argmin_distance(q2, r2, k=2, metric=metric, metric_params={"r": 0.05}, return_distance=True)
        oi, od = oracle.argmin(metric, q2, r2, k=2, r=0.05, n_jobs=0)
        _eq(idx, oi, metric + " idx")
        _eq(dist, od, metric + " dist")


def test_cfg5_full_length_random_entries(W, oracle):
    """BASELINE configs[4] at full length: a 64-row share against all 2000 series of length 4096 (msm / twe / dtw, r = 0.05),
    80 random entries per metric against the oracle -- the shipped engine per metric (strip for msm / twe at this pair count,
    cooperative for dtw)."""
    x, y = random_walks(2000, 4096, 1)[:64], random_walks(2000, 4096, 2)
    rng = np.random.default_rng(11)
    ii, jj = rng.integers(0, 64, 80), rng.integers(0, 2000, 80)
    for metric, engine in (("msm", 2), ("twe", 2), ("dtw", 4)):
        got = W.pairwise_distance(x, y, metric=metric, metric_params={"r": 0.05})
        assert W.last_stats()["engine"] == engine, (metric, W.last_stats())
        want = oracle.paired(metric, np.ascontiguousarray(y[jj]), np.ascontiguousarray(x[ii]), r=0.05, n_jobs=0)
        _eq(got[ii, jj], want, "cfg5 full length " + metric)


def test_concurrent_calls_from_several_threads(W, oracle):
    """The C ABI is re-entrant: every call leases its own device context (streams, events) and the page-locked pool and the
    context free list are shared under a lock.  Four Python threads (the GIL is released inside the library) run different
    metrics at once; every result must equal the oracle."""
    import threading
    x, y = random_walks(150, 100, 101), random_walks(400, 100, 102)
    jobs = [("dtw", 0.1), ("msm", 0.2), ("twe", 0.1), ("erp", 0.3), ("lcss", 0.2), ("adtw", 0.1), ("wdtw", 0.2), ("edr", 0.1)]
    want = {m: oracle.pairwise(m, x, y, r=r, n_jobs=0) for m, r in jobs}
    got, errs = {}, []

    def work(m, r):
        try:
            for _ in range(3):
                got[m] = W.pairwise_distance(x, y, metric=m, metric_params={"r": r})
                i, d = W.argmin_distance(x[:20], y, k=2, metric=m, metric_params={"r": r}, return_distance=True)
                got[m + "|argmin"] = (i, d)
        except Exception as e:  # noqa: BLE001
            errs.append((m, repr(e)))

    threads = [threading.Thread(target=work, args=j) for j in jobs]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errs, errs
    for m, r in jobs:
        _eq(got[m], want[m], "concurrent " + m)
        oi, od = oracle.argmin(m, x[:20], y, k=2, r=r, n_jobs=0)
        _eq(got[m + "|argmin"][0], oi, "concurrent argmin idx " + m)
        _eq(got[m + "|argmin"][1], od, "concurrent argmin dist " + m)


def test_fitted_set_keeps_the_cascade_operands_between_calls(W, oracle):
    """A device-resident reference set builds the reference-side operands of the LB cascade once (first argmin call) and
    reuses them; a call with another window must not be served from that cache.  Every call equals the oracle's scan."""
    from wildboar_b200 import _shim
    from wildboar_b200.distance import DtwMetric
    refs = random_walks(6000, 96, 121)
    fit = _shim.FittedSet(refs.reshape(6000, 1, 96), devices=[0])
    try:
        for r, seed in ((0.05, 1), (0.05, 2), (0.2, 3), (0.05, 4), (0.2, 5)):
            q = random_walks(40, 96, 130 + seed)
            m = DtwMetric(r=r)
            idx, dist = _shim.argmin_fitted(m.metric_id, m._params(), q, fit, 3, use_device_lb=True)
            oi, od = oracle.argmin("dtw", q, refs, k=3, r=r, n_jobs=0)
            _eq(idx, oi, f"fitted argmin r={r} call {seed} idx")
            _eq(dist, od, f"fitted argmin r={r} call {seed} dist")
            assert W.last_stats()["lb_kim_pruned"] + W.last_stats()["lb_keogh_pruned"] > 0
    finally:
        fit.close()


@pytest.mark.parametrize("metric,params", [("dtw", {"r": 0.1}), ("wdtw", {"r": 0.2, "g": 0.1}), ("adtw", {"r": 0.1, "p": 0.5})])
def test_argmin_pipelined_reference_upload(W, oracle, metric, params, monkeypatch):
    """Host-resident references are uploaded piece by piece while the scan is running (run_argmin / ensure_refs): tiny
    pieces and tiny chunks force dozens of pieces, with chunk and piece boundaries that do not line up; strided rows
    take the 2-D copy.  Results must equal the sequential scan and the all-at-once upload."""
    q, refs_full = random_walks(23, 96, 61), random_walks(1100, 200, 62)
    refs = refs_full[:, 50:146]            # rows 200 elements apart: a strided second operand
    assert not refs.flags.c_contiguous
    refs_c = np.ascontiguousarray(refs)
    oi, od = oracle.argmin(metric, q, refs_c, k=3, n_jobs=0, **params)
    monkeypatch.setenv("WILDBOAR_CUDA_ARGMIN_CHUNK", "96")
    for kb, y in (("0", refs_c), ("8", refs_c), ("8", refs), ("30", refs_c), ("1", refs_c)):
        monkeypatch.setenv("WILDBOAR_CUDA_PIPED_UPLOAD_KB", kb)
        for lb in (True, False):
            idx, dist = W.argmin_distance(q, y, k=3, metric=metric, metric_params=params, return_distance=True, device_lower_bound=lb)
            _eq(idx, oi, f"{metric} piece={kb} KB cascade={lb} idx"); _eq(dist, od, f"{metric} piece={kb} KB cascade={lb} dist")


def test_argmin_threshold_seeding_is_exact(W, oracle, monkeypatch):
    """k = 1: the thresholds are seeded with the exact distance to sketch-nearest candidates (run_argmin / k_seed_candidates).
    The seed only removes pairs that cannot be the minimum; ties must still resolve to the FIRST index, the candidate
    itself must survive, and k > 1 must not be seeded (heap order)."""
    monkeypatch.setenv("WILDBOAR_CUDA_SEED_MIN", "256")
    monkeypatch.setenv("WILDBOAR_CUDA_ARGMIN_CHUNK", "96")
    q, refs = random_walks(70, 128, 71), random_walks(1500, 128, 72)
    refs[1200] = q[3]; refs[40] = q[3]; refs[700] = q[3]       # the same exact match three times: index 40 wins
    refs[1499] = q[69]                                          # exact match in the last column
    refs[900] = refs[17]; refs[18] = refs[17]                   # duplicated references: ties for whoever is nearest to them
    q[5] = refs[17] + 1e-3
    for r in (0.1, 0.02, 1.0):
        oi, od = oracle.argmin("dtw", q, refs, k=1, r=r, n_jobs=0)
        monkeypatch.delenv("WILDBOAR_CUDA_NO_SEED", raising=False)
        idx, dist = W.argmin_distance(q, refs, k=1, metric="dtw", metric_params={"r": r}, return_distance=True)
        st_seed = W.last_stats()
        _eq(idx, oi, f"seeded r={r} idx"); _eq(dist, od, f"seeded r={r} dist")
        monkeypatch.setenv("WILDBOAR_CUDA_NO_SEED", "1")
        idx, dist = W.argmin_distance(q, refs, k=1, metric="dtw", metric_params={"r": r}, return_distance=True)
        st_plain = W.last_stats()
        _eq(idx, oi, f"unseeded r={r} idx"); _eq(dist, od, f"unseeded r={r} dist")
        if r < 1.0:
            assert st_seed["pairs"] < st_plain["pairs"], (st_seed, st_plain)   # fewer pairs reach the DP
    monkeypatch.delenv("WILDBOAR_CUDA_NO_SEED", raising=False)
    # pipelined upload: the candidates come from the first piece only (300 of the 1500 references)
    monkeypatch.setenv("WILDBOAR_CUDA_PIPED_UPLOAD_KB", "300")
    oi, od = oracle.argmin("dtw", q, refs, k=1, r=0.1, n_jobs=0)
    idx, dist = W.argmin_distance(q, refs, k=1, metric="dtw", metric_params={"r": 0.1}, return_distance=True)
    _eq(idx, oi, "seeded + piped idx"); _eq(dist, od, "seeded + piped dist")
    monkeypatch.delenv("WILDBOAR_CUDA_PIPED_UPLOAD_KB", raising=False)
    oi, od = oracle.argmin("dtw", q, refs, k=4, r=0.1, n_jobs=0)
    idx, dist = W.argmin_distance(q, refs, k=4, metric="dtw", metric_params={"r": 0.1}, return_distance=True)
    _eq(idx, oi, "k=4 idx (heap order)"); _eq(dist, od, "k=4 dist")
    # a fitted (device-resident) set takes the same path with all references as candidates
    from wildboar_b200 import _shim as sh
    from wildboar_b200.distance import DtwMetric
    m = DtwMetric(r=0.1)
    fit = sh.FittedSet(refs.reshape(len(refs), 1, -1), devices=[sh._first_device()])
    try:
        oi, od = oracle.argmin("dtw", q, refs, k=1, r=0.1, n_jobs=0)
        for _ in range(2):
            ridx, rdist = sh.argmin_fitted(m.metric_id, m._params(), q, fit, 1, use_device_lb=True)
            _eq(ridx, oi, "fitted seeded idx"); _eq(rdist, od, "fitted seeded dist")
    finally:
        fit.close()


@pytest.mark.parametrize("T,r", [(128, 0.1), (131, 0.05), (9, 0.5), (5, 1.0), (64, 0.3)])
def test_lb_prune_kernel_variants_prune_identically(W, oracle, T, r, monkeypatch):
    """The register-tiled LB pass (Q queries per warp, query tiles staged in shared memory by bulk copies, 4 or 8 time steps
    per register block, stragglers continued one pair per lane) computes per pair the sums of the one-query kernel in the
    same order: same results for every variant, and the same pruning counts as a pass that never leaves a block early;
    ragged query groups and reference blocks included."""
    monkeypatch.setenv("WILDBOAR_CUDA_ARGMIN_CHUNK", "160")
    monkeypatch.setenv("WILDBOAR_CUDA_NO_SEED", "1")
    monkeypatch.setenv("WILDBOAR_CUDA_LB_KEEP", "1")   # no adaptive switch-off of the pass: the counts are compared
    q, refs = random_walks(37, T, 81), random_walks(1003, T, 82)
    refs[600] = q[11]
    oi, od = oracle.argmin("dtw", q, refs, k=2, r=r, n_jobs=0)
    counts = set()
    for env in ({"LB_Q": "0"}, {"LB_Q": "2"}, {"LB_Q": "4"}, {"LB_Q": "4", "LB_BS": "8"}, {"LB_Q": "4", "LB_MINB": "2"},
                {"LB_Q": "8"}, {"LB_Q": "4", "LB_RB": "1"}, {"LB_Q": "8", "LB_RB": "3"}):
        for k_ in ("LB_Q", "LB_BS", "LB_MINB", "LB_RB"):
            monkeypatch.delenv("WILDBOAR_CUDA_" + k_, raising=False)
        for k_, v in env.items():
            monkeypatch.setenv("WILDBOAR_CUDA_" + k_, v)
        idx, dist = W.argmin_distance(q, refs, k=2, metric="dtw", metric_params={"r": r}, return_distance=True)
        _eq(idx, oi, f"{env} idx"); _eq(dist, od, f"{env} dist")
        st = W.last_stats()
        if env["LB_Q"] != "0":
            counts.add((st["lb_kim_pruned"], st["lb_keogh_pruned"], st["pairs"]))
        else:
            one_query = (st["lb_kim_pruned"], st["lb_keogh_pruned"], st["pairs"])
    assert len(counts) == 1, counts
    # the tiled pass finishes the sums of its stragglers (one pair per lane at the end of a CTA task) where the one-query
    # kernel hands them to the DP undecided: same LB_Kim count, at least as many LB_Keogh prunes, no more DP pairs
    tiled = next(iter(counts))
    assert tiled[0] == one_query[0] and tiled[1] >= one_query[1] and tiled[2] <= one_query[2], (tiled, one_query)
    # ... and exactly what the one-query kernel decides when it never leaves a block early
    monkeypatch.setenv("WILDBOAR_CUDA_LB_Q", "0"); monkeypatch.setenv("WILDBOAR_CUDA_LB_STRAG", "0,0")
    for k_ in ("LB_BS", "LB_MINB", "LB_RB"):
        monkeypatch.delenv("WILDBOAR_CUDA_" + k_, raising=False)
    idx, dist = W.argmin_distance(q, refs, k=2, metric="dtw", metric_params={"r": r}, return_distance=True)
    _eq(idx, oi, "no stragglers idx"); _eq(dist, od, "no stragglers dist")
    st = W.last_stats()
    assert (st["lb_kim_pruned"], st["lb_keogh_pruned"], st["pairs"]) == tiled, (st, tiled)


def test_argmin_neighbour_set_mode(W, oracle, monkeypatch):
    """use_device_lb bit 1 (KNeighborsClassifier.predict_proba only counts the k nearest): thresholds seeded for k > 1; the
    k neighbours per query are the reference's as a SET (and their distances as a multiset), ties at the kth distance are
    reported as ambiguous and the shim then falls back to the exact scan."""
    from wildboar_b200 import _shim as sh
    from wildboar_b200.distance import DtwMetric
    monkeypatch.setenv("WILDBOAR_CUDA_SEED_MIN", "256")
    monkeypatch.setenv("WILDBOAR_CUDA_ARGMIN_CHUNK", "128")
    q, refs = random_walks(60, 96, 91), random_walks(1400, 96, 92)
    m = DtwMetric(r=0.1)
    fit = sh.FittedSet(refs.reshape(len(refs), 1, -1), devices=[sh._first_device()])
    try:
        for k in (2, 5, 8):
            oi, od = oracle.argmin("dtw", q, refs, k=k, r=0.1, n_jobs=0)
            idx, dist = sh._argmin_fitted(m.metric_id, m._params(), q, fit, k, None, 3)
            st = sh._tls.stats
            assert st["ambiguous"] == 0, st
            _eq(np.sort(idx, axis=1), np.sort(oi, axis=1), f"k={k} neighbour sets")
            _eq(np.sort(dist, axis=1), np.sort(od, axis=1), f"k={k} distance multisets")
            plain_idx, _ = sh._argmin_fitted(m.metric_id, m._params(), q, fit, k, None, 1)
            assert st["pairs"] < sh._tls.stats["pairs"], (st, sh._tls.stats)   # the seeded scan sends fewer pairs to the DP
            _eq(plain_idx, oi, f"k={k} exact scan keeps the heap order")
    finally:
        fit.close()
    # ties at the kth distance: four copies of one reference, all equally near to query 0 -- which of them the reference keeps
    # depends on its scan history, so the mode reports the query and argmin_fitted(neighbour_set=True) returns the exact scan
    refs2 = refs.copy()
    for j in (100, 500, 900, 1300):
        refs2[j] = q[0] + 0.01
    fit = sh.FittedSet(refs2.reshape(len(refs2), 1, -1), devices=[sh._first_device()])
    try:
        oi, od = oracle.argmin("dtw", q, refs2, k=3, r=0.1, n_jobs=0)
        sh._argmin_fitted(m.metric_id, m._params(), q, fit, 3, None, 3)
        assert sh._tls.stats["ambiguous"] >= 1, sh._tls.stats
        idx, dist = sh.argmin_fitted(m.metric_id, m._params(), q, fit, 3, use_device_lb=True, neighbour_set=True)
        _eq(idx, oi, "fallback idx"); _eq(dist, od, "fallback dist")
    finally:
        fit.close()


def test_argmin_sorted_uses_the_set_mode_and_stays_exact(W, oracle, monkeypatch):
    """argmin_distance(sorted=True) and NearestNeighbors.kneighbors return the neighbours by distance, so they may run in the
    neighbour-set mode; rows with equal distances (duplicated references among the k nearest) must come back exactly as the
    reference orders them -- through the exact scan."""
    from wildboar_b200.neighbors import NearestNeighbors
    monkeypatch.setenv("WILDBOAR_CUDA_SEED_MIN", "256")
    q, refs = random_walks(50, 80, 95), random_walks(1200, 80, 96)

    def want(refs_, k):
        oi, od = oracle.argmin("dtw", q, refs_, k=k, r=0.1, n_jobs=0)
        order = np.argsort(od, axis=1, kind="stable")
        return np.take_along_axis(oi, order, axis=1), np.take_along_axis(od, order, axis=1)

    for k in (3, 5):
        idx, dist = W.argmin_distance(q, refs, k=k, metric="dtw", metric_params={"r": 0.1}, sorted=True, return_distance=True)
        st = W.last_stats()
        wi, wd = want(refs, k)
        _eq(idx, wi, f"k={k} sorted idx"); _eq(dist, wd, f"k={k} sorted dist")
        assert st["ambiguous"] == 0
    nn = NearestNeighbors(n_neighbors=4, metric="dtw", metric_params={"r": 0.1}).fit(refs)
    nd, ni = nn.kneighbors(q)
    wi, wd = want(refs, 4)
    _eq(ni, wi, "kneighbors idx"); _eq(nd, wd, "kneighbors dist")
    # the nearest reference of query 0 three times: equal distances inside the row -> exact scan, the reference's order
    refs2 = refs.copy()
    j0 = int(want(refs, 1)[0][0, 0])
    refs2[(j0 + 400) % 1200] = refs2[j0]; refs2[(j0 + 800) % 1200] = refs2[j0]
    idx, dist = W.argmin_distance(q, refs2, k=3, metric="dtw", metric_params={"r": 0.1}, sorted=True, return_distance=True)
    wi, wd = want(refs2, 3)
    _eq(idx, wi, "tied rows idx"); _eq(dist, wd, "tied rows dist")


def test_argmin_seeding_is_off_with_a_caller_lower_bound(W, oracle, monkeypatch):
    """The elastic ensemble masks every sample's own column with +inf in `lower_bound` (the reference skips pairs whose bound
    reaches the threshold whatever their distance): the distance to a sketch-nearest candidate -- here the sample itself,
    distance 0 -- is then no bound on what the scan accepts, so the thresholds must not be seeded."""
    monkeypatch.setenv("WILDBOAR_CUDA_SEED_MIN", "256")
    from wildboar_b200 import _shim as sh
    from wildboar_b200.distance import DtwMetric
    m = DtwMetric(r=0.1)
    x = random_walks(700, 64, 97)
    mask = np.full((700, 700), -np.inf)
    np.fill_diagonal(mask, np.inf)
    for k in (1, 3):
        oi, od = oracle.argmin("dtw", x, x, k=k, r=0.1, lower_bound=mask, n_jobs=0)
        # (the public argmin_distance refuses non-finite bounds like the reference's check_array; the ensemble calls the shim)
        idx, dist = sh.argmin(m.metric_id, m._params(), x, x, k, lower_bound=mask, use_device_lb=True)
        _eq(idx, oi, f"masked self join k={k} idx"); _eq(dist, od, f"masked self join k={k} dist")
        assert not (idx[:, 0] == np.arange(700)).any()


@pytest.mark.parametrize("metric,params", [("ddtw", {"r": 0.1}), ("adtw", {"r": 0.1, "p": 0.3}), ("adtw", {"r": 0.2, "p": 0.0})])
@pytest.mark.parametrize("k", [1, 4])
def test_argmin_cascade_covers_ddtw_and_adtw(W, oracle, metric, params, k, monkeypatch):
    """The LB cascade bounds the banded DTW of the prepared operands: valid for ddtw (slope series, eadistance's window) and for
    adtw with a non-negative penalty (DTW's cost plus penalties); seeded (k = 1) and unseeded (k = 4), many chunks."""
    monkeypatch.setenv("WILDBOAR_CUDA_SEED_MIN", "256")
    monkeypatch.setenv("WILDBOAR_CUDA_ARGMIN_CHUNK", "96")
    q, refs = random_walks(45, 100, 98), random_walks(1300, 100, 99)
    refs[900] = q[4]; refs[77] = q[4]
    oi, od = oracle.argmin(metric, q, refs, k=k, n_jobs=0, **params)
    for lb in (True, False):
        idx, dist = W.argmin_distance(q, refs, k=k, metric=metric, metric_params=params, return_distance=True, device_lower_bound=lb)
        _eq(idx, oi, f"{metric} cascade={lb} idx"); _eq(dist, od, f"{metric} cascade={lb} dist")
        st = W.last_stats()
        if lb and metric == "adtw":   # (ddtw of random walks compares noise: its bounds prune next to nothing)
            assert st["lb_kim_pruned"] + st["lb_keogh_pruned"] > 0, st


def test_argmin_cascade_switches_itself_off_where_it_does_not_prune(W, oracle, monkeypatch):
    """ddtw of random walks: the slope series are noise, their envelopes contain almost every sample and the LB pass prunes
    next to nothing -- after the first columns it is dropped (the pruned fraction is read back once); results unchanged."""
    q, refs = random_walks(30, 128, 101), random_walks(6000, 128, 102)
    oi, od = oracle.argmin("ddtw", q, refs, k=2, r=0.05, n_jobs=0)
    idx, dist = W.argmin_distance(q, refs, k=2, metric="ddtw", metric_params={"r": 0.05}, return_distance=True)
    _eq(idx, oi, "ddtw idx"); _eq(dist, od, "ddtw dist")
    st_auto = W.last_stats()
    monkeypatch.setenv("WILDBOAR_CUDA_LB_KEEP", "1")
    idx, dist = W.argmin_distance(q, refs, k=2, metric="ddtw", metric_params={"r": 0.05}, return_distance=True)
    _eq(idx, oi, "ddtw idx (pass kept)"); _eq(dist, od, "ddtw dist (pass kept)")
    st_keep = W.last_stats()
    assert st_auto["launches"] < st_keep["launches"], (st_auto, st_keep)
