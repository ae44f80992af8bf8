"""SURVEY 8f "next" rows: multivariate dim="mean"/"full" combined on the device.

CPU tests pin the oracle (+ numpy's reduction order) to the reference's golden vectors; `-m gpu` tests
compare the CUDA path (public API -> ctypes -> C ABI) with the oracle and the golden vectors, bit for bit.
"""
import os

import numpy as np
import pytest

from util import random_walks

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
ND_METRICS = ["dtw", "wdtw", "ddtw", "adtw", "lcss", "erp", "edr", "msm", "twe"]


@pytest.fixture(scope="module")
def next_golden():
    with np.load(os.path.join(ROOT, "tests", "golden", "next_golden.npz")) as z:
        return {k: z[k] for k in z.files}


def _oracle_nd(oracle, kind, metric, x, y, dim, **mp):
    """Per-dimension oracle calls combined like the reference (np.mean / np.stack over axis 0)."""
    nd = x.shape[1]
    if kind == "pairwise":
        per = [oracle.pairwise(metric, x[:, d], y[:, d], **mp) for d in range(nd)]
    elif kind == "self":
        per = [oracle.pairwise(metric, x[:, d], None, **mp) for d in range(nd)]
    else:
        per = [oracle.paired(metric, x[:, d], y[:, d], **mp) for d in range(nd)]
    return np.mean(per, axis=0) if dim == "mean" else np.stack(per, axis=0)


# ---------------------------------------------------------------------------------------------
# CPU: the oracle restatement equals the reference on multivariate input
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("metric", ND_METRICS)
def test_oracle_multivariate_matches_reference_golden(oracle, next_golden, metric):
    x, y = next_golden["nd|x"], next_golden["nd|y"]
    for dim in ("mean", "full"):
        assert np.array_equal(_oracle_nd(oracle, "pairwise", metric, x, y, dim, r=0.25), next_golden[f"nd|{metric}|{dim}|pairwise"])
        assert np.array_equal(_oracle_nd(oracle, "self", metric, x, None, dim, r=0.25), next_golden[f"nd|{metric}|{dim}|self"])
        assert np.array_equal(_oracle_nd(oracle, "paired", metric, x[:5], y, dim, r=0.25), next_golden[f"nd|{metric}|{dim}|paired"])


# ---------------------------------------------------------------------------------------------
# GPU
# ---------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def W(wb):
    if wb.device_count() < 1:
        pytest.skip("no CUDA device")
    wb.set_devices([0])
    return wb


@pytest.mark.gpu
@pytest.mark.parametrize("metric", ND_METRICS)
def test_multivariate_golden(W, next_golden, metric):
    x, y = next_golden["nd|x"], next_golden["nd|y"]
    mp = {"r": 0.25}
    for dim in ("mean", "full"):
        assert np.array_equal(W.pairwise_distance(x, y, dim=dim, metric=metric, metric_params=mp), next_golden[f"nd|{metric}|{dim}|pairwise"])
        assert np.array_equal(W.pairwise_distance(x, dim=dim, metric=metric, metric_params=mp), next_golden[f"nd|{metric}|{dim}|self"])
        assert np.array_equal(W.paired_distance(x[:5], y, dim=dim, metric=metric, metric_params=mp), next_golden[f"nd|{metric}|{dim}|paired"])
        assert W.last_stats()["launches"] >= x.shape[1]  # every dimension ran on the device, in one library call


@pytest.mark.gpu
def test_multivariate_large_strided_and_chunked(W, oracle):
    """Several result slabs per matrix (48 MB each), strided 3-D views, unequal lengths, multi-device split."""
    rng = np.random.default_rng(5)
    big = np.cumsum(rng.standard_normal((900, 4, 2, 48)), axis=3)   # view [:, ::2] has a dim stride of 2 * 2 * 48
    x = big[:, ::2, 0, :]                                            # (900, 2, 48), non-contiguous samples and dims
    y = np.cumsum(rng.standard_normal((8000, 2, 40)), axis=2)        # 900 x 8000 doubles = 57.6 MB -> two slabs
    for metric in ("dtw", "erp"):
        for dim in ("mean", "full"):
            got = W.pairwise_distance(x, y, dim=dim, metric=metric, metric_params={"r": 0.1})
            sel = rng.integers(0, 900, 40)
            want = _oracle_nd(oracle, "pairwise", metric, np.ascontiguousarray(x[sel]), y, dim, r=0.1)
            assert np.array_equal(got[..., sel, :], want), (metric, dim)
    xs = np.ascontiguousarray(x[:300])
    got = W.pairwise_distance(xs, dim="mean", metric="msm", metric_params={"r": 0.2})
    assert np.array_equal(got, _oracle_nd(oracle, "self", "msm", xs, None, "mean", r=0.2))
    if W.device_count() >= 2:
        W.set_devices([0, 1])
        try:
            two = W.pairwise_distance(xs, dim="mean", metric="msm", metric_params={"r": 0.2})
        finally:
            W.set_devices([0])
        assert np.array_equal(two, got)
