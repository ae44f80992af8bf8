#!/usr/bin/env python
"""Measurement of the SURVEY 8f "next" rows on ONE GPU, with the reference timed beside them.

Not the driver's bench line (that is bench.py).  Every row is one JSON line: what ran, the wall time of the
public API call (host buffers in and out), the device time of the kernels where the library reports it, and the
UNMODIFIED reference (oracle/_ref) timed on the box's host cores on a stated, bounded sample of the same
workload (`ref_*` keys; GCUPS / per-item times are size independent for these loops).

  python scripts/bench_next.py [--quick] [--no-ref]
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import wildboar_b200 as wb  # noqa: E402
from wildboar_b200 import dtw as wdtw  # noqa: E402
from wildboar_b200.neighbors import KMeans, KNeighborsClassifier  # noqa: E402

QUICK = "--quick" in sys.argv
NO_REF = "--no-ref" in sys.argv
NCPU = os.cpu_count() or 1


def rw(n, T, seed, dims=None):
    shape = (n, T) if dims is None else (n, dims, T)
    return np.cumsum(np.random.default_rng(seed).standard_normal(shape), axis=-1)


def cells(T, r, Tb=None):
    Tb = T if Tb is None else Tb
    R = max(int(np.floor(min(T, Tb) * r)), 1)
    return T * (2 * R - 1) - R * (R - 1) if T == Tb else None


def timed(fn, reps=2):
    fn()
    best = 1e30
    for _ in range(reps):
        t0 = time.perf_counter()
        out = fn()
        best = min(best, time.perf_counter() - t0)
    return best, out


def emit(**row):
    print(json.dumps(row), flush=True)


def ref_modules():
    if NO_REF:
        return None, None
    from oracle import ref
    wd = ref.load()
    if wd is None:
        return None, None
    from wildboar.distance import dtw as rd
    return wd, rd


def main():
    wb.set_devices([0])
    wd, rd = ref_modules()

    # ---- 8f-3a: multivariate dim="mean", combined on the device vs one library call per dimension ----
    n, nd, T = (800, 3, 140) if QUICK else (3000, 3, 140)
    x3, y3 = rw(n, T, 1, nd), rw(n, T, 2, nd)
    for metric in ("dtw", "msm"):
        t_fused, got = timed(lambda: wb.pairwise_distance(x3, y3, dim="mean", metric=metric))
        st = wb.last_stats()
        t_loop, want = timed(lambda: np.mean([wb.pairwise_distance(np.ascontiguousarray(x3[:, d]), np.ascontiguousarray(y3[:, d]), metric=metric)
                                              for d in range(nd)], axis=0))
        row = dict(row="8f-3 multivariate dim=mean", metric=metric, shape=f"{n}x{nd}x{T} vs {n}x{nd}x{T} r=1.0",
                   cells=st["cells"], fused_e2e_ms=round(t_fused * 1e3, 2), fused_kernel_ms=round(st["kernel_ms"], 2),
                   fused_e2e_gcups=round(st["cells"] / t_fused / 1e9, 1), per_dim_loop_e2e_ms=round(t_loop * 1e3, 2),
                   bit_equal_to_per_dim_loop=bool(np.array_equal(got, want)))
        if wd is not None:
            ns = 96 if QUICK else 192
            t_ref, _ = timed(lambda: wd.pairwise_distance(x3[:ns], y3[:ns], dim="mean", metric=metric, n_jobs=NCPU), reps=1)
            row.update(ref_sample=f"first {ns} x rows vs first {ns} y rows, n_jobs={NCPU}",
                       ref_gcups=round(ns * ns * nd * cells(T, 1.0) / t_ref / 1e9, 3), ref_cores=NCPU)
        emit(**row)

    # ---- 8f-1: KNeighborsClassifier.predict with the training set resident on the device ----
    ntrain, nq, T = (20000, 256, 256) if QUICK else (200000, 2500, 256)
    Xtr, Q = rw(ntrain, T, 4), rw(nq, T, 3)
    ytr = np.random.default_rng(0).integers(0, 5, ntrain)
    mp = {"r": 0.05}
    clf = KNeighborsClassifier(n_neighbors=1, metric="dtw", metric_params=mp)
    t0 = time.perf_counter(); clf.fit(Xtr, ytr); t_fit = time.perf_counter() - t0
    t_res, pred = timed(lambda: clf.predict(Q))
    st = wb.last_stats()
    t_host, idx = timed(lambda: wb.argmin_distance(Q, Xtr, k=1, metric="dtw", metric_params=mp))
    nominal = nq * ntrain * cells(T, 0.05)
    row = dict(row="8f-1 KNeighborsClassifier.predict (k=1, dtw r=0.05)", shape=f"{nq} queries vs {ntrain} training series x {T}",
               fit_upload_ms=round(t_fit * 1e3, 1), predict_resident_ms=round(t_res * 1e3, 1), kernel_ms=round(st["kernel_ms"], 1),
               argmin_host_buffers_ms=round(t_host * 1e3, 1), nominal_gcups_resident=round(nominal / t_res / 1e9, 1),
               nominal_gcups_host_buffers=round(nominal / t_host / 1e9, 1),
               same_labels=bool(np.array_equal(pred, ytr[idx[:, 0]])), lb_pruned=st["lb_kim_pruned"] + st["lb_keogh_pruned"], pairs=nq * ntrain)
    if wd is not None:
        nqs, nts = (16, 2000) if QUICK else (64, 4096)
        rclf = wd.KNeighborsClassifier(n_neighbors=1, metric="dtw", metric_params=mp, n_jobs=NCPU).fit(Xtr[:nts], ytr[:nts])
        t_ref, rp = timed(lambda: rclf.predict(Q[:nqs]), reps=1)
        ours = KNeighborsClassifier(n_neighbors=1, metric="dtw", metric_params=mp).fit(Xtr[:nts], ytr[:nts]).predict(Q[:nqs])
        row.update(ref_sample=f"{nqs} queries vs {nts} training series, n_jobs={NCPU}", ref_ms=round(t_ref * 1e3, 1),
                   ref_nominal_gcups=round(nqs * nts * cells(T, 0.05) / t_ref / 1e9, 2), ref_cores=NCPU,
                   ref_labels_equal=bool(np.array_equal(rp, ours)))
    emit(**row)
    clf.release()

    # ---- 8f-3b: warping paths (batched) ----
    n, T, r = (2000, 512, 0.1) if QUICK else (10000, 512, 0.1)
    a, b = rw(n, T, 5), rw(n, T, 6)
    t_paths, (lo, hi) = timed(lambda: wdtw.dtw_paths(a, b, r=r))
    st = wb.last_stats()
    row = dict(row="8f-3 dtw_paths (alignment + back-walk on the device)", shape=f"{n} pairs x {T} r={r}", e2e_ms=round(t_paths * 1e3, 2),
               kernel_ms=round(st["kernel_ms"], 2), kernel_gcups=round(st["cells"] / (st["kernel_ms"] * 1e-3) / 1e9, 1),
               pairs_per_s=round(n / t_paths), launches=st["launches"])
    if rd is not None:
        ns = 16 if QUICK else 48
        t_ref, _ = timed(lambda: [rd.dtw_mapping(a[i], b[i], r=r) for i in range(ns)], reps=1)
        row.update(ref_sample=f"{ns} pairs, dtw_alignment + dtw_mapping (1 core: the back-walk is a Python loop)",
                   ref_pairs_per_s=round(ns / t_ref, 1), ref_cores=1)
    emit(**row)

    # ---- 8f-3c: DBA (dtw_average, method="mm") ----
    n, T, r, ep = (500, 256, 0.1, 5) if QUICK else (2000, 512, 0.1, 5)
    X = rw(n, T, 7)
    t_dba, (mean, cost) = timed(lambda: wdtw.dtw_average(X, r=r, init=X[0], max_epoch=ep, tol=0.0, return_cost=True), reps=1)
    row = dict(row="8f-3 dtw_average (DBA, mm)", shape=f"{n} samples x {T}, r={r}, {ep} epochs", e2e_ms=round(t_dba * 1e3, 1),
               ms_per_epoch=round(t_dba * 1e3 / ep, 2), alignments_per_s=round(n * ep / t_dba))
    if rd is not None:
        ns, eps = (24, 2) if QUICK else (64, 2)
        t_ref, (rmean, rcost) = timed(lambda: rd.dtw_average(X[:ns], r=r, init=X[0], max_epoch=eps, tol=0.0, return_cost=True), reps=1)
        omean, ocost = wdtw.dtw_average(X[:ns], r=r, init=X[0], max_epoch=eps, tol=0.0, return_cost=True)
        row.update(ref_sample=f"{ns} samples, {eps} epochs (1 core: Python loops over samples and path cells)",
                   ref_alignments_per_s=round(ns * eps / t_ref, 1), ref_cores=1,
                   ref_bit_equal=bool(np.array_equal(rmean, omean) and rcost == ocost))
    emit(**row)

    # ---- 8f-4: subsequence search (shapelet distances), DTW family ----
    n, T, ns, m = (300, 256, 8, 48) if QUICK else (2000, 512, 64, 64)
    Xs = rw(n, T, 8)
    rng = np.random.default_rng(9)
    shp = [Xs[rng.integers(0, n), o:o + m].copy() for o in rng.integers(0, T - m, ns)]
    for metric, mp in (("dtw", {"r": 0.1}), ("wdtw", {"r": 0.1, "g": 0.05}), ("scaled_dtw", {"r": 0.1})):
        t_ss, (d, i) = timed(lambda: wb.pairwise_subsequence_distance(shp, Xs, metric=metric, metric_params=mp, return_index=True))
        st = wb.last_stats()
        row = dict(row="8f-4 pairwise_subsequence_distance", metric=metric, shape=f"{ns} subsequences x {m} vs {n} samples x {T}, r={mp['r']}",
                   windows=n * (T - m + 1) * ns, cells=st["cells"], e2e_ms=round(t_ss * 1e3, 1), kernel_ms=round(st["kernel_ms"], 1),
                   kernel_gcups=round(st["cells"] / (st["kernel_ms"] * 1e-3) / 1e9, 1), launches=st["launches"])
        if wd is not None:
            nss, nxs = (4, 48) if QUICK else (8, 128)
            t_ref, (rd_, ri_) = timed(lambda: wd.pairwise_subsequence_distance(shp[:nss], Xs[:nxs], metric=metric, metric_params=mp,
                                                                             return_index=True, n_jobs=NCPU), reps=1)
            row.update(ref_sample=f"{nss} subsequences vs {nxs} samples, n_jobs={NCPU} (early abandoning)", ref_ms=round(t_ref * 1e3, 1),
                       ref_windows_per_s=round(nss * nxs * (T - m + 1) / t_ref), windows_per_s=round(n * (T - m + 1) * ns / t_ss),
                       ref_cores=NCPU, ref_bit_equal=bool(np.array_equal(rd_, d[:nxs, :nss]) and np.array_equal(ri_, i[:nxs, :nss])))
        emit(**row)

    # ---- 8f-1: ElasticEnsembleClassifier.fit (leave-one-out grid search over 9 metrics, 74 candidates) ----
    from wildboar_b200.ensemble import ElasticEnsembleClassifier
    n, T = (60, 64) if QUICK else (400, 128)
    Xe = np.concatenate([rw(n // 2, T, 21), rw(n // 2, T, 22) + 2.0])
    ye = np.repeat([0, 1], n // 2)
    t_ee, clf = timed(lambda: ElasticEnsembleClassifier(n_neighbors=1, metric="auto").fit(Xe, ye), reps=1)
    row = dict(row="8f-1 ElasticEnsembleClassifier(metric=auto).fit", shape=f"{n} samples x {T}, 9 metrics, 74 candidates, leave-one-out",
               e2e_ms=round(t_ee * 1e3, 1), folds=n * 74, best=[(m, round(float(s_), 4)) for m, s_ in clf.scores_][:3])
    if wd is not None:
        from wildboar.ensemble import ElasticEnsembleClassifier as RefEE
        ns = 30 if QUICK else 60
        sel = np.r_[0:ns // 2, n // 2:n // 2 + ns // 2]
        t_ref, rclf = timed(lambda: RefEE(n_neighbors=1, metric="auto", n_jobs=NCPU).fit(Xe[sel], ye[sel]), reps=1)
        t_our, oclf = timed(lambda: ElasticEnsembleClassifier(n_neighbors=1, metric="auto").fit(Xe[sel], ye[sel]), reps=1)
        row.update(ref_sample=f"{ns} samples (reference: GridSearchCV + LeaveOneOut, n_jobs={NCPU})", ref_ms=round(t_ref * 1e3, 1),
                   ours_same_sample_ms=round(t_our * 1e3, 1), ref_cores=NCPU,
                   ref_equal=bool([m for m, _ in rclf.scores_] == [m for m, _ in oclf.scores_] and
                                  np.array_equal([s_ for _, s_ in rclf.scores_], [s_ for _, s_ in oclf.scores_]) and
                                  np.array_equal(rclf.predict_proba(Xe[::7]), oclf.predict_proba(Xe[::7]))))
    emit(**row)

    # ---- 8f-1: KMeans(metric="dtw") ----
    n, T, K = (300, 128, 4) if QUICK else (2000, 256, 8)
    Xk = np.concatenate([rw(n // K, T, 10 + c) + 8.0 * c for c in range(K)])
    args = dict(n_clusters=K, metric="dtw", r=0.1, random_state=1, max_iter=10)
    t_km, km = timed(lambda: KMeans(**args).fit(Xk), reps=1)
    row = dict(row="8f-1 KMeans(metric=dtw).fit", shape=f"{Xk.shape[0]} samples x {T}, K={K}, r=0.1, max_iter=10", e2e_ms=round(t_km * 1e3, 1),
               n_iter=int(km.n_iter_), inertia=float(km.inertia_))
    if wd is not None:
        ns = 60 if QUICK else 120
        sub = np.ascontiguousarray(Xk[:: max(1, Xk.shape[0] // ns)][:ns])
        sargs = dict(args, n_clusters=3, max_iter=3)
        t_ref, rkm = timed(lambda: wd.KMeans(**sargs).fit(sub), reps=1)
        t_our, okm = timed(lambda: KMeans(**sargs).fit(sub), reps=1)
        row.update(ref_sample=f"{sub.shape[0]} samples, K=3, max_iter=3 (reference, 1 core)", ref_ms=round(t_ref * 1e3, 1),
                   ours_same_sample_ms=round(t_our * 1e3, 1),
                   ref_bit_equal=bool(np.array_equal(rkm.cluster_centers_, okm.cluster_centers_) and np.array_equal(rkm.labels_, okm.labels_)))
    emit(**row)


if __name__ == "__main__":
    main()
