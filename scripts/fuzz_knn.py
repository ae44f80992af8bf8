"""DEV TOOLING: KNeighborsClassifier.predict_proba / NearestNeighbors.kneighbors (neighbour-set mode, fitted sets) and larger
argmin cases against the CPU oracle."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import wildboar_b200 as wb
from wildboar_b200.neighbors import KNeighborsClassifier, NearestNeighbors
from oracle import oracle as O

rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 3)
wb.set_devices([0])
os.environ["WILDBOAR_CUDA_SEED_MIN"] = "256"
bad = 0
t0 = time.time()
for case in range(60):
    T = int(rng.choice([16, 50, 96, 128])); r = float(rng.choice([0.05, 0.1, 0.3]))
    nq = int(rng.integers(1, 90)); ny = int(rng.integers(300, 5000)); k = int(rng.choice([1, 2, 3, 5, 7, 8]))
    metric = str(rng.choice(["dtw", "dtw", "adtw", "ddtw"]))
    q = np.cumsum(rng.standard_normal((nq, T)), axis=1); refs = np.cumsum(rng.standard_normal((ny, T)), axis=1)
    if rng.random() < 0.5:
        for _ in range(int(rng.integers(1, 5))):
            a, b = int(rng.integers(0, ny)), int(rng.integers(0, ny))
            refs[b] = refs[a]
            if rng.random() < 0.5: q[int(rng.integers(0, nq))] = refs[a] + 1e-3
    y = rng.integers(0, 3, ny)
    params = {"r": r}
    oi, od = O.argmin(metric, q, refs, k=k, n_jobs=0, **params)
    votes = y[oi]
    want = np.stack([(votes == c).sum(1) / k for c in range(3)], axis=1)
    clf = KNeighborsClassifier(k, metric=metric, metric_params=params).fit(refs, y)
    got = clf.predict_proba(q)
    ok1 = np.array_equal(got, want)
    nn = NearestNeighbors(n_neighbors=k, metric=metric, metric_params=params).fit(refs)
    nd, ni = nn.kneighbors(q)
    order = np.argsort(od, axis=1, kind="stable")
    ok2 = np.array_equal(ni, np.take_along_axis(oi, order, axis=1)) and np.array_equal(nd, np.take_along_axis(od, order, axis=1))
    if not (ok1 and ok2):
        bad += 1
        print("MISMATCH knn case", case, dict(T=T, r=r, nq=nq, ny=ny, k=k, metric=metric), ok1, ok2, flush=True)
print(f"knn fuzz: 60 cases, {bad} mismatches, {time.time() - t0:.0f} s", flush=True)
# larger scans: wide chunks, many CTA tasks, straggler queues, pipelined upload in default pieces
for case in range(6):
    for k_ in ("WILDBOAR_CUDA_PIPED_UPLOAD_KB",):
        os.environ.pop(k_, None)
    nq = int(rng.choice([3, 64, 300])); ny = int(rng.choice([20000, 60000])); T = 256; k = int(rng.choice([1, 1, 5]))
    if case % 2: os.environ["WILDBOAR_CUDA_PIPED_UPLOAD_KB"] = "4096"
    q = np.cumsum(rng.standard_normal((nq, T)), axis=1); refs = np.cumsum(rng.standard_normal((ny, T)), axis=1)
    refs[ny // 2] = q[0]; refs[ny // 3] = q[0]
    srt = k > 1
    t1 = time.time()
    oi, od = O.argmin("dtw", q, refs, k=k, r=0.05, n_jobs=os.cpu_count() or 1)
    if srt:
        order = np.argsort(od, axis=1, kind="stable"); oi, od = np.take_along_axis(oi, order, axis=1), np.take_along_axis(od, order, axis=1)
    idx, dist = wb.argmin_distance(q, refs, k=k, metric="dtw", metric_params={"r": 0.05}, sorted=srt, return_distance=True)
    ok = np.array_equal(idx, oi) and np.array_equal(dist, od)
    print("large case", case, dict(nq=nq, ny=ny, k=k, piped=os.environ.get("WILDBOAR_CUDA_PIPED_UPLOAD_KB")), "ok" if ok else "MISMATCH", f"{time.time() - t1:.1f} s", wb.last_stats()["pairs"], flush=True)
    bad += 0 if ok else 1
sys.exit(1 if bad else 0)
