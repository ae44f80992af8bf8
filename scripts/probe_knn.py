"""DEV TOOLING: KNeighborsClassifier(k).predict on the cfg4 shape (2500 queries vs 200 000 resident series x 256)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import wildboar_b200 as wb
from wildboar_b200.neighbors import KNeighborsClassifier
wb.set_devices([0])
rw = lambda n, T, s: np.cumsum(np.random.default_rng(s).standard_normal((n, T)), axis=1)
q, refs = rw(20000, 256, 3)[:2500], rw(200000, 256, 4)
y = np.random.default_rng(5).integers(0, 4, len(refs))
for k in (1, 5):
    clf = KNeighborsClassifier(k, metric="dtw", metric_params={"r": 0.05}).fit(refs, y)
    for rep in range(3):
        t0 = time.perf_counter()
        p = clf.predict(q)
        dt = time.perf_counter() - t0
        st = wb.last_stats()
        print("k =", k, "predict %.1f ms" % (dt * 1e3), {n: st[n] for n in ("kernel_ms", "launches", "pairs", "ambiguous")}, flush=True)
    if k == 5:
        os.environ["WILDBOAR_CUDA_NO_SEED"] = "1"
        t0 = time.perf_counter(); p2 = clf.predict(q); dt = time.perf_counter() - t0
        print("k = 5 unseeded predict %.1f ms" % (dt * 1e3), "same labels:", bool(np.array_equal(p, p2)))
        os.environ.pop("WILDBOAR_CUDA_NO_SEED")
