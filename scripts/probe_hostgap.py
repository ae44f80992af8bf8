"""DEV TOOLING: where do the 50-70 ms between the device-side total and the wall clock of ONE of bench.py's cfg5 calls go?"""
import os, sys, time, gc
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import wildboar_b200 as wb
from wildboar_b200 import _shim, distance
wb.set_devices([0])
rw = lambda n, T, s: np.cumsum(np.random.default_rng(s).standard_normal((n, T)), axis=1)
q, refs = rw(20000, 256, 3)[:2500], rw(200000, 256, 4)
wb.argmin_distance(q, refs, k=1, metric="dtw", metric_params={"r": 0.05})
for _ in range(3):
    c = wb.pinned_copy(refs); pinned = type(getattr(c, "base", None)).__name__ == "_PinnedBlock"; del c
    print("pinned copy of the references page-locked:", pinned, flush=True)
    if pinned: break
    time.sleep(0.75)
del refs, q
time.sleep(1.0)
x5, y5 = rw(2000, 4096, 1), rw(2000, 4096, 2)
xs = np.ascontiguousarray(x5[:250])
for rep in range(3):
    for m in ("msm", "twe"):
        wb.pairwise_distance(xs[:32], y5, metric=m, metric_params={"r": 0.05})
        t0 = time.perf_counter()
        xa = distance.check_array(xs, allow_3d=True, ensure_2d=False, dtype=float); ya = distance.check_array(y5, allow_3d=True, ensure_2d=False, dtype=float)
        t1 = time.perf_counter()
        mm = distance._make_metric(m, {"r": 0.05})
        res = _shim.pairwise(mm.metric_id, mm._params(), xa, ya)
        t2 = time.perf_counter()
        st = wb.last_stats()
        print(m, rep, "validate %.1f ms, C call %.1f ms (device total %.1f, kernels %.1f)" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, st["total_ms"], st["kernel_ms"]), "gc", gc.get_count(), flush=True)
