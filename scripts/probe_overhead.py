"""DEV TOOLING: fixed cost of an argmin call with a large reference set (upload, allocation, bookkeeping)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import wildboar_b200 as wb
from wildboar_b200 import _shim
wb.set_devices([0])
rw = lambda n, T, s: np.cumsum(np.random.default_rng(s).standard_normal((n, T)), axis=1)
q, refs = rw(64, 256, 3), rw(200000, 256, 4)
for rep in range(3):
    t0 = time.perf_counter()
    wb.argmin_distance(q[:1], refs, k=1, metric="dtw", metric_params={"r": 0.05})
    print("argmin 1 query, host buffers", round((time.perf_counter() - t0) * 1e3, 1), "ms", {k: round(wb.last_stats()[k], 1) for k in ("kernel_ms", "total_ms")})
t0 = time.perf_counter(); fit = _shim.FittedSet(refs.reshape(200000, 1, 256), devices=[0]); print("fit", round((time.perf_counter() - t0) * 1e3, 1), "ms")
from wildboar_b200.distance import DtwMetric
m = DtwMetric(r=0.05)
for rep in range(3):
    t0 = time.perf_counter()
    _shim.argmin_fitted(m.metric_id, m._params(), q[:1], fit, 1, use_device_lb=True)
    print("argmin 1 query, resident", round((time.perf_counter() - t0) * 1e3, 1), "ms", {k: round(wb.last_stats()[k], 1) for k in ("kernel_ms", "total_ms")})
