"""DEV TOOLING: one cfg4-shaped argmin call (2500 queries x 200k refs x 256, dtw r=0.05, device LB cascade)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import wildboar_b200 as wb
wb.set_devices([0])
rw = lambda n, T, s: np.cumsum(np.random.default_rng(s).standard_normal((n, T)), axis=1)
nq = int(sys.argv[1]) if len(sys.argv) > 1 else 2500
q, refs = rw(20000, 256, 3)[:nq], rw(200000, 256, 4)
for rep in range(2):
    t0 = time.perf_counter()
    idx, dist = wb.argmin_distance(q, refs, k=1, metric="dtw", metric_params={"r": 0.05}, return_distance=True)
    print("argmin", round((time.perf_counter() - t0) * 1e3, 1), "ms", {k: wb.last_stats()[k] for k in ("kernel_ms", "total_ms", "launches", "cells", "lb_kim_pruned", "lb_keogh_pruned")})
