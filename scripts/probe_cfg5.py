"""DEV TOOLING: cfg5 share (250 x 2000 x 4096, r = 0.05) three times per metric: is the first full-size call slower than the next?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import wildboar_b200 as wb
wb.set_devices([0])
rw = lambda n, T, s: np.cumsum(np.random.default_rng(s).standard_normal((n, T)), axis=1)
x, y = rw(2000, 4096, 1)[:250], rw(2000, 4096, 2)
for m in ("msm", "twe"):
    wb.pairwise_distance(x[:8], y, metric=m, metric_params={"r": 0.05})
    for rep in range(3):
        t0 = time.perf_counter()
        wb.pairwise_distance(x, y, metric=m, metric_params={"r": 0.05})
        dt = time.perf_counter() - t0
        st = wb.last_stats()
        print(m, rep, "wall %.1f ms" % (dt * 1e3), "total_ms %.1f kernel_ms %.1f" % (st["total_ms"], st["kernel_ms"]), flush=True)
