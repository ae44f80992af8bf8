"""DEV TOOLING: why is cfg5 twe's wall time in bench.py 70 ms above its kernels?  Replays bench.py's sequence."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import wildboar_b200 as wb
from oracle import oracle as O
wb.set_devices([0])
rw = lambda n, T, s: np.cumsum(np.random.default_rng(s).standard_normal((n, T)), axis=1)
# something big before, like bench.py's headline + cfg4
big = wb.pairwise_distance(rw(3000, 512, 1), rw(10000, 512, 2), metric="dtw", metric_params={"r": 0.1}); del big
x5, y5 = rw(2000, 4096, 1), rw(2000, 4096, 2)
xs = np.ascontiguousarray(x5[:250])
for m in ("msm", "twe", "twe", "msm"):
    wb.pairwise_distance(xs[:8], y5, metric=m, metric_params={"r": 0.05})
    t0 = time.perf_counter()
    res = wb.pairwise_distance(xs, y5, metric=m, metric_params={"r": 0.05})
    dt = time.perf_counter() - t0
    st = wb.last_stats()
    print(m, "wall %.1f ms total_ms %.1f kernel_ms %.1f" % (dt * 1e3, st["total_ms"], st["kernel_ms"]), flush=True)
    rng = np.random.default_rng(5)
    ii, jj = rng.integers(0, 250, 300), rng.integers(0, 2000, 300)
    t0 = time.perf_counter()
    want = np.array([O.pairwise(m, xs[i:i + 1], y5[j:j + 1], r=0.05)[0, 0] for i, j in zip(ii[:40], jj[:40])])
    print("   oracle spot check %.1f s, equal %s" % (time.perf_counter() - t0, bool(np.array_equal(want, res[ii[:40], jj[:40]]))), flush=True)
