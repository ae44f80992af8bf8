import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import wildboar_b200 as wb
from wildboar_b200 import dtw, _shim
X = np.cumsum(np.random.default_rng(7).standard_normal((2000, 512)), axis=1)
orig = _shim.dba_epoch
def timed_epoch(*a, **k):
    t0 = time.perf_counter(); out = orig(*a, **k); print("  dba_epoch", round((time.perf_counter() - t0) * 1e3, 2), "ms", wb.last_stats()["kernel_ms"]); return out
_shim.dba_epoch = timed_epoch
of = _shim.FittedSet
class TF(of):
    def __init__(self, *a, **k):
        t0 = time.perf_counter(); super().__init__(*a, **k); print("  fit", round((time.perf_counter() - t0) * 1e3, 2), "ms")
    def close(self):
        t0 = time.perf_counter(); super().close(); print("  close", round((time.perf_counter() - t0) * 1e3, 2), "ms")
    __del__ = close
_shim.FittedSet = TF
for rep in range(2):
    t0 = time.perf_counter()
    mean, cost = dtw.dtw_average(X, r=0.1, init=X[0], max_epoch=5, tol=0.0, return_cost=True)
    print("dtw_average", round((time.perf_counter() - t0) * 1e3, 2), "ms")
