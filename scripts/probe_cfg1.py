"""DEV TOOLING: cfg1 (200 x 200 x 150, dtw r = 0.1) on every engine."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import wildboar_b200 as wb
wb.set_devices([0])
rw = lambda n, T, s: np.cumsum(np.random.default_rng(s).standard_normal((n, T)), axis=1)
x, y = rw(200, 150, 1), rw(200, 150, 2)
ref = None
for eng in (None, "strip", "band", "rowscan", "coop"):
    if eng: os.environ["WILDBOAR_CUDA_ENGINE"] = eng
    else: os.environ.pop("WILDBOAR_CUDA_ENGINE", None)
    best = 1e9
    try:
        for rep in range(30):
            d = wb.pairwise_distance(x, y, metric="dtw", metric_params={"r": 0.1})
            best = min(best, wb.last_stats()["kernel_ms"])
        if ref is None: ref = d
        print(eng, "kernel %.4f ms" % best, "engine", wb.last_stats()["engine"], "equal", bool(np.array_equal(d, ref)))
    except RuntimeError as e:
        print(eng, "n/a", str(e)[-70:])
