#!/bin/bash
# round 2, call aq: cascade for ddtw / adtw (p >= 0); no seeding with a caller lower_bound; full suite; ddtw / adtw argmin timing
mkdir -p gpurun_out
{
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
cat > /tmp/fam.py <<'PY'
import sys, time
sys.path.insert(0, ".")
import numpy as np
import wildboar_b200 as wb
rw = lambda n, T, s: np.cumsum(np.random.default_rng(s).standard_normal((n, T)), axis=1)
q, refs = rw(20000, 256, 3)[:2500], rw(200000, 256, 4)[:50000]
for metric, p in (("dtw", {"r": 0.05}), ("ddtw", {"r": 0.05}), ("adtw", {"r": 0.05, "p": 1.0})):
    for lb in (True, False):
        for rep in range(2):
            t0 = time.perf_counter()
            i, d = wb.argmin_distance(q, refs, k=1, metric=metric, metric_params=p, return_distance=True, device_lower_bound=lb)
            dt = time.perf_counter() - t0
        st = wb.last_stats()
        print(metric, "cascade" if lb else "plain  ", "2500 x 50000: %.1f ms" % (dt * 1e3), {k: st[k] for k in ("kernel_ms", "pairs", "lb_kim_pruned", "lb_keogh_pruned")}, flush=True)
PY
timeout 900 python /tmp/fam.py
} 2>&1 | tee gpurun_out/r02aq.log
