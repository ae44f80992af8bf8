#!/usr/bin/env python
"""argmin of the non-DTW metrics: band-register engine (default when H <= 32) vs the row-scan engine it replaces."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import wildboar_b200 as wb  # noqa: E402

wb.set_devices([0])
rw = lambda n, T, s: np.cumsum(np.random.default_rng(s).standard_normal((n, T)), axis=1)
for (nq, nr, T, r) in ((400, 400, 128, 0.1), (400, 400, 128, 0.05), (2000, 20000, 128, 0.05), (1000, 5000, 140, 0.1)):
    q, x = rw(nq, T, 3), rw(nr, T, 4)
    for metric in ("msm", "twe", "erp", "lcss", "edr"):
        row = dict(shape=f"{nq}x{nr}x{T} r={r}", metric=metric)
        res = {}
        for eng in ("band", "rowscan"):
            os.environ["WILDBOAR_CUDA_ENGINE"] = eng
            wb.argmin_distance(q, x, k=1, metric=metric, metric_params={"r": r})
            best = 1e9
            for _ in range(3):
                t0 = time.perf_counter()
                res[eng] = wb.argmin_distance(q, x, k=1, metric=metric, metric_params={"r": r}, return_distance=True)
                best = min(best, time.perf_counter() - t0)
            st = wb.last_stats()
            row[eng + "_e2e_ms"] = round(best * 1e3, 2)
            row[eng + "_kernel_ms"] = round(st["kernel_ms"], 2)
            row[eng + "_engine"] = st["engine"]
        row["equal"] = bool(np.array_equal(res["band"][0], res["rowscan"][0]) and np.array_equal(res["band"][1], res["rowscan"][1]))
        print(json.dumps(row), flush=True)
