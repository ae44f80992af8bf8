#!/usr/bin/env python
"""Multi-GPU through the library's own device sharding (one process, one host thread per device):
full BASELINE configs[2], [3], [4] on all visible GPUs via the public API (host buffers)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import wildboar_b200 as wb  # noqa: E402


def rw(n, T, seed):
    return np.cumsum(np.random.default_rng(seed).standard_normal((n, T)), axis=1)


def timed(fn, reps=2):
    fn()
    best = 1e30
    for _ in range(reps):
        t0 = time.perf_counter(); out = fn(); best = min(best, time.perf_counter() - t0)
    return best, out


def main():
    n = wb.device_count()
    rows = []
    all_only = "--all-only" in sys.argv
    for devs in (([0], list(range(n))) if not all_only else (list(range(n)),)) if n > 1 else ([0],):
        wb.set_devices(devs)
        g = len(devs)
        x, y = rw(10000, 512, 1), rw(10000, 512, 2)
        dt, out = timed(lambda: wb.pairwise_distance(x, y, metric="dtw", metric_params={"r": 0.1}), 2)
        st = wb.last_stats()
        rows.append(dict(config="cfg3 full 10000x512 dtw r=0.1", gpus=g, e2e_s=round(dt, 3), e2e_gcups=round(st["cells"] / dt / 1e9, 1),
                         kernel_ms_max=round(st["kernel_ms"], 1), checksum=float(out[::997, ::991].sum())))
        print(json.dumps(rows[-1]), flush=True)
        if g == 1 and n > 1:
            continue  # the long configs only on all GPUs
        x, y = rw(2000, 4096, 1), rw(2000, 4096, 2)
        for m in ("msm", "twe"):
            dt, out = timed(lambda m=m: wb.pairwise_distance(x, y, metric=m, metric_params={"r": 0.05}), 1)
            st = wb.last_stats()
            rows.append(dict(config=f"cfg5 full 2000x4096 {m} r=0.05", gpus=g, e2e_s=round(dt, 3), e2e_gcups=round(st["cells"] / dt / 1e9, 1),
                             kernel_ms_max=round(st["kernel_ms"], 1), checksum=float(out[::97, ::91].sum())))
            print(json.dumps(rows[-1]), flush=True)
        q, refs = rw(20000, 256, 3), rw(200000, 256, 4)
        for lb in (True, False):
            dt, (idx, dist) = timed(lambda lb=lb: wb.argmin_distance(q, refs, k=1, metric="dtw", metric_params={"r": 0.05},
                                                                  return_distance=True, device_lower_bound=lb), 1)
            st = wb.last_stats()
            nominal = 20000 * 200000 * 5756
            rows.append(dict(config=f"cfg4 full argmin k=1 20000 q x 200000 refs x 256 dtw r=0.05 device_lb={lb}", gpus=g, e2e_s=round(dt, 3),
                             nominal_gcups=round(nominal / dt / 1e9, 1), dp_pairs=st["pairs"], pruned_kim=st["lb_kim_pruned"],
                             pruned_keogh=st["lb_keogh_pruned"], pruning_rate=round(1 - st["pairs"] / 4e9, 4),
                             kernel_ms_max=round(st["kernel_ms"], 1), idx_checksum=int(idx.sum()), dist_checksum=float(dist.sum())))
            print(json.dumps(rows[-1]), flush=True)
    json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "bench_multi.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
