#!/usr/bin/env python
"""DEV / evidence tooling: the same pairwise call on different DP engines (WILDBOAR_CUDA_ENGINE), kernel GCUPS from the
library's CUDA events.  One JSON line per (shape, metric, engine)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import wildboar_b200 as wb  # noqa: E402
from wildboar_b200 import _shim  # noqa: E402

OPS = {"dtw": 5, "ddtw": 5, "wdtw": 6, "adtw": 7, "lcss": 4, "erp": 6, "edr": 7, "msm": 8, "twe": 10}


def rw(n, T, seed):
    return np.cumsum(np.random.default_rng(seed).standard_normal((n, T)), axis=1)


SHAPES = {
    "cfg1": (200, 200, 150, 0.1, ["dtw"]),
    "cfg1x4": (400, 400, 150, 0.1, ["dtw", "msm"]),
    "cfg5_250": (250, 2000, 4096, 0.05, ["msm", "twe", "dtw"]),
    "cfg5_32": (32, 2000, 4096, 0.05, ["msm", "twe"]),
    "cfg2_1000": (1000, 5000, 140, 1.0, ["dtw", "msm", "twe"]),
    "cfg3_500": (500, 10000, 512, 0.1, ["dtw"]),
    "t1024_r05": (64, 2000, 1024, 0.05, ["dtw", "msm", "twe"]),
    "dba_like": (1, 2000, 512, 0.1, ["dtw"]),
}


def main():
    which = sys.argv[1].split(",") if len(sys.argv) > 1 else list(SHAPES)
    engines = sys.argv[2].split(",") if len(sys.argv) > 2 else ["auto", "strip", "coop"]
    wb.set_devices([0])
    peak = _shim.fp64_peak(0)[0] / 1e9
    print(json.dumps({"fp64_peak_g_lane_inst_per_s": peak}), flush=True)
    for name in which:
        nx, ny, T, r, metrics = SHAPES[name]
        x, y = rw(nx, T, 1), rw(ny, T, 2)
        for m in metrics:
            ref = None
            for eng in engines:
                if eng == "auto":
                    os.environ.pop("WILDBOAR_CUDA_ENGINE", None)
                else:
                    os.environ["WILDBOAR_CUDA_ENGINE"] = eng
                try:
                    wb.pairwise_distance(x, y, metric=m, metric_params={"r": r})
                    best = None
                    for _ in range(3):
                        t0 = time.perf_counter()
                        out = wb.pairwise_distance(x, y, metric=m, metric_params={"r": r})
                        dt = time.perf_counter() - t0
                        st = wb.last_stats()
                        if best is None or st["kernel_ms"] < best[0]["kernel_ms"]:
                            best = (st, dt)
                    st, dt = best
                    same = None if ref is None else bool(np.array_equal(out, ref))
                    if ref is None:
                        ref = out.copy()
                    g = st["cells"] / (st["kernel_ms"] * 1e-3) / 1e9
                    print(json.dumps(dict(shape=name, metric=m, engine_req=eng, engine=st["engine"], cfg=[st["strip_w"], st["strip_nr"], st["strip_warps"], st["strip_gring"]],
                                          kernel_ms=round(st["kernel_ms"], 3), e2e_ms=round(dt * 1e3, 3), kernel_gcups=round(g, 1),
                                          frac=round(g * OPS[m] / peak, 4), equal_to_first=same)), flush=True)
                except RuntimeError as e:
                    print(json.dumps(dict(shape=name, metric=m, engine_req=eng, error=str(e)[:120])), flush=True)
    os.environ.pop("WILDBOAR_CUDA_ENGINE", None)


if __name__ == "__main__":
    main()
