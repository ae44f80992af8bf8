#!/usr/bin/env bash
# compute-sanitizer (memcheck) over one small invocation of every engine: smoke() (strip / band / row-scan / cooperative / argmin
# replay / scans) plus forced-engine pairwise calls.  Not a bench; the log goes to profiles/.
set -x
mkdir -p gpurun_out
cat > /tmp/san.py <<'PY'
import os, sys
sys.path.insert(0, ".")
import numpy as np
import __graft_entry__ as g
g.smoke()
import wildboar_b200 as wb
rw = lambda n, T, s: np.cumsum(np.random.default_rng(s).standard_normal((n, T)), axis=1)
x, y = rw(12, 150, 1), rw(40, 150, 2)
for eng in ("strip", "band", "rowscan", "coop"):
    os.environ["WILDBOAR_CUDA_ENGINE"] = eng
    for m in ("dtw", "msm", "twe", "erp", "lcss"):
        try:
            wb.pairwise_distance(x, y, metric=m, metric_params={"r": 0.1})
            print(eng, m, "engine", wb.last_stats()["engine"])
        except RuntimeError as e:
            print(eng, m, "n/a:", str(e)[-60:])
os.environ.pop("WILDBOAR_CUDA_ENGINE")
X = rw(300, 40, 3)
wb.pairwise_distance(X, metric="msm", metric_params={"r": 0.3})   # self join mirrored on the device
wb.set_precision("fp64_fma"); wb.pairwise_distance(x, y, metric="dtw", metric_params={"r": 0.1}); wb.set_precision(None)
print("sanitize run complete")
PY
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 3 --log-file gpurun_out/sanitizer_memcheck.log python /tmp/san.py > gpurun_out/sanitizer_stdout.log 2>&1; echo "sanitizer rc=$?" | tee -a gpurun_out/sanitizer_stdout.log
tail -5 gpurun_out/sanitizer_stdout.log; tail -8 gpurun_out/sanitizer_memcheck.log
