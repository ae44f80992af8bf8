#!/bin/bash
# round 2, call al: candidate search split over reference slices (few queries); memcheck / racecheck of the new cascade kernels
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests -m gpu -q -x -k "argmin or neighbors or knn or cascade or fitted or lb_prune or seeding" 2>&1 | tail -3
echo "== default"; timeout 300 python scripts/probe_cfg4.py | tail -1
echo "== 1 query"; timeout 300 python scripts/probe_cfg4.py 1 | tail -1
echo "== 64 queries"; timeout 300 python scripts/probe_cfg4.py 64 | tail -1
timeout 300 python scripts/probe_overhead.py | tail -4
cat > /tmp/san2.py <<'PY'
import os, sys
sys.path.insert(0, ".")
import numpy as np
import wildboar_b200 as wb
rw = lambda n, T, s: np.cumsum(np.random.default_rng(s).standard_normal((n, T)), axis=1)
os.environ["WILDBOAR_CUDA_SEED_MIN"] = "256"
os.environ["WILDBOAR_CUDA_ARGMIN_CHUNK"] = "160"
q, refs = rw(37, 131, 81), rw(1003, 131, 82)
for env in ({}, {"WILDBOAR_CUDA_LB_Q": "0"}, {"WILDBOAR_CUDA_LB_Q": "2"}, {"WILDBOAR_CUDA_LB_Q": "8", "WILDBOAR_CUDA_LB_RB": "3"}, {"WILDBOAR_CUDA_LB_BS": "8"},
            {"WILDBOAR_CUDA_PIPED_UPLOAD_KB": "100"}, {"WILDBOAR_CUDA_NO_SEED": "1"}):
    os.environ.update(env)
    for k in (1, 3):
        i, d = wb.argmin_distance(q, refs, k=k, metric="dtw", metric_params={"r": 0.1}, return_distance=True)
    print(env, wb.last_stats()["lb_keogh_pruned"])
    for key in env: os.environ.pop(key)
print("sanitize run complete")
PY
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 --log-file gpurun_out/r02al_sanitizer_memcheck.log python /tmp/san2.py > gpurun_out/r02al_sanitizer_stdout.log 2>&1; echo "memcheck rc=$?"
tail -3 gpurun_out/r02al_sanitizer_stdout.log; tail -4 gpurun_out/r02al_sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 3 --log-file gpurun_out/r02al_sanitizer_racecheck.log python /tmp/san2.py > gpurun_out/r02al_sanitizer_race_stdout.log 2>&1; echo "racecheck rc=$?"
tail -2 gpurun_out/r02al_sanitizer_race_stdout.log; tail -6 gpurun_out/r02al_sanitizer_racecheck.log
} 2>&1 | tee gpurun_out/r02al.log
