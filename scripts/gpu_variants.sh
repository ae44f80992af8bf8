#!/usr/bin/env bash
# One gpurun call: GPU parity tests, strip-variant sweeps on the BASELINE shapes, ncu captures of the non-DTW kernels.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
V=bench/ubench/strip_variants
timeout 600 $V 1024 10000 512 0.1 2 0 x > gpurun_out/variants_cfg3.log 2>&1
timeout 300 $V 1024 5000 140 1.0 2 0 x > gpurun_out/variants_cfg2.log 2>&1
timeout 300 $V 2048 10000 256 0.05 2 1 x > gpurun_out/variants_cfg4.log 2>&1
timeout 300 $V 2048 4000 150 0.1 2 1 x > gpurun_out/variants_cfg1.log 2>&1
timeout 900 $V 512 5000 140 1.0 2 2 x > gpurun_out/variants_cfg2_metrics.log 2>&1
timeout 900 $V 32 2000 4096 0.05 1 3 x > gpurun_out/variants_cfg5.log 2>&1
grep -h -E "GCUPS|problem" gpurun_out/variants_cfg*.log
for w in cfg5_msm cfg5_twe; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_strip -s 3 -c 1 -f -o gpurun_out/prof_$w \
    python bench.py --workload $w --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 1 --profile-rows 32 > gpurun_out/ncu_$w.log 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_strip -s 3 -c 1 -f -o gpurun_out/prof_cfg2_wdtw \
    python bench.py --workload cfg2_wdtw --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 1 --profile-rows 512 > gpurun_out/ncu_cfg2_wdtw.log 2>&1
ls -la gpurun_out
