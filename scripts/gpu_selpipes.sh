#!/usr/bin/env bash
# One gpurun call: which pipe runs the selects of the fp64 min?  (issue-rate ubench + production strip kernel with
# dmin2 as FSEL+FSEL (m0) vs SEL+FSEL (m1))
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 300 bench/ubench/issue_rates > gpurun_out/issue_rates3.log 2>&1
cat gpurun_out/issue_rates3.log
for v in 0 1; do
  timeout 600 bench/ubench/strip_variants_m$v 1024 10000 512 0.1 2 0 x x > gpurun_out/variants_m${v}_cfg3.log 2>&1
  timeout 600 bench/ubench/strip_variants_m$v 512 5000 140 1.0 2 2 x x > gpurun_out/variants_m${v}_cfg2_metrics.log 2>&1
done
grep -h -E "GCUPS|problem" gpurun_out/variants_m*_cfg3.log
grep -h -E "GCUPS|problem" gpurun_out/variants_m*_cfg2_metrics.log
