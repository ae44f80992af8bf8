#!/usr/bin/env bash
# One gpurun call: GPU tests, A/B of the L2 persisting window (bench + full-size DRAM traffic), fp32 bench.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
for P in 1 0; do
  WILDBOAR_CUDA_L2_PERSIST=$P timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_persist$P.json 2> gpurun_out/bench_persist$P.err
  cut -c1-200 gpurun_out/bench_persist$P.json
  WILDBOAR_CUDA_L2_PERSIST=$P timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:k_strip -s 3 -c 1 --csv --log-file gpurun_out/traffic_persist$P.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_traffic$P.log 2>&1
  tail -4 gpurun_out/traffic_persist$P.csv | cut -c150-400
done
for w in cfg5_twe cfg5_msm; do
  for P in 1 0; do
  WILDBOAR_CUDA_L2_PERSIST=$P timeout 600 python bench.py --workload $w --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 1 --profile-rows 250 > gpurun_out/bench_${w}_persist$P.json 2>> gpurun_out/bench.err
  python -c "import json;d=json.load(open('gpurun_out/bench_${w}_persist$P.json'));print('$w persist=$P', d['value'], d['roofline']['kernel'])"
  done
done
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --precision fp32 > gpurun_out/bench_fp32.json 2> gpurun_out/bench_fp32.err; cut -c1-200 gpurun_out/bench_fp32.json
