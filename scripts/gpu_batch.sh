#!/usr/bin/env bash
# One gpurun call: GPU tests, bench + full-size DRAM traffic after the task-order change, per-config table.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
cat gpurun_out/bench.json
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,lts__t_bytes.sum --clock-control none -k regex:k_strip -s 3 -c 1 --csv --log-file gpurun_out/traffic_full.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_traffic.log 2>&1
tail -6 gpurun_out/traffic_full.csv | cut -c150-400
timeout 900 python scripts/bench_configs.py > gpurun_out/bench_configs.log 2>&1; echo "bench_configs rc=$?"
cat gpurun_out/bench_configs.log
