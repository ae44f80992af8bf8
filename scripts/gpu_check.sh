#!/usr/bin/env bash
# One gpurun call: GPU parity tests, per-config table, bench line.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 900 python scripts/bench_configs.py > gpurun_out/bench_configs.log 2>&1; echo "bench_configs rc=$?"
cat gpurun_out/bench_configs.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench.json
