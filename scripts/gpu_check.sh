#!/usr/bin/env bash
# One gpurun call: GPU parity tests, bench line (fp64) + fp32-mode evidence.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench.json
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --precision fp32 > gpurun_out/bench_fp32.json 2> gpurun_out/bench_fp32.err; echo "bench rc=$?"
cat gpurun_out/bench_fp32.json
for w in cfg2_msm cfg2_twe cfg2_adtw cfg5_msm cfg5_twe; do
timeout 600 python bench.py --workload $w --steps 2 --warmup 3 --no-cpu-baseline --precision fp32 --e2e-steps 1 > gpurun_out/bench_fp32_$w.json 2>> gpurun_out/bench_fp32.err; cat gpurun_out/bench_fp32_$w.json | cut -c1-300
done
