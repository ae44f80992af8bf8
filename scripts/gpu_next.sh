#!/usr/bin/env bash
# One gpurun call: full GPU parity suite, the next-row measurements, the headline bench, launch list of the next-row kernels.
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 900 python scripts/bench_next.py > gpurun_out/bench_next.jsonl 2> gpurun_out/bench_next.err; echo "rc=$?"
cat gpurun_out/bench_next.jsonl; tail -5 gpurun_out/bench_next.err
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "rc=$?"
cat gpurun_out/bench_n1.json | cut -c1-400
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_next_quick.csv \
  python scripts/bench_next.py --quick --no-ref > gpurun_out/ncu_next_quick.log 2>&1
grep -c k_ gpurun_out/launches_next_quick.csv
