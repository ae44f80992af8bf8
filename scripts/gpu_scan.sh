#!/usr/bin/env bash
# One gpurun call: the new subsequence-scan parity tests first, then the whole GPU suite, the scan measurements, the headline bench.
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_subsequence_scan.py -m gpu -q > gpurun_out/pytest_scan.log 2>&1; echo "pytest scan rc=$?" >> gpurun_out/pytest_scan.log
tail -40 gpurun_out/pytest_scan.log
timeout 400 python scripts/bench_scan.py > gpurun_out/bench_scan.jsonl 2> gpurun_out/bench_scan.err; echo "bench_scan rc=$?"
cat gpurun_out/bench_scan.jsonl; tail -5 gpurun_out/bench_scan.err
timeout 1200 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
cut -c1-600 gpurun_out/bench_n1.json
