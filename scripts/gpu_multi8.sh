#!/usr/bin/env bash
# 8-GPU refresh: multi-GPU parity, the driver's bench at N=8, full configs through the library's own sharding.
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.sm --format=csv > gpurun_out/gpus_n$N.txt
timeout 300 python -m pytest tests -m gpu -x -q -k "multi_gpu or fitted_set or multivariate_large or subsequence_larger" 2>&1 | tail -2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
cat gpurun_out/bench_n$N.json | cut -c1-300
timeout 900 python scripts/bench_multi.py --all-only 2>&1 | tee gpurun_out/bench_multi.log | cut -c1-330
