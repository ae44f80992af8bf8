#!/usr/bin/env bash
# N-GPU session: multi-GPU parity test, torchrun bench at N, N/2, N/4, full configs via the library's own sharding.
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.sm --format=csv > gpurun_out/gpus_n$N.txt
python -m pytest tests -m gpu -x -q -k "multi_gpu" 2>&1 | tail -2
n=$N
while [ $n -ge 2 ]; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 3 --warmup 3 > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err
  cat gpurun_out/bench_n$n.json
  n=$((n/2))
done
timeout 900 python scripts/bench_multi.py 2>&1 | tee gpurun_out/bench_multi.log
