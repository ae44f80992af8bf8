#!/usr/bin/env bash
# Round 2, call F (2 GPUs): the N > 1 path of bench.py (one rank per GPU under torchrun) incl. the in-process multi-GPU
# run of the library (e2e_inprocess) and the per-config evidence with cfg4 / cfg5 sharded over the ranks; plus the one GPU
# test that needs two devices.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "multi_gpu or self_join_mirrored" > gpurun_out/pytest_2gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_2gpu.log; tail -5 gpurun_out/pytest_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench rc=$?"
tail -5 gpurun_out/bench_n2.err
python - <<'PY'
import json
b = json.loads(open("gpurun_out/bench_n2.json").read().strip().splitlines()[-1])
print("value", b["value"], "e2e", b["e2e"]["value"], "parity", b["parity"], "inproc", b.get("e2e_inprocess"))
c = b.get("configs", {})
print("cfg4", c.get("cfg4")); print("cfg5", c.get("cfg5")); print("cpu", b.get("cpu_baseline"))
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_ref_n2.json 2> gpurun_out/bench_ref_n2.err; echo "ref rc=$?"; cat gpurun_out/bench_ref_n2.json
