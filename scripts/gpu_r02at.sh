#!/usr/bin/env bash
# round 2, call at (8 GPUs): bench.py at N = 8 as the driver launches it; cfg4 / cfg5 sharded over the 8 ranks = the whole configs
mkdir -p gpurun_out
{
nvidia-smi --query-gpu=index,name --format=csv | head -9
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r02at_bench_n8.json 2> gpurun_out/r02at_bench_n8.err; echo "bench rc=$?"
tail -3 gpurun_out/r02at_bench_n8.err
python - <<'PY'
import json
b = json.loads(open("gpurun_out/r02at_bench_n8.json").read().strip().splitlines()[-1])
print("value", b["value"], "e2e", b["e2e"]["value"], "ms", b["ms_per_step"], b["e2e"]["ms_per_step"], "parity", b["parity"]["ok"])
print("inproc", b.get("e2e_inprocess"))
c = b.get("configs", {})
print("cfg4", c.get("cfg4")); print("cfg5", {k: (v["kernel_gcups"], v["e2e_gcups"], v["parity"]) for k, v in c.get("cfg5", {}).items()})
PY
} 2>&1 | tee gpurun_out/r02at.log
