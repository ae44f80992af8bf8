#!/usr/bin/env bash
# Round 2, call J (8 GPUs): bench.py at N = 8 as the driver launches it (one rank per GPU), incl. e2e_inprocess (the library's
# own 8-device path, checked bit-equal against the per-rank slabs) and cfg4 / cfg5 sharded over the 8 ranks = the whole configs.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv | head -9
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/bench_n4.json 2> gpurun_out/bench_n4.err; echo "bench rc=$?"
tail -5 gpurun_out/bench_n4.err
python - <<'PY'
import json
b = json.loads(open("gpurun_out/bench_n4.json").read().strip().splitlines()[-1])
print("value", b["value"], "e2e", b["e2e"]["value"], "ms", b["ms_per_step"], b["e2e"]["ms_per_step"], "parity", b["parity"])
print("inproc", b.get("e2e_inprocess"))
c = b.get("configs", {})
print("cfg4", c.get("cfg4")); print("cfg5", c.get("cfg5")); print("cpu", b.get("cpu_baseline")); print("fma", b.get("fp64_fma"))
PY
