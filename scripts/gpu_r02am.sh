#!/usr/bin/env bash
# round 2, call am (2 GPUs): the two-device GPU test, in-process two-device argmin through the pipelined upload (both worker
# threads upload all references), bench.py under torchrun with 2 ranks
mkdir -p gpurun_out
{
nvidia-smi --query-gpu=index,name --format=csv
timeout 900 python -m pytest tests -m gpu -q -k "multi_gpu or self_join_mirrored or pipelined or seeding" 2>&1 | tail -3
cat > /tmp/two.py <<'PY'
import sys, time
sys.path.insert(0, ".")
import numpy as np
import wildboar_b200 as wb
rw = lambda n, T, s: np.cumsum(np.random.default_rng(s).standard_normal((n, T)), axis=1)
q, refs = rw(20000, 256, 3)[:5000], rw(200000, 256, 4)
res = {}
for devs in ([0], [0, 1]):
    wb.set_devices(devs)
    for rep in range(2):
        t0 = time.perf_counter()
        idx, dist = wb.argmin_distance(q, refs, k=1, metric="dtw", metric_params={"r": 0.05}, return_distance=True)
        dt = time.perf_counter() - t0
    res[len(devs)] = (idx, dist)
    print("devices", devs, "5000 queries x 200000 refs: %.1f ms" % (dt * 1e3), {k: wb.last_stats()[k] for k in ("kernel_ms", "total_ms", "launches")})
print("two devices == one device:", bool(np.array_equal(res[1][0], res[2][0]) and np.array_equal(res[1][1], res[2][1])))
PY
timeout 600 python /tmp/two.py
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r02am_bench_n2.json 2> gpurun_out/r02am_bench_n2.err; echo "bench rc=$?"
tail -3 gpurun_out/r02am_bench_n2.err
python - <<'PY'
import json
b = json.loads(open("gpurun_out/r02am_bench_n2.json").read().strip().splitlines()[-1])
print("value", b["value"], "e2e", b["e2e"]["value"], "parity", b["parity"]["ok"], "inproc", b.get("e2e_inprocess"))
c = b.get("configs", {})
print("cfg4", c.get("cfg4")); print("cfg5", {k: (v["kernel_gcups"], v["parity"]) for k, v in c.get("cfg5", {}).items()})
PY
} 2>&1 | tee gpurun_out/r02am.log
