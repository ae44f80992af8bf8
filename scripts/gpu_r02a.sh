#!/usr/bin/env bash
# Round 2, call A (1 GPU): smoke, GPU tests, bench with the per-config evidence, small-call overhead probe.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log; tail -2 gpurun_out/smoke.log
timeout 1200 python -m pytest tests -m gpu -q --maxfail=20 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -40 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -5 gpurun_out/bench.err
python - <<'PY'
import json
try:
    b = json.loads(open("gpurun_out/bench.json").read().strip().splitlines()[-1])
    for k in ("value", "ms_per_step", "e2e", "parity", "fp64_fma", "cpu_baseline"):
        print(k, b.get(k))
    print("roofline frac", b["roofline"]["frac"], b["roofline"]["kernel"])
    c = b.get("configs", {})
    print("cfg1", c.get("cfg1"))
    for k, v in (c.get("cfg2") or {}).items():
        print("cfg2", k, v["kernel_gcups"], v["e2e_gcups"], v["frac"], v["parity"])
    print("cfg4", c.get("cfg4"))
    print("cfg5", c.get("cfg5"))
    if "error" in c: print("CONFIGS ERROR", c["error"])
except Exception as e:
    print("bench parse failed", repr(e))
PY
python - <<'PY' 2>&1 | tee gpurun_out/overhead.log
import time, numpy as np, sys
sys.path.insert(0, '.')
import wildboar_b200 as wb
from wildboar_b200 import _shim
wb.set_devices([0])
rw = lambda n, T, s: np.cumsum(np.random.default_rng(s).standard_normal((n, T)), axis=1)
x, y = rw(200, 150, 1), rw(200, 150, 2)
m = wb.check_metric("dtw")(r=0.1)
for name, fn in (("api", lambda: wb.pairwise_distance(x, y, metric="dtw", metric_params={"r": 0.1})),
                 ("shim", lambda: _shim.pairwise(m.metric_id, m._params(), x, y))):
    for _ in range(20): fn()
    t0 = time.perf_counter()
    for _ in range(200): fn()
    dt = (time.perf_counter() - t0) / 200
    st = wb.last_stats()
    print(f"cfg1 {name}: {dt*1e3:.3f} ms per call; kernel_ms {st['kernel_ms']:.3f} total_ms(device) {st['total_ms']:.3f}")
X = rw(5000, 140, 1)
for met in ("dtw", "msm"):
    for _ in range(2): wb.pairwise_distance(X, metric=met)
    t0 = time.perf_counter(); wb.pairwise_distance(X, metric=met); dt = time.perf_counter() - t0
    st = wb.last_stats()
    print(f"cfg2 singleton {met}: e2e {dt*1e3:.1f} ms kernel {st['kernel_ms']:.1f} ms device total {st['total_ms']:.1f} ms")
PY
