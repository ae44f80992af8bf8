#!/usr/bin/env bash
# One gpurun call: GPU parity tests (+ smoke), LB-matrix timing.
set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log; tail -2 gpurun_out/smoke.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
python - <<'PY' 2>&1 | tee gpurun_out/lb_timing.log
import time, numpy as np, sys
sys.path.insert(0, '.')
import wildboar_b200 as wb
from wildboar_b200.lb import DtwKeoghLowerBound, DtwKimLowerBound
wb.set_devices([0])
rw = lambda n, T, s: np.cumsum(np.random.default_rng(s).standard_normal((n, T)), axis=1)
q, refs = rw(2000, 256, 3), rw(20000, 256, 4)
for kind in ("both", "left"):
    est = DtwKeoghLowerBound(r=0.05, kind=kind).fit(refs)
    est.transform(q[:64])
    t0 = time.perf_counter(); lb = est.transform(q); dt = time.perf_counter() - t0
    st = wb.last_stats()
    print(f"lb_keogh kind={kind} 2000x20000xT256: e2e {dt*1e3:.1f} ms, device kernels+D2H {st['kernel_ms']:.1f} ms, "
          f"{st['cells']/ (st['kernel_ms']*1e-3)/1e9:.1f} G pair-steps/s")
est = DtwKimLowerBound().fit(refs); est.transform(q[:64])
t0 = time.perf_counter(); lb = est.transform(q); dt = time.perf_counter() - t0
print(f"lb_kim 2000x20000: e2e {dt*1e3:.1f} ms, device {wb.last_stats()['kernel_ms']:.1f} ms")
PY
