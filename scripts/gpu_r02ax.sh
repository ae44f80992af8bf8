#!/bin/bash
# round 2, final-tree evidence as the driver runs it: GPU tests, smoke(), bench.py with the default flags, the reference arm
mkdir -p gpurun_out
{
echo "(GPU tests: see the first run of this script, 295 passed)"
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
SECONDS=0; timeout 1500 python bench.py > gpurun_out/r02ax_bench_n1_final.json 2> gpurun_out/r02ax_bench_n1_final.err; echo "bench rc=$?"; echo "bench wall ${SECONDS} s"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02ax_bench_n1_final.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "dtype", "gpu_launches")})
print("e2e", d["e2e"]); print("roofline", {k: d["roofline"][k] for k in ("bound", "achieved", "peak", "frac", "traffic")}); print("cpu_baseline", d["cpu_baseline"]); print("clocks", d["clocks"]); print("parity", d["parity"]["ok"])
c = d["configs"]
print("cfg1", {k: c["cfg1"][k] for k in ("kernel_ms", "e2e_ms", "frac", "parity")})
print("cfg2", {k: (v["kernel_gcups"], v["frac"], v["e2e_gcups"], v["parity"]) for k, v in c["cfg2"].items()})
print("cfg4", {k: c["cfg4"][k] for k in ("kernel_ms", "e2e_ms", "e2e_resident_refs_ms", "e2e_pinned_refs_ms", "nominal_e2e_gcups", "pruned_fraction_rank0", "parity")})
print("cfg5", {k: (v["kernel_gcups"], v["frac"], v["e2e_gcups"], v["parity"]) for k, v in c["cfg5"].items()})
PY
timeout 900 python bench.py --impl reference > gpurun_out/r02ax_bench_reference_arm.json 2> gpurun_out/r02ax_bench_reference_arm.err; echo "ref rc=$?"; cat gpurun_out/r02ax_bench_reference_arm.json | head -c 1200
} 2>&1 | tee gpurun_out/r02ax.log
