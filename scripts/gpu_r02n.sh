#!/usr/bin/env bash
# bench.py exactly as the driver runs it at N = 1, wall clock around it
set -x
mkdir -p gpurun_out
SECONDS=0
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$? wall=${SECONDS}s"
tail -3 gpurun_out/bench.err
python - <<'PY'
import json
b = json.loads(open("gpurun_out/bench.json").read().strip().splitlines()[-1])
print("value", b["value"], "e2e", b["e2e"]["value"], "ms", b["ms_per_step"], b["e2e"]["ms_per_step"], "parity", b["parity"]["ok"], "frac", b["roofline"]["frac"])
c = b.get("configs", {})
print("cfg1", c.get("cfg1")); print("cfg5", {m: (v["kernel_gcups"], v["e2e_gcups"], v["parity"]) for m, v in c.get("cfg5", {}).items()})
print("cfg4", {k: c["cfg4"][k] for k in ("kernel_ms", "e2e_ms", "parity")})
print("all parity", all(v["parity"] for v in c["cfg2"].values()) and c["cfg1"]["parity"] and c["cfg4"]["parity"] and all(v["parity"] for v in c["cfg5"].values()))
print("cpu", b.get("cpu_baseline")); print("clocks", b.get("clocks"))
PY
