#!/usr/bin/env bash
# Round 2, call M (1 GPU): final tree -- smoke, full GPU suite, bench as the driver runs it (--steps 20 --warmup 5), launch list.
set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log; tail -2 gpurun_out/smoke.log
timeout 1200 python -m pytest tests -m gpu -q --maxfail=20 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"

python - <<'PY'
import json
b = json.loads(open("gpurun_out/bench.json").read().strip().splitlines()[-1])
print("value", b["value"], "e2e", b["e2e"]["value"], "ms", b["ms_per_step"], b["e2e"]["ms_per_step"], "parity", b["parity"]["ok"], "frac", b["roofline"]["frac"])
c = b.get("configs", {})
print("cfg1", c.get("cfg1")); print("cfg5", {m: (v["kernel_gcups"], v["e2e_gcups"], v["parity"]) for m, v in c.get("cfg5", {}).items()})
print("cfg4", {k: c["cfg4"][k] for k in ("kernel_ms", "e2e_ms", "parity")})
print("all parity", all(v["parity"] for v in c["cfg2"].values()) and c["cfg1"]["parity"] and c["cfg4"]["parity"] and all(v["parity"] for v in c["cfg5"].values()))
print("cpu", b.get("cpu_baseline"))
PY
timeout 600 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2>/dev/null; cat gpurun_out/bench_ref.json | cut -c1-300
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs --profile-rows 1024 > gpurun_out/launches.log 2>&1
python - <<'PY'
import csv, io, collections
txt = open("gpurun_out/launches_bench.csv").read(); i = txt.index('"ID"')
tot = collections.Counter(); n = collections.Counter()
for r in csv.DictReader(io.StringIO(txt[i:])):
    k = r["Kernel Name"][:60]; tot[k] += float(r["Metric Value"]); n[k] += 1
s = sum(tot.values())
for k, v in tot.most_common(8): print(f"{k:62s} launches {n[k]:4d}  time {v/1e6:9.3f} ms  share {100*v/s:5.1f} %")
PY
