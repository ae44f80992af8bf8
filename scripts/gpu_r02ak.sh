#!/bin/bash
# round 2, call ak: DP work list appended by the LB pass itself (no count / scan / fill passes); full GPU suite; bench
mkdir -p gpurun_out
{
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
P0="WILDBOAR_CUDA_PIPED_UPLOAD_KB=0"
echo "== default (piped, seeded from the first piece)"; timeout 300 python scripts/probe_cfg4.py | tail -1
echo "== not piped, seeded from all"; env $P0 timeout 300 python scripts/probe_cfg4.py | tail -1
echo "== not piped, no seed"; env $P0 WILDBOAR_CUDA_NO_SEED=1 timeout 300 python scripts/probe_cfg4.py | tail -1
echo "== 1 query"; timeout 300 python scripts/probe_cfg4.py 1 | tail -1
echo "== 64 queries"; timeout 300 python scripts/probe_cfg4.py 64 | tail -1
timeout 300 python scripts/probe_overhead.py | tail -4
env $P0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02ak_launches_cfg4.csv python scripts/probe_cfg4.py > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(l for l in open('gpurun_out/r02ak_launches_cfg4.csv') if l.startswith('"')))
h = rows[0]; ki = h.index('Kernel Name'); vi = h.index('Metric Value')
rows = rows[1:]
half = len(rows) // 2   # two identical calls: take the second
agg = collections.OrderedDict()
for r in rows[half:]:
    n = r[ki].split('(')[0][:60]
    a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += float(r[vi].replace(',', '')) / 1e6
for n, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]): print(f"{ms:9.3f} ms {c:5d}  {n}")
PY
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02ak_bench_n1.json 2> gpurun_out/r02ak_bench_n1.err; tail -c 600 gpurun_out/r02ak_bench_n1.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02ak_bench_n1.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step")}, d["e2e"]["value"], d["parity"]["ok"])
print(json.dumps(d["configs"]["cfg4"], indent=0))
print({k: (v["kernel_gcups"], v["parity"]) for k, v in d["configs"]["cfg5"].items()}, d["configs"]["cfg1"]["kernel_ms"], d["configs"]["cfg1"]["parity"])
PY
} 2>&1 | tee gpurun_out/r02ak.log
