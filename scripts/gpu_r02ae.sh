#!/bin/bash
# round 2, call ae: software-pipelined LB loop, first chunk 128, pipelined reference upload; launch list of one cfg4 share
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests -m gpu -q -x -k "argmin or neighbors or knn or cascade or fitted" 2>&1 | tail -3
echo "== default"; python scripts/probe_cfg4.py | tail -2
echo "== no piped upload"; WILDBOAR_CUDA_PIPED_UPLOAD_KB=0 python scripts/probe_cfg4.py | tail -1
echo "== lb warps 16"; WILDBOAR_CUDA_LB_WARPS=16 python scripts/probe_cfg4.py | tail -1
echo "== lb warps 4"; WILDBOAR_CUDA_LB_WARPS=4 python scripts/probe_cfg4.py | tail -1
echo "== 1 query"; python scripts/probe_cfg4.py 1 | tail -1
echo "== 64 queries"; python scripts/probe_cfg4.py 64 | tail -1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02ae_launches_cfg4.csv python scripts/probe_cfg4.py > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(l for l in open('gpurun_out/r02ae_launches_cfg4.csv') if l.startswith('"')))
h = rows[0]; ki = h.index('Kernel Name'); vi = h.index('Metric Value')
rows = rows[1:]
half = len(rows) // 2   # two identical calls: take the second
agg = collections.OrderedDict()
for r in rows[half:]:
    n = r[ki].split('(')[0][:60]
    a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += float(r[vi].replace(',', '')) / 1e6
for n, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]): print(f"{ms:9.3f} ms {c:5d}  {n}")
PY
} 2>&1 | tee gpurun_out/r02ae.log
