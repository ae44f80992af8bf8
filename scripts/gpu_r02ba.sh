#!/bin/bash
# round 2, call ba: three task buffers in the LB tile kernel
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests -m gpu -q -x -k "argmin or neighbors or knn or cascade or fitted or lb_prune or seeding" 2>&1 | tail -3
P0="WILDBOAR_CUDA_PIPED_UPLOAD_KB=0"
echo "== not piped (seeded from all)"; env $P0 timeout 300 python scripts/probe_cfg4.py | tail -1
echo "== not piped, RB=32"; env $P0 WILDBOAR_CUDA_LB_RB=32 timeout 300 python scripts/probe_cfg4.py | tail -1
echo "== not piped, RB=8"; env $P0 WILDBOAR_CUDA_LB_RB=8 timeout 300 python scripts/probe_cfg4.py | tail -1
echo "== default"; timeout 300 python scripts/probe_cfg4.py | tail -1
timeout 300 python scripts/fuzz_argmin.py 300 41 | tail -1
env $P0 timeout 600 ncu --metrics gpu__time_duration.sum,sm__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum --clock-control none -k regex:k_lb_prune_tile -s 16 -c 14 --csv --log-file gpurun_out/r02ba_ncu_lb.csv python scripts/probe_cfg4.py > /dev/null 2>&1
python - <<'PY'
import csv
rows = list(csv.reader(l for l in open('gpurun_out/r02ba_ncu_lb.csv') if l.startswith('"')))
h = rows[0]; mi = h.index('Metric Name'); vi = h.index('Metric Value')
import collections
agg = collections.defaultdict(list)
for r in rows[1:]: agg[r[mi]].append(float(r[vi].replace(',', '')))
for k, v in agg.items(): print(k, "mean %.4g" % (sum(v) / len(v)), "n", len(v))
PY
} 2>&1 | tee gpurun_out/r02ba.log
