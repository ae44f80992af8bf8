#!/bin/bash
# round 2, call aj: stragglers of the LB pass continued lane-per-pair inside the CTA task; unrolled replay / list fill; tiled
# envelope kernel; wider chunks when seeded
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests -m gpu -q -x -k "argmin or neighbors or knn or cascade or fitted or lb_prune or seeding" 2>&1 | tail -3
P0="WILDBOAR_CUDA_PIPED_UPLOAD_KB=0"
echo "== default (piped, seeded from the first piece)"; timeout 300 python scripts/probe_cfg4.py | tail -2
echo "== resident-like (not piped, seeded from all)"; env $P0 timeout 300 python scripts/probe_cfg4.py | tail -1
for c in 1600 3200 12800; do echo "== not piped, chunk $c"; env $P0 WILDBOAR_CUDA_ARGMIN_CHUNK=$c timeout 300 python scripts/probe_cfg4.py | tail -1; done
for sg in "1,63" "2,31" "4,63" "2,127"; do echo "== not piped, stragglers $sg"; env $P0 WILDBOAR_CUDA_LB_STRAG=$sg timeout 300 python scripts/probe_cfg4.py | tail -1; done
echo "== not piped, no seed"; env $P0 WILDBOAR_CUDA_NO_SEED=1 timeout 300 python scripts/probe_cfg4.py | tail -1
echo "== not piped, plain envelope kernel"; env $P0 WILDBOAR_CUDA_ENVELOPE_PLAIN=1 timeout 300 python scripts/probe_cfg4.py | tail -1
echo "== 1 query"; timeout 300 python scripts/probe_cfg4.py 1 | tail -1
echo "== 64 queries"; timeout 300 python scripts/probe_cfg4.py 64 | tail -1
env $P0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02aj_launches_cfg4.csv python scripts/probe_cfg4.py > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(l for l in open('gpurun_out/r02aj_launches_cfg4.csv') if l.startswith('"')))
h = rows[0]; ki = h.index('Kernel Name'); vi = h.index('Metric Value')
rows = rows[1:]
half = len(rows) // 2   # two identical calls: take the second
agg = collections.OrderedDict()
for r in rows[half:]:
    n = r[ki].split('(')[0][:60]
    a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += float(r[vi].replace(',', '')) / 1e6
for n, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]): print(f"{ms:9.3f} ms {c:5d}  {n}")
PY
} 2>&1 | tee gpurun_out/r02aj.log
