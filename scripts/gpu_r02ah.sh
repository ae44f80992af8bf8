#!/bin/bash
# round 2, call ah: threshold seeding (k = 1) + barrier-free register-tiled LB pass; variants
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests -m gpu -q -x -k "argmin or neighbors or knn or cascade or fitted or lb_prune or seeding" 2>&1 | tail -3
P0="WILDBOAR_CUDA_PIPED_UPLOAD_KB=0"
echo "== default (seeded from the first piece, piped)"; timeout 300 python scripts/probe_cfg4.py | tail -1
echo "== seeded from all refs, not piped"; env $P0 timeout 300 python scripts/probe_cfg4.py | tail -1
echo "== no seed, not piped"; env $P0 WILDBOAR_CUDA_NO_SEED=1 timeout 300 python scripts/probe_cfg4.py | tail -1
echo "== no seed: MINB=2"; env $P0 WILDBOAR_CUDA_NO_SEED=1 WILDBOAR_CUDA_LB_MINB=2 timeout 300 python scripts/probe_cfg4.py | tail -1
echo "== no seed: BS=8"; env $P0 WILDBOAR_CUDA_NO_SEED=1 WILDBOAR_CUDA_LB_BS=8 timeout 300 python scripts/probe_cfg4.py | tail -1
echo "== no seed: Q=8"; env $P0 WILDBOAR_CUDA_NO_SEED=1 WILDBOAR_CUDA_LB_Q=8 timeout 300 python scripts/probe_cfg4.py | tail -1
echo "== no seed: Q=0"; env $P0 WILDBOAR_CUDA_NO_SEED=1 WILDBOAR_CUDA_LB_Q=0 timeout 300 python scripts/probe_cfg4.py | tail -1
for c in 3200 6400; do echo "== seeded, chunk $c"; env $P0 WILDBOAR_CUDA_ARGMIN_CHUNK=$c timeout 300 python scripts/probe_cfg4.py | tail -1; done
echo "== seeded Q=8"; env $P0 WILDBOAR_CUDA_LB_Q=8 timeout 300 python scripts/probe_cfg4.py | tail -1
echo "== 1 query"; timeout 300 python scripts/probe_cfg4.py 1 | tail -1
echo "== 64 queries"; timeout 300 python scripts/probe_cfg4.py 64 | tail -1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02ah_launches_cfg4.csv python scripts/probe_cfg4.py > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(l for l in open('gpurun_out/r02ah_launches_cfg4.csv') if l.startswith('"')))
h = rows[0]; ki = h.index('Kernel Name'); vi = h.index('Metric Value')
rows = rows[1:]
half = len(rows) // 2   # two identical calls: take the second
agg = collections.OrderedDict()
for r in rows[half:]:
    n = r[ki].split('(')[0][:60]
    a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += float(r[vi].replace(',', '')) / 1e6
for n, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]): print(f"{ms:9.3f} ms {c:5d}  {n}")
PY
timeout 600 ncu --metrics gpu__time_duration.sum,lts__t_bytes.sum,sm__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio,smsp__average_warps_issue_stalled_membar_per_issue_active.ratio --clock-control none -k regex:k_lb_prune -s 20 -c 3 --csv --log-file gpurun_out/r02ah_ncu_lb_prune.csv env WILDBOAR_CUDA_NO_SEED=1 python scripts/probe_cfg4.py > /dev/null 2>&1
grep "k_lb_prune" gpurun_out/r02ah_ncu_lb_prune.csv | awk -F'","' '{print $13" | "$15}' | sort | uniq | awk 'NR%3==1'
} 2>&1 | tee gpurun_out/r02ah.log
