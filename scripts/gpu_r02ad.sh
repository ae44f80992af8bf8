#!/bin/bash
# round 2, call ad: LB pass with CTA-shared reference blocks + fused survivor counts; first-chunk sweep; argmin tests
mkdir -p gpurun_out
{
timeout 600 python -m pytest tests -m gpu -q -x -k "argmin or neighbors or knn or cascade or fitted" 2>&1 | tail -3
echo "== default"; python scripts/probe_cfg4.py | tail -1
for f in 64 128 256 512; do echo "== first=$f"; WILDBOAR_CUDA_ARGMIN_FIRST=$f python scripts/probe_cfg4.py | tail -1; done
echo "== 1 query"; python scripts/probe_cfg4.py 1 | tail -1
echo "== 64 queries"; python scripts/probe_cfg4.py 64 | tail -1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,lts__t_bytes.sum,sm__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio --clock-control none -k regex:k_lb_prune -s 20 -c 4 --csv --log-file gpurun_out/r02ad_ncu_lb_prune.csv python scripts/probe_cfg4.py > /dev/null 2>&1
grep -c k_lb_prune gpurun_out/r02ad_ncu_lb_prune.csv
} 2>&1 | tee gpurun_out/r02ad.log
