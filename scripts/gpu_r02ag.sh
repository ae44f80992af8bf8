#!/bin/bash
# round 2, call ag: LB pass as a register tile with TMA-staged query tiles (k_lb_prune_tile<Q>); Q / task-size sweep
mkdir -p gpurun_out
{
timeout 600 python -m pytest tests -m gpu -q -x -k "argmin or neighbors or knn or cascade or fitted" 2>&1 | tail -3
for q in 0 2 4 8; do echo "== Q=$q"; WILDBOAR_CUDA_PIPED_UPLOAD_KB=0 WILDBOAR_CUDA_LB_Q=$q timeout 300 python scripts/probe_cfg4.py | tail -1; done
for rb in 8 32; do echo "== Q=4 rb=$rb"; WILDBOAR_CUDA_PIPED_UPLOAD_KB=0 WILDBOAR_CUDA_LB_RB=$rb timeout 300 python scripts/probe_cfg4.py | tail -1; done
echo "== default (piped)"; timeout 300 python scripts/probe_cfg4.py | tail -1
echo "== 1 query"; timeout 300 python scripts/probe_cfg4.py 1 | tail -1
echo "== 64 queries"; timeout 300 python scripts/probe_cfg4.py 64 | tail -1
for q in 4; do
WILDBOAR_CUDA_LB_Q=$q timeout 600 ncu --metrics gpu__time_duration.sum,l1tex__t_sector_hit_rate.pct,lts__t_bytes.sum,sm__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active --clock-control none -k regex:k_lb_prune -s 20 -c 3 --csv --log-file gpurun_out/r02ag_ncu_lb_prune_q$q.csv python scripts/probe_cfg4.py > /dev/null 2>&1
grep "k_lb_prune" gpurun_out/r02ag_ncu_lb_prune_q$q.csv | awk -F'","' '{print $13" | "$15}' | sort | uniq | awk 'NR%3==1'
done
} 2>&1 | tee gpurun_out/r02ag.log
