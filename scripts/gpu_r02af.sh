#!/bin/bash
# round 2, call af: register-tiled LB pass (Q queries per warp task), Q sweep
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests -m gpu -q -x -k "argmin or neighbors or knn or cascade or fitted" 2>&1 | tail -3
for q in 1 2 4 8; do echo "== Q=$q"; WILDBOAR_CUDA_LB_Q=$q python scripts/probe_cfg4.py | tail -1; done
echo "== default, resident timing (no piped upload)"; WILDBOAR_CUDA_PIPED_UPLOAD_KB=0 python scripts/probe_cfg4.py | tail -1
echo "== 1 query"; python scripts/probe_cfg4.py 1 | tail -1
echo "== 64 queries"; python scripts/probe_cfg4.py 64 | tail -1
for q in 4 8; do
WILDBOAR_CUDA_LB_Q=$q ncu --metrics gpu__time_duration.sum,l1tex__t_sector_hit_rate.pct,lts__t_bytes.sum,sm__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:k_lb_prune -s 20 -c 3 --csv --log-file gpurun_out/r02af_ncu_lb_prune_q$q.csv python scripts/probe_cfg4.py > /dev/null 2>&1
grep "k_lb_prune" gpurun_out/r02af_ncu_lb_prune_q$q.csv | awk -F'","' '{print $13" | "$15}' | sort | uniq | awk 'NR%3==1'
done
} 2>&1 | tee gpurun_out/r02af.log
