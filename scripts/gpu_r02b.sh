#!/usr/bin/env bash
# Round 2, call B (1 GPU): cooperative engine -- parity, engine comparison, ncu of the long-series kernels.
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "coop or cfg1 or cfg5 or edge or golden" --maxfail=10 > gpurun_out/pytest_coop.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_coop.log
tail -30 gpurun_out/pytest_coop.log
timeout 900 python scripts/bench_engines.py > gpurun_out/engines.jsonl 2> gpurun_out/engines.err; echo "engines rc=$?"
cat gpurun_out/engines.jsonl; tail -5 gpurun_out/engines.err
# ncu: the cooperative msm / twe kernels on a 32-row share of cfg5 (pipes, issue, stalls, dram bytes, L2 hit)
M="smsp__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active,sm__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,gpu__time_duration.sum,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio,smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio,launch__registers_per_thread,launch__grid_size,launch__block_size"
timeout 600 ncu --metrics $M --clock-control none -k regex:k_coop -c 4 --csv --log-file gpurun_out/ncu_coop_cfg5.csv python scripts/bench_engines.py cfg5_32 coop > gpurun_out/ncu_coop.log 2>&1
timeout 600 ncu --metrics $M --clock-control none -k regex:k_strip -c 4 --csv --log-file gpurun_out/ncu_strip_cfg5.csv python scripts/bench_engines.py cfg5_32 strip > gpurun_out/ncu_strip.log 2>&1
timeout 300 ncu --metrics $M --clock-control none -k regex:k_coop -c 2 --csv --log-file gpurun_out/ncu_coop_cfg1.csv python scripts/bench_engines.py cfg1x4 coop > gpurun_out/ncu_coop1.log 2>&1
ls -la gpurun_out
