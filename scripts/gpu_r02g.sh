#!/usr/bin/env bash
# Round 2, call G (1 GPU): full suite on the tree with the blocked band rows + background page-locking; scan / argmin
# timings of the band engine against round 1 (profiles/r01o_scan_rows.jsonl, r01m_band_vs_rowscan_argmin.jsonl); ncu of k_band.
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --maxfail=20 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 600 python scripts/bench_scan.py --no-ref > gpurun_out/scan_rows.jsonl 2> gpurun_out/scan.err; echo "scan rc=$?"; cut -c1-330 gpurun_out/scan_rows.jsonl; tail -3 gpurun_out/scan.err
timeout 600 python scripts/probe_band_argmin.py > gpurun_out/band_argmin.jsonl 2>&1; cat gpurun_out/band_argmin.jsonl
M="sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active,sm__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,gpu__time_duration.sum,launch__registers_per_thread,l1tex__t_sector_hit_rate.pct"
timeout 600 ncu --metrics $M --clock-control none -k regex:k_band -c 8 --csv --log-file gpurun_out/ncu_band.csv python scripts/probe_band_argmin.py > gpurun_out/ncu_band.log 2>&1
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
b = json.loads(open("gpurun_out/bench.json").read().strip().splitlines()[-1])
print("value", b["value"], "e2e", b["e2e"]["value"], "ms", b["ms_per_step"], b["e2e"]["ms_per_step"], "parity", b["parity"]["ok"], "traffic", b["roofline"]["traffic"])
c = b.get("configs", {})
print("cfg1", c.get("cfg1")); print("cfg5", {m: (v["kernel_gcups"], v["e2e_gcups"], v["parity"]) for m, v in c.get("cfg5", {}).items()})
print("cfg2 e2e/kernel", {k: round(v["e2e_gcups"] / v["kernel_gcups"], 3) for k, v in c.get("cfg2", {}).items()})
PY
