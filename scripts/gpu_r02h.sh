#!/usr/bin/env bash
# Round 2, call H (1 GPU): band-register engine, plain rows vs blocked interior rows, on the scans and on argmin.
set -x
mkdir -p gpurun_out
for B in 0 1; do
  WILDBOAR_CUDA_BAND_BLOCKED=$B timeout 600 python scripts/bench_scan.py --no-ref > gpurun_out/scan_rows_blk$B.jsonl 2> gpurun_out/scan$B.err; echo "scan blk=$B rc=$?"
  python - <<PY
import json
for ln in open("gpurun_out/scan_rows_blk$B.jsonl"):
    r = json.loads(ln); print("blk=$B", r["row"][:28], r["metric"], r.get("e2e_ms"), r.get("kernel_ms"))
PY
  WILDBOAR_CUDA_BAND_BLOCKED=$B timeout 600 python scripts/probe_band_argmin.py 2>&1 | python -c "
import sys, json
for ln in sys.stdin:
    try: r = json.loads(ln)
    except Exception: print(ln.strip()); continue
    print('blk=$B', r['shape'], r['metric'], 'band', r['band_kernel_ms'], 'rowscan', r['rowscan_kernel_ms'], r['equal'])
" | tee gpurun_out/band_argmin_blk$B.txt
done
