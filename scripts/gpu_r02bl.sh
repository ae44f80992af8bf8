#!/bin/bash
# round 2, call bl: staged upload for wb_cuda_fit and the replicated operand of pairwise calls; full suite
mkdir -p gpurun_out
{
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
timeout 300 python scripts/probe_overhead.py | tail -4
timeout 300 python scripts/probe_cfg5.py | tail -3
echo "== default"; timeout 300 python scripts/probe_cfg4.py | tail -1
timeout 300 python scripts/fuzz_argmin.py 300 91 | tail -1
} 2>&1 | tee gpurun_out/r02bl.log
