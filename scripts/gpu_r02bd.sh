#!/bin/bash
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests -m gpu -q -x -k "argmin or seeding or pipelined" 2>&1 | tail -2
for sp in 1 2 4 8 16; do echo "== seed pieces $sp"; WILDBOAR_CUDA_SEED_PIECES=$sp timeout 300 python scripts/probe_cfg4.py | tail -1; done
echo "== 1 query"; timeout 300 python scripts/probe_cfg4.py 1 | tail -1
} 2>&1 | tee gpurun_out/r02bd.log
