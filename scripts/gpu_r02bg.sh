#!/bin/bash
mkdir -p gpurun_out
{
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
echo "== unseeded, resident-like"; WILDBOAR_CUDA_PIPED_UPLOAD_KB=0 WILDBOAR_CUDA_NO_SEED=1 timeout 300 python scripts/probe_cfg4.py | tail -1
echo "== default"; timeout 300 python scripts/probe_cfg4.py | tail -1
timeout 600 python scripts/probe_knn.py | tail -5
timeout 300 python scripts/fuzz_argmin.py 600 61 | tail -1
timeout 300 python scripts/fuzz_knn.py 9 | tail -3
} 2>&1 | tee gpurun_out/r02bg.log
