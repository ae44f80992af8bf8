#!/bin/bash
# round 2, call bi: the restored build (golden-ratio order): argmin tests, probes, wider fuzz
mkdir -p gpurun_out
{
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
echo "== resident-like"; WILDBOAR_CUDA_PIPED_UPLOAD_KB=0 timeout 300 python scripts/probe_cfg4.py | tail -1
echo "== default"; timeout 300 python scripts/probe_cfg4.py | tail -1
timeout 600 python scripts/fuzz_argmin.py 2500 81 | tail -1
timeout 300 python scripts/fuzz_knn.py 17 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" | tail -1 | cut -c1-80
} 2>&1 | tee gpurun_out/r02bi.log
