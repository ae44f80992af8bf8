#!/bin/bash
# round 2, call bh: order of the time steps in the LB cascade taken from the references' spread (build_time_order)
mkdir -p gpurun_out
{
timeout 1200 python -m pytest tests -m gpu -q -x -k "argmin or neighbors or knn or cascade or fitted or lb_prune or seeding or ensemble or concurrent" 2>&1 | tail -3
P0="WILDBOAR_CUDA_PIPED_UPLOAD_KB=0"
echo "== resident-like, data-driven order"; env $P0 timeout 300 python scripts/probe_cfg4.py | tail -1
echo "== resident-like, golden-ratio order"; env $P0 WILDBOAR_CUDA_LB_GOLDEN_ORDER=1 timeout 300 python scripts/probe_cfg4.py | tail -1
echo "== default (host refs), data-driven"; timeout 300 python scripts/probe_cfg4.py | tail -1
echo "== default (host refs), golden"; WILDBOAR_CUDA_LB_GOLDEN_ORDER=1 timeout 300 python scripts/probe_cfg4.py | tail -1
echo "== unseeded resident-like, data-driven"; env $P0 WILDBOAR_CUDA_NO_SEED=1 timeout 300 python scripts/probe_cfg4.py | tail -1
echo "== 1 query"; timeout 300 python scripts/probe_cfg4.py 1 | tail -1
timeout 300 python scripts/probe_overhead.py | tail -3
timeout 300 python scripts/fuzz_argmin.py 600 71 | tail -1
timeout 300 python scripts/fuzz_knn.py 13 | tail -2
} 2>&1 | tee gpurun_out/r02bh.log
