#!/bin/bash
# round 2, call bc: pageable references staged through page-locked buffers by several copy threads
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests -m gpu -q -x -k "argmin or neighbors or knn or cascade or fitted or lb_prune or seeding or concurrent" 2>&1 | tail -3
echo "== default (staged, 4 copy threads)"; timeout 300 python scripts/probe_cfg4.py | tail -2
for t in 1 2 8; do echo "== copy threads $t"; WILDBOAR_CUDA_COPY_THREADS=$t timeout 300 python scripts/probe_cfg4.py | tail -1; done
echo "== no staging (driver's pageable path)"; WILDBOAR_CUDA_NO_STAGING=1 timeout 300 python scripts/probe_cfg4.py | tail -1
for kb in 4096 65536; do echo "== piece $kb KB"; WILDBOAR_CUDA_PIPED_UPLOAD_KB=$kb timeout 300 python scripts/probe_cfg4.py | tail -1; done
echo "== 1 query"; timeout 300 python scripts/probe_cfg4.py 1 | tail -1
echo "== 64 queries"; timeout 300 python scripts/probe_cfg4.py 64 | tail -1
timeout 300 python scripts/fuzz_argmin.py 400 51 | tail -1
} 2>&1 | tee gpurun_out/r02bc.log
