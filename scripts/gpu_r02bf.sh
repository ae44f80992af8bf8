#!/bin/bash
# round 2, call bf: chunk width of the UNSEEDED scan (heap-order k > 1, or WILDBOAR_CUDA_NO_SEED): first chunk 128, growing x4 up to C
mkdir -p gpurun_out
{
P0="WILDBOAR_CUDA_PIPED_UPLOAD_KB=0 WILDBOAR_CUDA_NO_SEED=1"
echo "== default (C = 4 Mi / nq = 1664)"; env $P0 timeout 300 python scripts/probe_cfg4.py | tail -1
for c in 3200 6400 12800; do echo "== C = $c, first 128"; env $P0 WILDBOAR_CUDA_ARGMIN_CHUNK=$c WILDBOAR_CUDA_ARGMIN_FIRST=128 timeout 300 python scripts/probe_cfg4.py | tail -1; done
} 2>&1 | tee gpurun_out/r02bf.log
