#!/bin/bash
mkdir -p gpurun_out
{ timeout 900 python scripts/fuzz_argmin.py 3000 21; timeout 900 python scripts/fuzz_knn.py 5; } 2>&1 | tee gpurun_out/r02av.log
