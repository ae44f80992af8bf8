#!/bin/bash
mkdir -p gpurun_out
{ timeout 1500 python scripts/fuzz_argmin.py 400 11; timeout 600 python scripts/fuzz_argmin.py 150 12; } 2>&1 | tee gpurun_out/r02au.log
