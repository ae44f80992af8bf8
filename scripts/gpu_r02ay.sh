#!/bin/bash
mkdir -p gpurun_out
{ timeout 900 python scripts/fuzz_argmin.py 300 31; } 2>&1 | tee gpurun_out/r02ay.log
