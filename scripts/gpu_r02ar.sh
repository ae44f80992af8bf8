#!/bin/bash
mkdir -p gpurun_out
{ timeout 600 python -m pytest tests -m gpu -q -x -k "caller_lower_bound" 2>&1 | tail -40; } 2>&1 | tee gpurun_out/r02ar.log
