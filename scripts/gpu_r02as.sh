#!/bin/bash
mkdir -p gpurun_out
{ timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -4; } 2>&1 | tee gpurun_out/r02as.log
