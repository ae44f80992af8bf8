#!/bin/bash
mkdir -p gpurun_out
{ timeout 900 python -m pytest tests -m gpu -q -x -k "argmin or neighbors or knn or golden" 2>&1 | tail -3; } 2>&1 | tee gpurun_out/r02ap.log
