#!/bin/bash
mkdir -p gpurun_out
{ timeout 600 python scripts/probe_hostgap.py; } 2>&1 | tee gpurun_out/r02bn.log
