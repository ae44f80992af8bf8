#!/bin/bash
mkdir -p gpurun_out
{ timeout 300 python scripts/probe_cfg1.py; } 2>&1 | tee gpurun_out/r02aw.log
