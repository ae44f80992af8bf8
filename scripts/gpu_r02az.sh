#!/bin/bash
# round 2, call az: one `ncu --set full` capture of the final LB pass kernel (cfg4 share, references resident-like) and of the survivor DP
mkdir -p gpurun_out
{
WILDBOAR_CUDA_PIPED_UPLOAD_KB=0 timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_lb_prune_tile -s 5 -c 1 -o gpurun_out/r02az_lb_tile_full -f python scripts/probe_cfg4.py > /dev/null 2>&1
ls -la gpurun_out/r02az_lb_tile_full.ncu-rep
ncu -i gpurun_out/r02az_lb_tile_full.ncu-rep --page details --csv > gpurun_out/r02az_lb_tile_full_details.csv 2>/dev/null
wc -l gpurun_out/r02az_lb_tile_full_details.csv
} 2>&1 | tee gpurun_out/r02az.log
