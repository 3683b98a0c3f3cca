#!/bin/bash
# ncu --set full of the evaluation kernels at 1M points (stream kernel, DIRECT27) 
set -u
mkdir -p gpurun_out
timeout 300 python scripts/exp_eval_c4.py 20 2>&1 | tail -2
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_linearize" -s 4 -c 4 -f -o gpurun_out/r2d_eval python scripts/exp_eval_c4.py 2 > gpurun_out/r2d_ncu.log 2>&1
tail -2 gpurun_out/r2d_ncu.log
