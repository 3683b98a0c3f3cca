#!/bin/bash
# ncu --set full of the k-NN search kernel (17k fixture cloud)
set -u
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_knn_search -s 3 -c 1 -f -o gpurun_out/r2d_knn_search python scripts/exp_knn.py 3 > gpurun_out/r2d_ncu.log 2>&1
tail -3 gpurun_out/r2d_ncu.log
ls -la gpurun_out/*.ncu-rep | tail -2
