#!/bin/bash
# ncu --set full of the k-NN stage kernels (17k fixture cloud)
set -u
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_knn_search|k_knn_deferred|k_sort_pass|k_grid_table" -s 12 -c 6 -f -o gpurun_out/r2d_knn python scripts/exp_knn.py 3 > gpurun_out/r2d_ncu.log 2>&1
tail -2 gpurun_out/r2d_ncu.log
