#!/bin/bash
# round 2, GPU call C: Morton-grid kNN + sorted voxel accumulate + RBF -- parity tests, timing, launch lists at 17k and 1M
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2c_gpu_suite.log 2>&1
echo "== gpu suite: $(tail -3 gpurun_out/r2c_gpu_suite.log)"
grep -E "FAILED|Error|assert " gpurun_out/r2c_gpu_suite.log | head -20
timeout 300 python scripts/exp_knn.py 300 2>&1 | tail -3
timeout 300 python scripts/exp_knn_1m.py 10 2>&1 | tail -1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -s 130 -c 130 --csv --log-file gpurun_out/r2c_knn_launches.csv python scripts/exp_knn.py 20 > /dev/null 2>&1
python profiles/summarise_launches.py gpurun_out/r2c_knn_launches.csv 2>/dev/null | head -20
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -s 28 -c 56 --csv --log-file gpurun_out/r2c_knn_launches_1m.csv python scripts/exp_knn_1m.py 4 > /dev/null 2>&1
python profiles/summarise_launches.py gpurun_out/r2c_knn_launches_1m.csv 2>/dev/null | head -20
