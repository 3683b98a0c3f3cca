#!/bin/bash
# round 2, GPU call C: fused cooperative grid build -- kNN parity tests, timing, launch lists at 17k and 1M
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "knn or c4_matches or large_cloud or align_matches" > gpurun_out/r2c_knn_tests.log 2>&1
echo "== knn tests: $(tail -3 gpurun_out/r2c_knn_tests.log)"
grep -E "FAILED|Error|assert " gpurun_out/r2c_knn_tests.log | head -20
timeout 300 python scripts/exp_knn.py 300 2>&1 | tail -3
timeout 300 python scripts/exp_knn_1m.py 10 2>&1 | tail -1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -s 60 -c 60 --csv --log-file gpurun_out/r2c_knn_launches.csv python scripts/exp_knn.py 20 > /dev/null 2>&1
python profiles/summarise_launches.py gpurun_out/r2c_knn_launches.csv 2>/dev/null | head -20
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -s 8 -c 16 --csv --log-file gpurun_out/r2c_knn_launches_1m.csv python scripts/exp_knn_1m.py 4 > /dev/null 2>&1
python profiles/summarise_launches.py gpurun_out/r2c_knn_launches_1m.csv 2>/dev/null | head -20
