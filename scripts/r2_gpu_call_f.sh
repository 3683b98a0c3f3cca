#!/bin/bash
# round 2, GPU call F (2 GPUs): sharded evaluation + stage-1 sharding test; C4 evaluation timing
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -q -x > gpurun_out/r2f_multi_gpu_test.log 2>&1
echo "== multi-gpu test: $(tail -3 gpurun_out/r2f_multi_gpu_test.log)"
grep -E "FAILED|Error|assert |rror" gpurun_out/r2f_multi_gpu_test.log | head -20
timeout 300 python scripts/exp_eval_c4.py 20 2>&1 | tail -2
timeout 300 python scripts/exp_eval.py 2>&1 | tail -5
