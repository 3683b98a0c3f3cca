#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02_gpu_suite.log 2>&1
echo "== gpu suite: $(tail -1 gpurun_out/r02_gpu_suite.log)"
grep -E "FAILED|Error" gpurun_out/r02_gpu_suite.log | head
bash scripts/r2_final_artifacts.sh
