#!/bin/bash
# round 2, GPU call A: baseline of the round-1 state (pending tests as hard tests, whole suite, C4 DIRECT1/27, bench c2/c2_direct1/c4)
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_input_prep.py tests/test_batch.py "tests/test_gpu_parity.py::test_fitness_score_against_kdtree" -m gpu --runxfail -q > gpurun_out/r2_pending_tests.log 2>&1
echo "== pending tests: $(tail -1 gpurun_out/r2_pending_tests.log)"
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2_gpu_suite.log 2>&1
echo "== gpu suite: $(tail -1 gpurun_out/r2_gpu_suite.log)"
VGICP_C4_METHOD=DIRECT1 timeout 300 python scripts/bench_c4_sharded.py 1 > gpurun_out/r2_c4_direct1.log 2>&1
echo "== c4 direct1: $(tail -1 gpurun_out/r2_c4_direct1.log | cut -c1-300)"
timeout 400 python bench.py > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err
echo "== bench c2: $(cut -c1-400 gpurun_out/r2_bench.json)"
timeout 300 python bench.py --workload c2_direct1 --no-cpu-baseline > gpurun_out/r2_bench_d1.json 2> gpurun_out/r2_bench_d1.err
echo "== bench c2_direct1: $(cut -c1-300 gpurun_out/r2_bench_d1.json)"
timeout 400 python bench.py --workload c4 --streams 2 --steps 6 --no-cpu-baseline > gpurun_out/r2_bench_c4.json 2> gpurun_out/r2_bench_c4.err
echo "== bench c4: $(cut -c1-300 gpurun_out/r2_bench_c4.json)"
timeout 400 python bench.py --workload c4_direct1 --streams 2 --steps 6 --no-cpu-baseline > gpurun_out/r2_bench_c4d1.json 2> gpurun_out/r2_bench_c4d1.err
echo "== bench c4_direct1: $(cut -c1-300 gpurun_out/r2_bench_c4d1.json)"
