#!/bin/bash
# Everything round 1 left unverified on hardware, in one gpurun call (about 6 GPU-minutes):
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash scripts/round2_first_gpu_call.sh'
# Output: gpurun_out/r2_*.log
set -u
mkdir -p gpurun_out
# 1. the tests that have never run on a GPU, as hard tests (--runxfail ignores the xfail marks)
timeout 600 python -m pytest tests/test_input_prep.py tests/test_batch.py "tests/test_gpu_parity.py::test_fitness_score_against_kdtree" -m gpu --runxfail -q > gpurun_out/r2_pending_tests.log 2>&1
echo "== pending tests: $(tail -1 gpurun_out/r2_pending_tests.log)"
# 2. the whole GPU suite
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/r2_gpu_suite.log 2>&1
echo "== gpu suite: $(tail -1 gpurun_out/r2_gpu_suite.log)"
# 3. compute-sanitizer over the round-1 final kernels (memcheck, racecheck, synccheck)
bash scripts/sanitize.sh 2>&1 | tail -4
# 4. the bandwidth-bound configuration of the evaluation kernel: 1M points, DIRECT1
VGICP_C4_METHOD=DIRECT1 timeout 300 python scripts/bench_c4_sharded.py 1 > gpurun_out/r2_c4_direct1.log 2>&1
echo "== c4 direct1: $(tail -1 gpurun_out/r2_c4_direct1.log | cut -c1-200)"
timeout 300 python scripts/bench_c4_sharded.py 1 > gpurun_out/r2_c4_direct27.log 2>&1
echo "== c4 direct27: $(tail -1 gpurun_out/r2_c4_direct27.log | cut -c1-200)"
# 5. input preparation timing (raw-like 70k-point scan)
timeout 120 python - > gpurun_out/r2_prep_timing.log 2>&1 <<'PY'
import sys, time, numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from test_input_prep import raw_like_cloud
from fast_gicp_b200.prep import InputPrep
import oracle as O
p = InputPrep(0)
c = raw_like_cloud(1, 70000)
for _ in range(3): p.approximate_voxel_grid(c, 0.1, True)
t0 = time.perf_counter()
for _ in range(20): out = p.approximate_voxel_grid(c, 0.1, True)
gpu = (time.perf_counter() - t0) / 20
t0 = time.perf_counter()
for _ in range(5): ref = O.approximate_voxel_grid(O.remove_near_origin(c), 0.1)
cpu = (time.perf_counter() - t0) / 5
print(f"approximate_voxel_grid(0.1) + origin filter, 70k points: device {gpu*1e3:.3f} ms (host buffers in and out), CPU oracle {cpu*1e3:.3f} ms, equal {np.array_equal(out, ref)}")
PY
echo "== prep: $(tail -1 gpurun_out/r2_prep_timing.log)"
# 6. the default bench (sanity: same numbers as profiles/r01_c_bench.json)
timeout 400 python bench.py > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err
python -c "
import json; d = json.load(open('gpurun_out/r2_bench.json')); print('== bench: value', round(d['value']), 'e2e', round(d['e2e']['value']), 'single-stream ms', round(d['single_stream']['ms_per_registration'], 3))"
# 7. the end-to-end arm through vgicp_batch_register (one C call per timed region) next to the default class-based arm
timeout 400 python bench.py --e2e-impl batch --no-cpu-baseline > gpurun_out/r2_bench_batch.json 2> gpurun_out/r2_bench_batch.err
python -c "
import json; d = json.load(open('gpurun_out/r2_bench_batch.json')); print('== bench --e2e-impl batch: value', round(d['value']), 'e2e', round(d['e2e']['value']), d['pose_check'])"

