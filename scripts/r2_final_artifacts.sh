#!/bin/bash
# round 2: the committed evidence in one gpurun call -- bench N=1, launch list, ncu full (C2 kernels + C4 evaluation) summarised on the
# box, k-NN stage timings, input-preparation timing.  (scripts/sanitize.sh is a separate call.)
set -u
mkdir -p gpurun_out
timeout 700 python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err
tail -3 gpurun_out/r02_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -s 150 -c 400 --csv --log-file gpurun_out/r02_launches.csv \
  python bench.py --streams 1 --steps 2 --warmup 3 --no-cpu-baseline --no-c4 > /dev/null 2>&1
python profiles/summarise_launches.py gpurun_out/r02_launches.csv > gpurun_out/r02_launches_summary.txt 2>/dev/null; head -8 gpurun_out/r02_launches_summary.txt
bash scripts/make_traffic.sh
timeout 600 ncu --set full --clock-control none -k regex:"k_linearize" -s 4 -c 4 -f -o /tmp/r02_c4_eval python scripts/exp_eval_c4.py 2 > gpurun_out/r02_c4_eval.log 2>&1
python profiles/summarise_ncu_full.py /tmp/r02_c4_eval.ncu-rep gpurun_out/r02_c4_eval_ncu_summary.json | tee gpurun_out/r02_c4_eval_ncu_summary.txt
timeout 300 python scripts/exp_knn.py 300 2>&1 | tail -1 | tee gpurun_out/r02_knn_timing.txt
timeout 300 python scripts/exp_knn_1m.py 10 2>&1 | tail -1 | tee -a gpurun_out/r02_knn_timing.txt
timeout 300 python scripts/exp_eval_c4.py 20 2>&1 | tail -2 | tee -a gpurun_out/r02_knn_timing.txt
timeout 120 python - > gpurun_out/r02_prep_timing.log 2>&1 <<'PY'
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
echo "== prep: $(tail -1 gpurun_out/r02_prep_timing.log)"
du -sh gpurun_out
