#!/bin/bash
# round 2, GPU call B: the new parity tests (RBF bit-exact, C4 golden, NDT idempotent maps) + whole suite
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2b_gpu_suite.log 2>&1
echo "== gpu suite: $(tail -3 gpurun_out/r2b_gpu_suite.log)"
grep -E "FAILED|Error|assert" gpurun_out/r2b_gpu_suite.log | head -20
timeout 120 python - <<'PY' 2>&1 | tail -5
import numpy as np, time, sys
sys.path.insert(0, ".")
from fast_gicp_b200.core import Core
import torch
d = np.load("tests/golden/pair_0p1.npz")
c = Core(0)
c.set_source_cloud(d["source"])
for _ in range(3): c.calculate_source_covariances_rbf(3)
c.synchronize()
t0 = time.perf_counter()
for _ in range(10): c.calculate_source_covariances_rbf(3)
c.synchronize()
print("rbf covariances (17k pts, PLANE): %.3f ms" % ((time.perf_counter() - t0) / 10 * 1e3))
PY
