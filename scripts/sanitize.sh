#!/bin/bash
# compute-sanitizer passes over the hot path (run under gpurun): memcheck + racecheck + synccheck on one full registration
# per k-NN engine / neighbour mode.  Output: gpurun_out/sanitizer_*.log
set -u
cat > /tmp/san_driver.py <<'PY'
import sys, numpy as np
sys.path.insert(0, ".")
from fast_gicp_b200.core import Core
d = np.load("tests/golden/pair_0p2.npz")
tgt, src = d["target"][::3].copy(), d["source"][::3].copy()
for method, knn_mode, align_mode, hint, vindex, spec in (("DIRECT27", 0, 1, 0, 0, 1), ("DIRECT7", 0, 0, 1, 0, 1), ("DIRECT1", 1, 1, 0, 1, 0), ("DIRECT_RADIUS", 0, 1, 1, 1, 1),
                                                         ("DIRECT27", 0, 1, 0, 1, 0)):
    c = Core(0)
    c.set_neighbor_search_method(method, 1.5)
    c.set_knn_mode(knn_mode); c.set_align_mode(align_mode); c.set_execution_hint(hint)
    c.set_voxel_index(vindex); c.set_speculation(spec)
    r = c.register(tgt, src)
    c.calculate_source_covariances_rbf(3)
    print(method, "converged", bool(r.converged), "fitness", c.fitness_score(np.eye(4)))
    c.close()
# a far outlier: the voxel bounding box no longer fits the direct-mapped index, the k-NN grid's levels are all coarse there and the deferred (block-cooperative) search takes the queries
c = Core(0)
c.set_neighbor_search_method("DIRECT27")
far = np.vstack([tgt, np.array([[3.0e6, -2.0e6, 1.0e6]], dtype=np.float32)])
r = c.register(far, src)
print("outlier", "converged", bool(r.converged))
c.close()
# DIRECT1 on a cloud above the streaming threshold: k_linearize_stream (bulk copies into shared memory, mbarriers), 1M-class sort path
from fast_gicp_b200.synthetic import kitti_like_pair
bt, bs, _ = kitti_like_pair(beams=48, az_steps=2083, seed=5, pose=(0.4, 0.05, 0.5), downsample=0.0)
c = Core(0)
c.set_resolution(0.5); c.set_neighbor_search_method("DIRECT1")
r = c.register(bt, bs)
print("stream", len(bs), "converged", bool(r.converged))
c.close()
# NDT D2D through the same evaluation kernels
c = Core(0)
c.set_problem(2); c.set_neighbor_search_method("DIRECT7")
c.set_target_cloud(tgt); c.set_source_cloud(src); c.ndt_create_voxelmaps()
r = c.align()
print("ndt d2d", "converged", bool(r.converged))
c.close()
PY
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python /tmp/san_driver.py > gpurun_out/r02_sanitizer_$tool.log 2>&1
  echo "== $tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/r02_sanitizer_$tool.log | tail -1)"
done
