"""C4 evaluation kernels alone: 1M-point pair prepared once, then repeated linearize() calls (DIRECT1 = streaming kernel, DIRECT27)."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from fast_gicp_b200.core import Core
from fast_gicp_b200.synthetic import kitti_like_pair

tgt, src, _ = kitti_like_pair(beams=128, az_steps=8192, seed=44, pose=(0.5, 0.0, 1.0), downsample=0.0, max_points=1_000_000)
c = Core(0)
c.set_resolution(0.5)
c.set_target_cloud(tgt); c.find_target_neighbors(20); c.calculate_target_covariances(3); c.create_target_voxelmap()
c.set_source_cloud(src); c.find_source_neighbors(20); c.calculate_source_covariances(3)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
for method in ("DIRECT1", "DIRECT27"):
    c.set_neighbor_search_method(method)
    T = np.eye(4)
    for _ in range(3):
        c.linearize(T)
    t0 = time.perf_counter()
    for _ in range(reps):
        c.linearize(T)
    print(method, "linearize: %.1f us per call (host-driven)" % ((time.perf_counter() - t0) / reps * 1e6))
