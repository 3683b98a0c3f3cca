"""Time the kNN stage alone (Morton grid build + k_knn_search + k_knn_deferred) on the fixture target cloud: wall clock over a loop of
find_target_neighbors + synchronize on one stream, and the per-category device times from the profiling hooks."""
import sys, time, numpy as np
sys.path.insert(0, ".")
from fast_gicp_b200.core import Core

d = np.load("tests/golden/pair_0p1.npz")
tgt = d["target"].copy()
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 300
c = Core(0)
c.set_target_cloud(tgt)
for _ in range(20):
    c.find_target_neighbors(20)
c.synchronize()
t0 = time.perf_counter()
for _ in range(reps):
    c.find_target_neighbors(20)
c.synchronize()
dt = (time.perf_counter() - t0) / reps
print(f"kNN stage (k=20, {len(tgt)} pts): {dt * 1e6:.1f} us per call (pipelined launches, one stream)")
if len(sys.argv) > 2:  # 1M synthetic
    from fast_gicp_b200.synthetic import kitti_like_pair
    big, _, _ = kitti_like_pair(beams=128, az_steps=8192, seed=44, pose=(0.5, 0.0, 1.0), downsample=0.0, max_points=int(sys.argv[2]))
    c.set_target_cloud(big)
    for _ in range(3):
        c.find_target_neighbors(20)
    c.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        c.find_target_neighbors(20)
    c.synchronize()
    print(f"kNN stage (k=20, {len(big)} pts): {(time.perf_counter() - t0) / 10 * 1e3:.3f} ms per call")
