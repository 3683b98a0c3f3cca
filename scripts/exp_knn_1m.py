"""kNN stage alone on the C4 target cloud (1M points): per-call time, or under ncu for the launch list."""
import sys, time
sys.path.insert(0, ".")
from fast_gicp_b200.core import Core
from fast_gicp_b200.synthetic import kitti_like_pair

big, _, _ = kitti_like_pair(beams=128, az_steps=8192, seed=44, pose=(0.5, 0.0, 1.0), downsample=0.0, max_points=1_000_000)
c = Core(0)
c.set_target_cloud(big)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
for _ in range(2):
    c.find_target_neighbors(20)
c.synchronize()
t0 = time.perf_counter()
for _ in range(reps):
    c.find_target_neighbors(20)
c.synchronize()
print(f"kNN stage (k=20, {len(big)} pts): {(time.perf_counter() - t0) / reps * 1e3:.3f} ms per call")
