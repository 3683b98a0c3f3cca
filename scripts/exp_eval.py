"""Micro-benchmark of one LM evaluation (fused lookup + derivatives) on a prepared handle: ms per vgicp_compute_error call."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fast_gicp_b200.core import Core, REG_PLANE

d = np.load("tests/golden/pair_0p1.npz")
tgt, src = d["target"], d["source"]
for method in ("DIRECT27", "DIRECT7", "DIRECT1"):
    for hint in (0, 1):
        c = Core(0)
        c.set_neighbor_search_method(method); c.set_execution_hint(hint)
        c.set_target_cloud(tgt); c.find_target_neighbors(20); c.calculate_target_covariances(REG_PLANE); c.create_target_voxelmap()
        c.set_source_cloud(src); c.find_source_neighbors(20); c.calculate_source_covariances(REG_PLANE)
        T = np.eye(4); T[:3, 3] = [0.4, 0.1, 0.0]
        c.update_correspondences(T)
        out = []
        for want_H in (True, False):
            for _ in range(20): c.compute_error(T, want_H)
            t0 = time.perf_counter()
            for _ in range(300): c.compute_error(T, want_H)
            out.append(1e6 * (time.perf_counter() - t0) / 300)
        print(f"{os.environ.get('VGICP_B200_LIB','default')[-28:]:28s} {method:9s} hint {hint}: linearize {out[0]:6.1f} us  error-only {out[1]:6.1f} us", flush=True)
        c.close()
