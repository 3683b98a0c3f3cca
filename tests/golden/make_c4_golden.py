#!/usr/bin/env python3
"""Generate tests/golden/c4_golden.json: the CPU oracle's results on BASELINE config 4 (synthetic 1M-pt pair, seeds 44/45, res 0.5).

    python tests/golden/make_c4_golden.py          (a few minutes on 8 cores; not run by the tests)

The C4 inputs are too large for the oracle to be re-run inside the GPU test suite, so its outputs are committed once:
  - SHA-256 of the inputs (the generator must reproduce them on the GPU box), of both 1M x 20 k-NN tables (kd-tree, ascending
    (d2, index)) and of the voxel bucket table (coordinates + ids);
  - num_buckets / num_voxels / per-voxel point-count digest;
  - err, H, b of one evaluation at the identity and at the generator's ground-truth pose, DIRECT27 and DIRECT1;
  - the final pose, iteration and evaluation counters of the whole LM registration, DIRECT27 and DIRECT1.
tests/test_gpu_parity.py::test_c4_matches_the_oracle_golden asserts the CUDA path against this file.
"""
import hashlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle as O  # noqa: E402
from fast_gicp_b200.synthetic import kitti_like_pair  # noqa: E402


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    t0 = time.time()
    tgt, src, T_gt = kitti_like_pair(beams=128, az_steps=8192, seed=44, pose=(0.5, 0.0, 1.0), downsample=0.0, max_points=1_000_000)
    out = {"generator": "fast_gicp_b200.synthetic.kitti_like_pair(beams=128, az_steps=8192, seed=44, pose=(0.5, 0.0, 1.0), downsample=0.0, max_points=1000000)",
           "n_target": len(tgt), "n_source": len(src), "sha_target": sha(tgt), "sha_source": sha(src), "T_gt": T_gt.tolist(), "res": 0.5, "k": 20}
    t_nbr = O.knn(tgt, 20, "kdtree")
    s_nbr = O.knn(src, 20, "kdtree")
    out["sha_knn_target"], out["sha_knn_source"] = sha(t_nbr.astype(np.int32)), sha(s_nbr.astype(np.int32))
    print("knn done", time.time() - t0, flush=True)
    t_cov = O.symmetrized(O.regularize(O.covariances(tgt, t_nbr), O.REG_PLANE))
    s_cov = O.symmetrized(O.regularize(O.covariances(src, s_nbr), O.REG_PLANE))
    out["sha_cov_target"], out["sha_cov_source"] = sha(t_cov), sha(s_cov)
    vm = O.VoxelMap(tgt, t_cov, 0.5, accum_double=True)
    out.update(num_buckets=int(vm.num_buckets), num_voxels=int(vm.num_voxels), sha_bucket_coord=sha(vm.bucket_coord), sha_bucket_id=sha(vm.bucket_id), sha_voxel_num_points=sha(vm.vox_n),
               sha_voxel_means=sha(vm.vox_mean))
    print("voxel map done", time.time() - t0, vm.num_buckets, vm.num_voxels, flush=True)
    for name, method in (("DIRECT27", O.DIRECT27), ("DIRECT1", O.DIRECT1)):
        offs = O.offsets(method)
        rec = {}
        for pname, T in (("identity", np.eye(4)), ("gt", T_gt)):
            e, H, b, nc = O.evaluate(vm, src, s_cov, offs, T, T, True)
            rec[pname] = {"err": float(e), "H": H.tolist(), "b": b.tolist(), "n_correspondences": int(nc)}
        r = O.align_f32(vm, src, s_cov, offs)
        rec["align"] = {"T": r.T.tolist(), "H": r.H.tolist(), "iterations": int(r.iterations), "converged": bool(r.converged), "n_linearize": int(r.n_linearize), "n_error": int(r.n_error)}
        out[name] = rec
        print(name, "done", time.time() - t0, r.iterations, r.converged, flush=True)
    with open(os.path.join(ROOT, "tests", "golden", "c4_golden.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("written", time.time() - t0)


if __name__ == "__main__":
    main()
