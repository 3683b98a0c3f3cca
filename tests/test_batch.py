"""Batch registration (include/vgicp_batch_b200.h): a pool of handles + worker threads over the public C ABI; every pair is one
vgicp_register on one handle, so the results must equal the sequential loop's bit for bit whatever the interleaving.

CPU: the library loads and exports exactly what its header declares.  GPU: batch vs sequential, in a child process."""
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_batch_library_exports_its_abi():
    import ctypes

    import build_native

    build_native.build_native()
    path = build_native.build_batch()
    from fast_gicp_b200 import batch, core

    core.load_library()
    lib = ctypes.CDLL(path)
    hdr = open(os.path.join(ROOT, "include", "vgicp_batch_b200.h")).read()
    declared = set(re.findall(r"VGICP_API\s+[\w\s\*]+?\b(vgicp_batch_\w+)\s*\(", hdr))
    assert declared == set(batch.EXPORTED_SYMBOLS)
    for sym in declared:
        assert hasattr(lib, sym)
    assert lib.vgicp_batch_num_streams(None) == 0


_CHILD = r"""
import sys, numpy as np
sys.path.insert(0, %r)
from fast_gicp_b200.batch import BatchRegistration
from fast_gicp_b200.core import Core, pose_from_c
d = np.load(%r)
tgt0, src0 = d["target"], d["source"]
rng = np.random.default_rng(9)
tgts, srcs = [], []
for i in range(24):
    yaw = rng.uniform(-0.05, 0.05)
    R = np.array([[np.cos(yaw), -np.sin(yaw), 0], [np.sin(yaw), np.cos(yaw), 0], [0, 0, 1]], dtype=np.float32)
    t = rng.uniform(-0.5, 0.5, 3).astype(np.float32)
    tgts.append((tgt0 @ R.T + t).astype(np.float32)); srcs.append((src0[: len(src0) - i] @ R.T + t).astype(np.float32))
b = BatchRegistration(0, 6)
b.configure(1.0, "DIRECT27")
poses, res, aligned = b.register(tgts, srcs, want_aligned=True)
c = Core(0); c.set_neighbor_search_method("DIRECT27")
for i in range(24):
    r = c.register(tgts[i], srcs[i])
    assert np.array_equal(np.array(r.T), np.array(res[i].T)), i
    assert (r.nr_iterations, r.converged, r.n_linearize, r.n_compute_error) == (res[i].nr_iterations, res[i].converged, res[i].n_linearize, res[i].n_compute_error)
    assert np.array_equal(c.transform_source(pose_from_c(r.T)), aligned[i]), i
assert all(r.converged for r in res)
print("batch ok")
"""


@pytest.mark.gpu
def test_gpu_batch_equals_the_sequential_loop():
    code = _CHILD % (ROOT, os.path.join(ROOT, "tests", "golden", "pair_0p2.npz"))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=180)
    assert r.returncode == 0 and "batch ok" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
