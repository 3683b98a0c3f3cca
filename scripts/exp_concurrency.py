"""Experiment: throughput of T concurrent registration streams (one Core/handle per host thread) on one GPU."""
import sys, os, time, threading
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fast_gicp_b200.core import Core, REG_PLANE

d = np.load("tests/golden/pair_0p1.npz")
tgt, src = d["target"], d["source"]
dev = torch.device("cuda", 0)
tgt_d = torch.from_numpy(tgt).to(dev).contiguous(); src_d = torch.from_numpy(src).to(dev).contiguous()
torch.cuda.synchronize()

ONECALL = len(sys.argv) > 1 and sys.argv[1] == "onecall"
def worker(core, n, out, i):
    if ONECALL:
        for _ in range(n):
            r = core.register_raw(tgt_d.data_ptr(), len(tgt), src_d.data_ptr(), len(src), 12, True)
        out[i] = r.nr_iterations
        return
    for _ in range(n):
        core.set_cloud_device("target", tgt_d.data_ptr(), len(tgt), 12)
        core.find_target_neighbors(20); core.calculate_target_covariances(REG_PLANE); core.create_target_voxelmap()
        core.set_cloud_device("source", src_d.data_ptr(), len(src), 12)
        core.find_source_neighbors(20); core.calculate_source_covariances(REG_PLANE)
        r = core.align()
    out[i] = r.nr_iterations

for T in [int(x) for x in os.environ.get("TS", "1,2,4,8,16,32").split(",")]:
    cores = [Core(0) for _ in range(T)]
    for c in cores:
        c.set_neighbor_search_method("DIRECT27")
        c.set_execution_hint(int(os.environ.get("HINT", "0")))
        c.set_speculation(int(os.environ.get("SPEC", "1")))
    out = [None] * T
    for c in cores: worker(c, 3, out, 0)
    torch.cuda.synchronize()
    per = max(200 // T, 10)
    th = [threading.Thread(target=worker, args=(cores[i], per, out, i)) for i in range(T)]
    t0 = time.perf_counter()
    for t in th: t.start()
    for t in th: t.join()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"T={T:3d}  {T*per/dt:9.1f} reg/s   ({1e3*dt/(per):.3f} ms per registration per stream)", flush=True)
    for c in cores: c.close()
