"""CPU (gloo, world_size 2) test of the N>1 bench path: barrier, max-over-ranks timing, sum of units, partition."""
import os
import socket
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent(
    """
    import importlib.util, json, os, sys
    spec = importlib.util.spec_from_file_location("dist_mod", os.path.join(sys.argv[1], "fast_gicp_b200", "distributed.py"))
    D = importlib.util.module_from_spec(spec); spec.loader.exec_module(D)
    rank, world = D.init("gloo")
    D.barrier()
    ms = [10.0 + 5.0 * rank, 3.0 - rank]
    mx = D.max_over_ranks(ms)
    value, worst = D.aggregate_throughput(local_units=100 * (rank + 1), local_ms=20.0 * (rank + 1))
    lo, hi = D.partition(11, rank, world)
    handles = D._all_gather_bytes(bytes([rank] * 64))  # the transport of the 64-byte IPC handles of the sharded path
    D.barrier()
    if rank == 0:
        print(json.dumps({"mx": mx, "value": value, "worst": worst, "part": [lo, hi], "world": world, "handles": [list(h[:2]) + [len(h)] for h in handles]}))
    else:
        print(json.dumps({"part": [lo, hi]}))
    D.finalize()
    """
)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_replica_plumbing_world2(tmp_path):
    import json

    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    port = _free_port()
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script), ROOT], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    outs = [p.communicate(timeout=120) for p in procs]
    for p, (o, e) in zip(procs, outs):
        assert p.returncode == 0, e
    r0 = json.loads(outs[0][0].strip().splitlines()[-1])
    r1 = json.loads(outs[1][0].strip().splitlines()[-1])
    assert r0["world"] == 2
    assert r0["mx"] == [15.0, 3.0]                 # element-wise max over ranks
    assert abs(r0["value"] - 300 / 0.040) < 1e-6   # (100 + 200) units / slowest rank (40 ms)
    assert r0["worst"] == 40.0
    assert r0["part"] == [0, 6] and r1["part"] == [6, 11]
    assert r0["handles"] == [[0, 0, 64], [1, 1, 64]]  # rank order, 64 bytes each
