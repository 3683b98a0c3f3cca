"""C4: 1M-pt synthetic pair, DIRECT27, res 0.5, source sharded over N GPUs with the in-kernel NVLink exchange.
   python scripts/bench_c4_sharded.py N      -> per-evaluation time and agreement with the single-GPU system
   (spawns its own N workers: run it with plain python, not under torchrun; VGICP_C4_METHOD=DIRECT1 selects the bandwidth-bound case)"""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
from test_multi_gpu import run_workers  # noqa: E402

if __name__ == "__main__":
    world = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    res = run_workers(world, args=("c4", "20"), timeout=1500)
    r0 = res[0]
    print(json.dumps({"workload": "C4 synthetic 1M-pt pair, %s res 0.5" % os.environ.get("VGICP_C4_METHOD", "DIRECT27"), "n_gpus": world, "ms_per_evaluation_max_over_ranks": max(r["ms_per_evaluation"] for r in res),
                      "H_rel_diff_vs_single_gpu": r0["H_rel_diff_vs_full"], "iters_sharded_vs_single": r0["iters"], "converged": r0["converged"],
                      "ranks_bit_identical": all(r["H_sum"] == r0["H_sum"] and r["T"] == r0["T"] for r in res), "comm_error": max(r["comm_error"] for r in res)}))
