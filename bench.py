#!/usr/bin/env python3
"""bench.py -- registrations/sec of the VGICP hot path on B200 (BASELINE.json metric), one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c2_direct1|c3|c4]

Workload (N=1 default, BASELINE configs[1]): the reference's benchmark pair (tests/golden/pair_0p1.npz = data/251370668.pcd
vs 251371071.pcd after align.cpp's filter + ApproximateVoxelGrid(0.1): 17047 / 17334 points), FastVGICPCuda, DIRECT27,
voxel_res 1.0, k=20, PLANE, LM defaults, identity initial guess.

One *step* = one registration under the reference's "100times" protocol (src/align.cpp:72-81): clearTarget, clearSource,
setInputTarget (upload, kNN, covariances, voxel map), setInputSource (upload, kNN, covariances), align.

  value  : registrations/s with both clouds already resident in HBM when the timed region starts (device pointers
           through the C ABI); CUDA events on the handle's stream around every step, L2 flushed between steps.
  e2e    : the same through the reference-facing class FastVGICPCuda with HOST (pinned) buffers: H2D of both clouds and D2H
           of the aligned cloud + pose inside the timed region.
  roofline: dominant kernel of the step, algorithmic bytes / CUDA-event time, against MEASURED_PEAKS.json.
  cpu_baseline: the reference's own CPU implementation of the path (OpenMP FastVGICP, restated in oracle/ because the
           reference cannot be compiled here) on the box's host cores, bounded sample.

N>1 (torchrun): the 17k-pt path does not shard usefully -> replicas, one registration stream per GPU, no
data-path collective; value = N*K registrations / max-over-ranks time ("weak" scaling).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "registrations/sec (VGICP, ~17k-pt pairs)"
UNIT = "registrations/s"


# ------------------------------------------------------------------------------------------------------------ inputs
def load_workload(name):
    if name in ("c2", "c2_direct1"):
        d = np.load(os.path.join(ROOT, "tests", "golden", "pair_0p1.npz"))
        return dict(
            name="C2: FastVGICPCuda %s res=1.0 k=20 PLANE LM, 17k-pt fixture pair (17047/17334 pts)" % ("DIRECT27" if name == "c2" else "DIRECT1"),
            target=np.ascontiguousarray(d["target"], dtype=np.float32), source=np.ascontiguousarray(d["source"], dtype=np.float32),
            method="DIRECT27" if name == "c2" else "DIRECT1", res=1.0,
            data="fixture: reference data/251370668.pcd vs 251371071.pcd, align.cpp filter + ApproximateVoxelGrid(0.1)")
    from fast_gicp_b200.synthetic import kitti_like_pair

    if name == "c3":
        t, s, _ = kitti_like_pair(beams=64, az_steps=2083, seed=42, pose=(0.8, 0.05, 0.7), downsample=0.25)
        return dict(name="C3: synthetic HDL-64 pair (0.25 m downsample) DIRECT27 res=1.0", target=t, source=s, method="DIRECT27", res=1.0, data="synthetic")
    if name == "c4":
        t, s, _ = kitti_like_pair(beams=128, az_steps=8192, seed=44, pose=(0.5, 0.0, 1.0), downsample=0.0, max_points=1_000_000)
        return dict(name="C4: synthetic 1M-pt pair DIRECT27 res=0.5", target=t, source=s, method="DIRECT27", res=0.5, data="synthetic")
    raise SystemExit("unknown workload " + name)


# --------------------------------------------------------------------------------------------------- clocks sampling
class ClockSampler:
    FIELDS = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.path = tempfile.mktemp(prefix="clocks_", suffix=".csv")
        self.proc = None
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            p = [x.strip() for x in line.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0]))
                mx.append(float(p[1]))
            except ValueError:
                continue
            for nme, v in zip(names, p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        os.unlink(self.path)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))
        return out


# -------------------------------------------------------------------------------------------------------- CPU arm
def cpu_registration_loop(w, min_seconds, min_regs, max_regs):
    """The reference's CPU implementation of this path (FastVGICP, OpenMP) restated in oracle/: per registration
    calculate_covariances(target), calculate_covariances(source) (kd-tree kNN, k=20, PLANE), voxel map, LM align."""
    import oracle as O

    offs = O.offsets(getattr(O, w["method"]))
    tgt, src = w["target"], w["source"]
    times = []
    t_end = time.perf_counter() + min_seconds
    T = None
    while (len(times) < min_regs or time.perf_counter() < t_end) and len(times) < max_regs:
        t0 = time.perf_counter()
        tc = O.covariances_f64(tgt, 20, O.REG_PLANE)
        sc = O.covariances_f64(src, 20, O.REG_PLANE)
        r = O.align_f64(tgt, tc, src, sc, res=w["res"], offs=offs)
        times.append(time.perf_counter() - t0)
        T = r.T
    return times, T, O.num_threads()


def run_reference_arm(args, w, rank, world):
    """--impl reference: times the CPU implementation on the host cores (rank 0 only)."""
    if rank != 0:
        return
    for _ in range(args.warmup):
        cpu_registration_loop(w, 0.0, 1, 1)
    times, _, threads = cpu_registration_loop(w, 0.0, args.steps, args.steps)
    total = float(np.sum(times))
    value = len(times) / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": len(times), "warmup": args.warmup,
        "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": w["data"],
        "config": {"workload": w["name"], "protocol": "align.cpp 100times (covariances recomputed every registration)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": "%d full registrations, restated OpenMP FastVGICP (reference not buildable here: no Eigen/PCL)" % len(times)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# -------------------------------------------------------------------------------------------------------- GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    w = load_workload(args.workload)

    if args.impl == "reference":
        run_reference_arm(args, w, rank, world)
        return

    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from fast_gicp_b200 import FastVGICPCuda
    from fast_gicp_b200.core import REG_PLANE, Core, pose_from_c

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    tgt, src = w["target"], w["source"]
    n_t, n_s = len(tgt), len(src)
    K, W = args.steps, args.warmup

    # device-resident inputs for `value`, pinned host inputs for `e2e`
    tgt_d = torch.from_numpy(tgt).to(dev).contiguous()
    src_d = torch.from_numpy(src).to(dev).contiguous()
    tgt_h = torch.from_numpy(tgt).clone().pin_memory()
    src_h = torch.from_numpy(src).clone().pin_memory()
    aligned_h = torch.empty((n_s, 3), dtype=torch.float32).pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    torch.cuda.synchronize()

    core = Core(local_rank)
    core.set_resolution(w["res"])
    core.set_neighbor_search_method(w["method"])
    stream = torch.cuda.ExternalStream(core.stream(), device=dev)

    def step_resident():
        core.set_cloud_device("target", tgt_d.data_ptr(), n_t, 12)
        core.find_target_neighbors(20)
        core.calculate_target_covariances(REG_PLANE)
        core.create_target_voxelmap()
        core.set_cloud_device("source", src_d.data_ptr(), n_s, 12)
        core.find_source_neighbors(20)
        core.calculate_source_covariances(REG_PLANE)
        return core.align()

    reg = FastVGICPCuda(local_rank)
    reg.setResolution(w["res"])
    reg.voxel_resolution_ = w["res"]
    reg.setNeighborSearchMethod(w["method"])
    tgt_np, src_np, aligned_np = tgt_h.numpy(), src_h.numpy(), aligned_h.numpy()

    def step_e2e():
        reg.clearTarget()
        reg.clearSource()
        reg.setInputTarget(tgt_np)
        reg.setInputSource(src_np)
        return reg.align(aligned_out=aligned_np)

    def timed(step_fn, steps):
        evs = []
        for _ in range(steps):
            flush.zero_()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            out = step_fn()
            b.record(stream)
            evs.append((a, b))
        torch.cuda.synchronize()
        return [a.elapsed_time(b) for a, b in evs], out

    # ---- value: resident inputs
    for _ in range(W):
        step_resident()
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    l0 = core.launch_count()
    t_wall0 = time.perf_counter()
    ms, res = timed(step_resident, K)
    barrier()
    wall_s = time.perf_counter() - t_wall0
    launches = core.launch_count() - l0
    total_ms = float(np.sum(ms))

    # ---- e2e: host buffers through the reference-facing class
    for _ in range(W):
        step_e2e()
    barrier()
    l1 = reg.vgicp_cuda_.launch_count()
    ms_e2e, T_e2e = timed(step_e2e, K)
    barrier()
    launches_e2e = reg.vgicp_cuda_.launch_count() - l1
    total_ms_e2e = float(np.sum(ms_e2e))
    clocks = sampler.stop() if sampler else None

    # ---- max over ranks
    if world > 1:
        t = torch.tensor([total_ms, total_ms_e2e], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, total_ms_e2e = float(t[0]), float(t[1])

    # ---- per-kernel profile (separate pass, events around every launch) -> roofline of the dominant kernel
    core.set_profiling(True)
    for _ in range(min(K, 20)):
        flush.zero_()
        torch.cuda.synchronize()
        step_resident()
    prof = core.get_profile()
    core.set_profiling(False)
    n_prof = min(K, 20)
    per_kernel = {k: {"ms_per_step": v[0] / n_prof, "launches_per_step": v[1] / n_prof} for k, v in prof.items() if v[1]}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak_gbs, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak_gbs, peak_src = 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"
    V, B = core.num_voxels(), core.num_buckets()
    # algorithmic bytes per launch, SURVEY.md 8(d)
    alg_bytes = {
        "knn": 52.0 * 0.5 * (n_t + n_s),                    # stage 1 (kNN+cov+reg) 52 B/pt, one cloud per launch
        "covariance": 52.0 * 0.5 * (n_t + n_s),
        "voxelmap_build": 52.0 * n_t + 52.0 * V + 16.0 * B,  # stage 2, whole build
        "linearize": 52.0 * n_s + 52.0 * V + 16.0 * B,       # stage 3, one LM evaluation
        "compute_error": 52.0 * n_s + 52.0 * V + 16.0 * B,
    }
    dom = max((k for k in per_kernel if k in alg_bytes), key=lambda k: per_kernel[k]["ms_per_step"])
    pk = per_kernel[dom]
    launches_dom = pk["launches_per_step"] / (7.0 if dom == "voxelmap_build" else 1.0)  # the build is ~7 small kernels: treat as one unit
    avg_ms = pk["ms_per_step"] / max(launches_dom, 1e-9)
    achieved = alg_bytes[dom] / (avg_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get(dom)
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak_gbs, "unit": "GB/s", "frac": achieved / peak_gbs, "traffic": traffic,
                "peak_source": peak_src, "avg_launch_ms": avg_ms, "algorithmic_bytes_per_launch": alg_bytes[dom],
                "kernel_share_of_step": pk["ms_per_step"] / (sum(v["ms_per_step"] for v in per_kernel.values()) or 1.0)}

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        times, T_cpu, threads = cpu_registration_loop(w, 10.0, 5, 400)
        cpu_baseline = {"value": len(times) / float(np.sum(times)), "unit": UNIT, "cores": threads, "kind": "port",
                        "sample": "%d full registrations (%.1f s) of the same pair, restated OpenMP FastVGICP in double (reference not buildable here)" % (len(times), np.sum(times))}

    T_val = pose_from_c(res.T)
    h2d = (n_t + n_s) * 12
    d2h = n_s * 12 + 16 * 4 + (res.n_linearize * 43 + res.n_compute_error) * 8
    line = {
        "metric": METRIC, "value": world * K / (total_ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": w["data"],
        "config": {"workload": w["name"], "protocol": "align.cpp 100times (covariances recomputed every registration)", "l2": "flushed between steps (256 MiB memset)",
                   "parallelism": "replicas x%d" % world, "n_target": n_t, "n_source": n_s, "num_voxels": V, "num_buckets": B,
                   "lm_iterations": int(res.nr_iterations) + 1, "evaluations": int(res.n_linearize + res.n_compute_error), "converged": bool(res.converged)},
        "e2e": {"value": world * K / (total_ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": total_ms_e2e / K,
                "api": "FastVGICPCuda.setInputTarget/setInputSource/align (pinned host buffers, aligned cloud + pose read back)"},
        "gpu_launches": int(launches), "gpu_launches_e2e": int(launches_e2e),
        "roofline": roofline, "cpu_baseline": cpu_baseline, "clocks": clocks, "per_kernel": per_kernel,
        "wall_ms_per_step_incl_flush": 1e3 * wall_s / K,
        "pose_check": {"translation": [float(x) for x in T_val[:3, 3]], "e2e_vs_resident_max_abs": float(np.abs(np.asarray(T_e2e, dtype=np.float64) - T_val).max())},
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
