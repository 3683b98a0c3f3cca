#!/usr/bin/env python
"""Summarise an `ncu --set full` report into per-kernel per-launch averages and the traffic table bench.py reads.

usage:  python profiles/summarise_ncu_full.py <report.ncu-rep> <summary.json> [traffic.json]

Reads `ncu -i <report> --page raw --csv` (units row included; byte-like units are normalised to bytes, time to microseconds),
groups launches by kernel name and averages.  `traffic.json` maps bench.py's kernel categories to
dram__bytes_read.sum + dram__bytes_write.sum per launch.
"""
import csv, io, json, subprocess, sys
from collections import defaultdict

SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "us": 1.0, "ns": 1e-3, "ms": 1e3, "s": 1e6, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3, "second": 1e6}
WANT = {
    "gpu__time_duration.sum": "duration_us",
    "dram__bytes_read.sum": "dram_rd",
    "dram__bytes_write.sum": "dram_wr",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "occupancy_pct",
    "launch__registers_per_thread": "regs",
    "smsp__inst_executed.sum": "warp_instr",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "l1tex__t_sector_hit_rate.pct": "l1_hit",
    "lts__t_sector_hit_rate.pct": "l2_hit",
    "sm__cycles_active.avg": "sm_cycles_active",
    "sm__cycles_elapsed.avg": "sm_cycles_elapsed",
    "smsp__issue_active.avg.per_cycle_active": "issue_per_cycle_active",
}
# bench.py category -> the kernels whose DRAM bytes add up to one launch of that category (one cloud's k-NN stage, one map build, ...)
CATEGORY = {"linearize": ("k_linearize_spec",), "compute_error": ("k_linearize<(int)27, (bool)0",),
            "knn": ("k_grid_bbox", "k_grid_codes", "k_sort_pass<unsigned long long>", "k_grid_levels", "k_grid_table", "k_knn_search", "k_knn_deferred"),
            "covariance": ("k_covariance_knn",),
            "voxelmap_build": ("k_voxel_coords", "k_fill_i32", "k_table_insert", "k_table_lookup_points", "k_table_verdict", "k_table_assign_ids", "k_voxel_sort_keys",
                               "k_sort_pass<unsigned int>", "k_voxel_segments", "k_voxel_reduce")}


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    head, units, body = rows[0], rows[1], rows[2:]
    col = {}
    for i, h in enumerate(head):
        for m, short in WANT.items():
            if h == m or h.endswith("." + m):
                col.setdefault(short, i)
    name_i = head.index("Kernel Name")
    grid_i, block_i = head.index("Grid Size"), head.index("Block Size")
    acc = defaultdict(lambda: defaultdict(list))
    for r in body:
        k = r[name_i]
        for short, i in col.items():
            try:
                v = float(r[i].replace(",", ""))
            except ValueError:
                continue
            acc[k][short].append(v * SCALE.get(units[i], 1.0))
        acc[k]["grid"].append(float(r[grid_i].strip("() ").split(",")[0]))
        acc[k]["block"].append(float(r[block_i].strip("() ").split(",")[0]))
    summary = []
    for k, d in acc.items():
        e = {s: sum(v) / len(v) for s, v in d.items()}
        e["kernel"], e["launches"] = k, len(d["grid"])
        e["traffic_bytes"] = e.get("dram_rd", 0.0) + e.get("dram_wr", 0.0)
        summary.append(e)
    json.dump(summary, open(out, "w"), indent=1)
    for e in summary:
        print(f"{e['kernel'][:60]:60s} x{e['launches']:<3d} {e.get('duration_us', 0):8.1f} us  dram {e['traffic_bytes'] / 1e6:7.2f} MB  instr {e.get('warp_instr', 0) / 1e6:6.2f} M  regs {e.get('regs', 0):4.0f}")
    if len(sys.argv) > 3:
        # per category: sum over its kernels of (DRAM bytes per launch x launches of that kernel per launch of the category)
        import hashlib, os
        traffic = {}
        for cat, pats in CATEGORY.items():
            tot, seen = 0.0, False
            for pat in pats:
                es = [e for e in summary if pat in e["kernel"]]
                if not es:
                    continue
                seen = True
                mult = 3.0 if pat.startswith("k_sort_pass<unsigned long long>") else (2.0 if pat.startswith("k_sort_pass<unsigned int>") else 1.0)  # passes per sort
                tot += mult * sum(e["traffic_bytes"] for e in es) / len(es)
            if seen:
                traffic[cat] = tot
        h = hashlib.sha256()
        d = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "fast_gicp_b200", "csrc")
        for f in sorted(os.listdir(d)):
            if f.endswith((".cu", ".cuh", ".hpp")):
                h.update(open(os.path.join(d, f), "rb").read())
        traffic["source_hash"] = h.hexdigest()[:16]  # bench.py quotes these figures only while the CUDA sources hash to this
        traffic["_note"] = ("dram__bytes_read.sum + dram__bytes_write.sum per launch of a bench.py category (its kernels added up), ncu --set full "
                            "(caches flushed before each replay), C2 workload single stream; written by scripts/make_traffic.sh from " + out)
        json.dump(traffic, open(sys.argv[3], "w"), indent=1)


if __name__ == "__main__":
    main()
