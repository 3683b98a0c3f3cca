#!/bin/bash
# Regenerates profiles/traffic.json (bench.py's roofline.traffic) and profiles/r02_ncu_full_summary.json from one ncu --set full capture
# of the C2 workload's kernels.  Run under gpurun:  gpurun --timeout 1200 -- 'bash scripts/make_traffic.sh'  then, here (needs ncu only):
#   python profiles/summarise_ncu_full.py gpurun_out/r02_full.ncu-rep profiles/r02_ncu_full_summary.json profiles/traffic.json
set -u
mkdir -p gpurun_out
timeout 1100 ncu --set full --clock-control none --import-source on -k regex:^k_ -s 150 -c 75 -f -o gpurun_out/r02_full \
  python bench.py --streams 1 --steps 2 --warmup 3 --no-cpu-baseline --no-c4 > gpurun_out/r02_full_bench.log 2>&1
tail -2 gpurun_out/r02_full_bench.log
