#!/bin/bash
# Regenerates profiles/traffic.json (bench.py's roofline.traffic) and the per-kernel ncu summary from one `ncu --set full` capture of the
# C2 workload's kernels (one registration on one stream).  Run under gpurun; the report is summarised on the GPU box (the raw .ncu-rep is
# tens of MB) and only the two JSON files come back in gpurun_out/:
#   gpurun --timeout 1200 -- 'bash scripts/make_traffic.sh'  &&  cp gpurun_out/r02_ncu_full_summary.json profiles/ && cp gpurun_out/traffic.json profiles/
set -u
mkdir -p gpurun_out
REP=/tmp/r02_full.ncu-rep
timeout 1100 ncu --set full --clock-control none -k regex:^k_ -s 150 -c 64 -f -o ${REP%.ncu-rep} \
  python bench.py --streams 1 --steps 2 --warmup 3 --no-cpu-baseline --no-c4 > gpurun_out/r02_full_bench.log 2>&1
tail -1 gpurun_out/r02_full_bench.log
python profiles/summarise_ncu_full.py $REP gpurun_out/r02_ncu_full_summary.json gpurun_out/traffic.json | tee gpurun_out/r02_ncu_full_summary.txt
python profiles/sass_histogram.py $REP k_knn_search 40 > gpurun_out/r02_knn_search_sass_histogram.txt 2>/dev/null
