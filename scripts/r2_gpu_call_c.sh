#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "knn or c4_matches or large_cloud or align_matches" > gpurun_out/r2c_knn_tests.log 2>&1
echo "== knn tests: $(tail -3 gpurun_out/r2c_knn_tests.log)"
grep -E "FAILED|Error|assert " gpurun_out/r2c_knn_tests.log | head -20
timeout 300 python scripts/exp_knn.py 300 2>&1 | tail -1
timeout 300 python scripts/exp_knn_1m.py 10 2>&1 | tail -1
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:^k_ -s 130 -c 65 --csv --log-file gpurun_out/r2c_knn_launches.csv python scripts/exp_knn.py 20 > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows=[l for l in open('gpurun_out/r2c_knn_launches.csv') if not l.startswith('==')]
agg=collections.defaultdict(lambda: collections.defaultdict(list))
for r in csv.DictReader(rows):
    agg[r['Kernel Name'].split('(')[0]][r['Metric Name']].append(float(r['Metric Value'].replace(',','')))
for k,v in agg.items():
    t=v.get('gpu__time_duration.sum',[0]); i=v.get('smsp__inst_executed.sum',[0])
    print(f"{k[:50]:50s} n={len(t):3d} avg={sum(t)/len(t)/1000:8.1f} us  instr={sum(i)/len(i)/1e6:7.2f} M")
PY
