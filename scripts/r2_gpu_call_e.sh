#!/bin/bash
# round 2, GPU call E: whole suite + kNN timing + default bench (C2 + C4 record)
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2e_gpu_suite.log 2>&1
echo "== gpu suite: $(tail -2 gpurun_out/r2e_gpu_suite.log)"
grep -E "FAILED|Error|assert " gpurun_out/r2e_gpu_suite.log | head -20
timeout 300 python scripts/exp_knn.py 300 2>&1 | tail -1
timeout 300 python scripts/exp_knn_1m.py 10 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err
tail -5 gpurun_out/r2e_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2e_bench.json'))
print('value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'single',round(d['single_stream']['ms_per_registration'],4),'cpu',d['cpu_baseline'])
for k,v in d['per_kernel'].items(): print('   ',k,{a:round(b,4) for a,b in v.items()})
print(json.dumps(d.get('c4'))[:3000])
PY
