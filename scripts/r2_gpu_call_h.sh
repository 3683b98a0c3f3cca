#!/bin/bash
# streams-per-GPU sweep at N=1 (resident and e2e arms)
set -u
mkdir -p gpurun_out
for S in 4 8 12 16 24; do
timeout 300 python bench.py --streams $S --steps 30 --warmup 3 --no-cpu-baseline --no-c4 > gpurun_out/r2h_bench_s$S.json 2> gpurun_out/r2h_bench_s$S.err
python - <<PY
import json
d=json.load(open('gpurun_out/r2h_bench_s$S.json'))
print('S=$S value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'single',round(d['single_stream']['ms_per_registration'],4), {k:round(v['single_stream_ms'],3) for k,v in d['published_configurations'].items()}, {k:round(v.get('value',v.get('registrations_per_s',0)),0) for k,v in d['published_configurations'].items()})
PY
done
