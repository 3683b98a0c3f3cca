#!/bin/bash
# Does the default of 8 hardware work queues (CUDA_DEVICE_MAX_CONNECTIONS) alias the streams of the bench's handles (2 per handle) onto each
# other?  Same bench over streams x connections.  Output: gpurun_out/exp_connections.txt (one trimmed JSON line per run).
mkdir -p gpurun_out
out=gpurun_out/exp_connections.txt
for conn in ${CONNS:-8 16 32 64 128}; do
  for s in ${STREAMS:-8 12 16 24}; do
    echo "== connections=$conn streams=$s ${EXTRA:-}" >> $out
    CUDA_DEVICE_MAX_CONNECTIONS=$conn timeout 300 python bench.py --steps ${STEPS:-60} --warmup 5 --streams $s --no-c4 --no-cpu-baseline ${EXTRA:-} 2>/dev/null \
      | python -c "
import sys, json
for line in sys.stdin:
    line=line.strip()
    if not line.startswith('{'): continue
    d=json.loads(line)
    print(json.dumps({'value': round(d['value']), 'e2e': round(d['e2e']['value']), 'ms_per_step': round(d['ms_per_step'],3), 'single_ms': round(d['single_stream']['ms_per_registration'],3), 'sm_mhz': d.get('clocks',{}).get('sm_mhz')}))
" >> $out
  done
done
cat $out
