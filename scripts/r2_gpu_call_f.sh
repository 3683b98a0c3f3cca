#!/bin/bash
# round 2 (2 GPUs): multi-GPU test + bench --gpus 2
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -q -x > gpurun_out/r02_multi_gpu_test.log 2>&1
echo "== multi-gpu test: $(tail -1 gpurun_out/r02_multi_gpu_test.log)"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r02_bench_2gpu.json 2> gpurun_out/r02_bench_2gpu.err
tail -3 gpurun_out/r02_bench_2gpu.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_2gpu.json'))
print('value',round(d['value'],1),'e2e',round(d['e2e']['value'],1), d.get('clocks'))
c4=d.get('c4') or {}
print({k:(round(v['ms_per_registration'],3)) for k,v in c4.items() if k.startswith('DIRECT')})
print(json.dumps(c4.get('sharded'))[:1200])
PY
