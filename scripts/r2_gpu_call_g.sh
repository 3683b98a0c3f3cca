#!/bin/bash
# round 2, GPU call G (8 GPUs): bench --gpus 8 (replicas on C2 + C4 sharded over 8 ranks), multi-GPU test on 4 ranks
set -u
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -q -x > gpurun_out/r2g_multi_gpu_test.log 2>&1
echo "== multi-gpu test (4 ranks): $(tail -2 gpurun_out/r2g_multi_gpu_test.log)"
for N in 8 4; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r2g_bench_${N}gpu.json 2> gpurun_out/r2g_bench_${N}gpu.err
tail -4 gpurun_out/r2g_bench_${N}gpu.err
python - <<PY
import json
d=json.load(open('gpurun_out/r2g_bench_${N}gpu.json'))
print('N=${N} value',round(d['value'],1),'e2e',round(d['e2e']['value'],1), d.get('host_placement'))
c4=d.get('c4') or {}
print(json.dumps(c4.get('sharded'))[:1500])
PY
done
