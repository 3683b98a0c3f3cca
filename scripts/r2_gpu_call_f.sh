#!/bin/bash
# round 2, GPU call F (2 GPUs): whole GPU suite on GPU 0, multi-GPU test, kNN timing (sub-warp vs warp), bench --gpus 2
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2f_gpu_suite.log 2>&1
echo "== gpu suite: $(tail -2 gpurun_out/r2f_gpu_suite.log)"
grep -E "FAILED|Error|assert " gpurun_out/r2f_gpu_suite.log | head -20
timeout 300 python scripts/exp_knn.py 300 2>&1 | tail -1
VGICP_KNN_SUBWARP=0 timeout 300 python scripts/exp_knn.py 300 2>&1 | tail -1
timeout 300 python scripts/exp_knn_1m.py 10 2>&1 | tail -1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2f_bench_2gpu.json 2> gpurun_out/r2f_bench_2gpu.err
tail -4 gpurun_out/r2f_bench_2gpu.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2f_bench_2gpu.json'))
print('value',round(d['value'],1),'e2e',round(d['e2e']['value'],1), d.get('clocks'))
c4=d.get('c4') or {}
print({k:(round(v['ms_per_registration'],3)) for k,v in c4.items() if k.startswith('DIRECT')})
print(json.dumps(c4.get('sharded'))[:1500])
PY
