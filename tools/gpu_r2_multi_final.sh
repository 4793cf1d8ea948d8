#!/bin/bash
# final multi-GPU evidence with the final build: (2 GPUs only) the multi-GPU parity tests, then the full bench line and the reference arm
mkdir -p gpurun_out
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
if [ "$N" = "2" ]; then timeout 900 python -m pytest tests/test_dist_gpu.py -q -m gpu 2>&1 | tail -2; fi
timeout 400 $TR bench.py --gpus $N --steps 100 --warmup 5 > gpurun_out/r02_final_${N}gpu.json 2> gpurun_out/r02_final_${N}gpu.err; echo "bench rc=$?"
timeout 300 $TR bench.py --impl reference --gpus $N --steps 10 --warmup 2 > gpurun_out/r02_final_${N}gpu_reference.json 2>/dev/null; echo "reference rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_final_${N}gpu.json').read().strip().splitlines()[-1])
ss=d.get('strong_scaling') or {}
print('value', round(d['value']), 'b2b', round(d['back_to_back']['value']), d['config']['parallelism'][-70:], 'parity', d.get('parity_ok'), 'clocks', d['clocks']['sm_mhz'], d['clocks']['reasons'])
print('mrf', round(ss.get('value',0)), round((ss.get('back_to_back') or {}).get('value',0)), (ss.get('config') or {}).get('parallelism','')[-70:])
r=json.loads(open('gpurun_out/r02_final_${N}gpu_reference.json').read().strip().splitlines()[-1]); print('reference', round(r['value'],1), r['cpu_baseline']['cores'])
PY
