#!/bin/bash
# multi-GPU: the push exchange (multimem.red inside the pass) against the one-shot pull exchange, parity first
mkdir -p gpurun_out
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
for mode in auto 1; do
  BDDB200_EXCHANGE_SHOTS=$mode timeout 300 $TR tools/gpu_dist_check.py > gpurun_out/push_check_${N}_$mode.log 2>&1; echo "check $mode rc=$?"
  grep -E "PARITY|MISMATCH|exchange=" gpurun_out/push_check_${N}_$mode.log | cut -c1-220 | tail -12
done
for mode in auto 1; do
  BDDB200_EXCHANGE_SHOTS=$mode timeout 400 $TR bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/push_${N}_full_$mode.json 2> gpurun_out/push_${N}_$mode.err; echo "full $mode rc=$?"
done
BDDB200_EXCHANGE_SHOTS=push timeout 300 $TR bench.py --gpus $N --steps 50 --warmup 5 --no-extras > gpurun_out/push_${N}_setcover_push.json 2>> gpurun_out/push_${N}_push.err; echo "set cover push rc=$?"
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/push_${N}_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        ss=d.get('strong_scaling') or {}
        print(f, 'value', round(d['value']), 'b2b', round(d['back_to_back']['value']), d['config']['parallelism'][-60:], 'parity', d.get('parity_ok'),
              '| mrf', round(ss.get('value',0)), round((ss.get('back_to_back') or {}).get('value',0)), (ss.get('config') or {}).get('parallelism','')[-60:])
    except Exception as e: print(f, 'ERR', e)
PY
