#!/bin/bash
# N GPUs: parity with the default exchange choice, then the full bench line (set cover weak scaling + grid_mrf_20m strong scaling) with the
# default choice (push for the MRF) and with the one-shot pull exchange forced
mkdir -p gpurun_out
N=${1:-4}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout 300 $TR tools/gpu_dist_check.py > gpurun_out/push_check_${N}_auto.log 2>&1; echo "check auto rc=$?"
grep -E "PARITY|MISMATCH" gpurun_out/push_check_${N}_auto.log; grep -c "push inside" gpurun_out/push_check_${N}_auto.log
for mode in auto 1; do
  BDDB200_EXCHANGE_SHOTS=$mode timeout 400 $TR bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/push_${N}_full_$mode.json 2> gpurun_out/push_${N}_$mode.err; echo "full $mode rc=$?"
done
timeout 300 $TR bench.py --impl reference --gpus $N --steps 10 --warmup 2 > gpurun_out/push_${N}_reference.json 2>/dev/null; echo "reference rc=$?"
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/push_${N}_full_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        ss=d.get('strong_scaling') or {}
        print(f, 'value', round(d['value']), 'b2b', round(d['back_to_back']['value']), d['config']['parallelism'][-60:], 'parity', d.get('parity_ok'),
              '| mrf', round(ss.get('value',0)), round((ss.get('back_to_back') or {}).get('value',0)), (ss.get('config') or {}).get('parallelism','')[-60:])
    except Exception as e: print(f, 'ERR', e)
PY
