#!/bin/bash
# 2-GPU session: parity of all exchange modes, then the bench line with extras
mkdir -p gpurun_out
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
for mode in auto 1 2; do
  echo "== exchange mode $mode"; BDDB200_EXCHANGE_SHOTS=$mode timeout 600 $TR tools/gpu_dist_check.py 2>&1 | grep -E "OK|MISMATCH|FAILED|Error|error" | tail -12
done
timeout 900 $TR bench.py --gpus $N --steps 100 --warmup 5 --no-extras > gpurun_out/bench_f_$N.json 2> gpurun_out/bench_f_$N.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_f_$N.err
python - "$N" <<'PY'
import json,sys
d=json.load(open(f"gpurun_out/bench_f_{sys.argv[1]}.json"))
print({k:d[k] for k in ("value","ms_per_step","gpu_launches","n_gpus")}, "b2b", d["back_to_back"], d["config"]["parallelism"], d["roofline"]["kernel_ms"])
PY
for mode in 1 2; do
BDDB200_EXCHANGE_SHOTS=$mode timeout 900 $TR bench.py --gpus $N --steps 100 --warmup 5 --no-extras > gpurun_out/bench_f_${N}_m$mode.json 2>> gpurun_out/bench_f_$N.err
python - "$N" "$mode" <<'PY'
import json,sys
d=json.load(open(f"gpurun_out/bench_f_{sys.argv[1]}_m{sys.argv[2]}.json"))
print({k:d[k] for k in ("value","ms_per_step","gpu_launches","n_gpus")}, "b2b", d["back_to_back"], d["config"]["parallelism"], d["roofline"]["kernel_ms"])
PY
done
timeout 1200 $TR bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_f_${N}_full.json 2>> gpurun_out/bench_f_$N.err; echo "full rc=$?"
python - "$N" <<'PY'
import json,sys
d=json.load(open(f"gpurun_out/bench_f_{sys.argv[1]}_full.json"))
s=d.get("strong_scaling") or {}
print("full:", d["value"], d["back_to_back"]["value"], "parity_ok", d.get("parity_ok"), "strong", s.get("value"), (s.get("back_to_back") or {}).get("value"), s.get("config",{}).get("parallelism"), d["clocks"])
PY
tail -3 gpurun_out/bench_f_$N.err
