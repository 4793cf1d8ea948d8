#!/bin/bash
# bench lines under several kernel tuning settings; each argument is a string of env assignments
# (BDDB200_* kernel knobs, BENCH_WORKLOAD=...), e.g. "BDDB200_LANES_PER_BDD=2 BENCH_WORKLOAD=qap_5m"
mkdir -p gpurun_out
if [ -n "$RUN_TESTS" ]; then timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log | cut -c1-300; fi
: > gpurun_out/tune.log
run() {
  echo "== $*" >> gpurun_out/tune.log
  env "$@" timeout 300 python bench.py --steps ${STEPS:-100} --warmup 5 --no-cpu 2>>gpurun_out/tune.err | python -c "
import json,sys
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('value',round(d['value']),'b2b',round(d['back_to_back']['value']),'kern_us',round(d['roofline']['kernel_ms']*1e3,2),'frac',round(d['roofline']['frac'],3),'e2e',round(d['e2e']['value']),'rs',round(d['e2e']['run_solver_loop']['value']), 'construct_ms', round(d['construct_ms'],1), 'N', d['config']['N'])
" >> gpurun_out/tune.log
}
for cfg in "$@"; do run $cfg; done
cat gpurun_out/tune.log
