#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_resident_gpu.py -x -q -m gpu 2>&1 | tail -5
timeout 600 python bench.py --steps 200 --warmup 10 --no-cpu > gpurun_out/bench_e.json 2> gpurun_out/bench_e.err; python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_e.json"))
print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, d["back_to_back"]["value"], d["e2e"], d["roofline"]["frac"])
PY
tail -3 gpurun_out/bench_e.err
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu --workload qap_5m > gpurun_out/bench_e_qap.json 2>> gpurun_out/bench_e.err; python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_e_qap.json"))
print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, d["back_to_back"]["value"], d["e2e"], d["roofline"]["frac"])
PY
