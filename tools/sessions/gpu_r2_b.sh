#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools/gpu_debug_resident2.py 2>&1 | grep -E "K=7" 
timeout 900 python -m pytest tests/test_resident_gpu.py -x -q -m gpu 2>&1 | tail -5
timeout 300 python tools/gpu_trace_resident.py 2>&1 | tee gpurun_out/trace_resident.txt | grep -E "==|phase"
timeout 600 python bench.py --steps 200 --warmup 10 --no-cpu > gpurun_out/bench_b.json 2> gpurun_out/bench_b.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_b.json"))
print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, d["back_to_back"], d["e2e"]["value"], d["e2e"]["run_solver_loop"]["value"], d["roofline"]["frac"], d["roofline"]["kernel"])
PY
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8
