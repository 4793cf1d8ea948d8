#!/bin/bash
mkdir -p gpurun_out
show() { python - "$1" <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
print(sys.argv[1], {k:round(d[k],5) if isinstance(d[k],float) else d[k] for k in ("value","ms_per_step","gpu_launches")}, "b2b", round(d["back_to_back"]["value"]), "e2e", round(d["e2e"]["value"]), "rs", round(d["e2e"]["run_solver_loop"]["value"]), "frac", round(d["roofline"]["frac"],3))
PY
}
timeout 600 python bench.py --steps 200 --warmup 10 --no-cpu > gpurun_out/bench_c_bal.json 2> gpurun_out/bench_c.err; show gpurun_out/bench_c_bal.json
BDDB200_NO_BALANCE=1 timeout 600 python bench.py --steps 200 --warmup 10 --no-cpu > gpurun_out/bench_c_nobal.json 2>> gpurun_out/bench_c.err; show gpurun_out/bench_c_nobal.json
BDDB200_LIB=$PWD/build_variants/libbdd_b200_gather8.so timeout 600 python bench.py --steps 200 --warmup 10 --no-cpu > gpurun_out/bench_c_g8.json 2>> gpurun_out/bench_c.err; show gpurun_out/bench_c_g8.json
BDDB200_RESIDENT=1 timeout 600 python bench.py --steps 200 --warmup 10 --no-cpu > gpurun_out/bench_c_res.json 2>> gpurun_out/bench_c.err; show gpurun_out/bench_c_res.json
tail -3 gpurun_out/bench_c.err
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
