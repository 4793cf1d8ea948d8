#!/bin/bash
mkdir -p gpurun_out
show() { python - "$1" <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
print(sys.argv[1], {k:round(d[k],5) if isinstance(d[k],float) else d[k] for k in ("value","ms_per_step","gpu_launches")}, "b2b", round(d["back_to_back"]["value"]), "e2e", round(d["e2e"]["value"]), "rs", round(d["e2e"]["run_solver_loop"]["value"]), "frac", round(d["roofline"]["frac"],3))
PY
}
for i in 1 2; do
timeout 600 python bench.py --steps 200 --warmup 10 --no-cpu > gpurun_out/bench_d_base.json 2> gpurun_out/bench_d.err; show gpurun_out/bench_d_base.json
BDDB200_LIB=$PWD/build_variants/libbdd_b200_prefetch.so timeout 600 python bench.py --steps 200 --warmup 10 --no-cpu > gpurun_out/bench_d_pf.json 2>> gpurun_out/bench_d.err; show gpurun_out/bench_d_pf.json
done
for w in qap_5m grid_mrf_20m; do
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu --workload $w > gpurun_out/bench_d_$w.json 2>> gpurun_out/bench_d.err; show gpurun_out/bench_d_$w.json
BDDB200_LIB=$PWD/build_variants/libbdd_b200_prefetch.so timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu --workload $w > gpurun_out/bench_d_pf_$w.json 2>> gpurun_out/bench_d.err; show gpurun_out/bench_d_pf_$w.json
done
tail -3 gpurun_out/bench_d.err
