#!/bin/bash
# round 2, session A: resident kernel correctness + first timings
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt
timeout 900 python -m pytest tests/test_resident_gpu.py -x -q -m gpu 2>&1 | tail -15
timeout 120 tools/microbench/scatter_cost 2>&1 | tee gpurun_out/scatter_cost.txt
timeout 300 python tools/gpu_trace_resident.py 2>&1 | tee gpurun_out/trace_resident.txt
timeout 600 python bench.py --steps 200 --warmup 10 --no-cpu > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err; echo "bench rc=$?"; cut -c1-1500 gpurun_out/bench_a.json; tail -3 gpurun_out/bench_a.err
BDDB200_NO_RESIDENT=1 timeout 600 python bench.py --steps 200 --warmup 10 --no-cpu > gpurun_out/bench_a_stream.json 2> gpurun_out/bench_a_stream.err; echo "bench(stream) rc=$?"; cut -c1-700 gpurun_out/bench_a_stream.json
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8
