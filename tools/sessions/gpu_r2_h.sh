#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout 600 $TR tools/gpu_dist_check.py 2>&1 | grep -E "PARITY|MISMATCH"
timeout 900 $TR bench.py --gpus $N --steps 100 --warmup 5 --no-extras > gpurun_out/bench_h_$N.json 2> gpurun_out/bench_h_$N.err; echo "bench rc=$?"
BDDB200_NO_PDL=1 timeout 900 $TR bench.py --gpus $N --steps 100 --warmup 5 --no-extras > gpurun_out/bench_h_${N}_nopdl.json 2>> gpurun_out/bench_h_$N.err
timeout 1200 $TR bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_h_${N}_full.json 2>> gpurun_out/bench_h_$N.err; echo "full rc=$?"
grep -v "Warning\|symm_mem\|^\*\|OMP_NUM" gpurun_out/bench_h_$N.err | tail -5
