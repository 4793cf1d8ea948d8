#!/bin/bash
# multi-GPU evidence: all exchange forms with the final build
mkdir -p gpurun_out
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout 600 $TR tools/gpu_dist_check.py 2>&1 | grep -E "PARITY|MISMATCH"
for mode in auto 1 2 mc; do
  BDDB200_EXCHANGE_SHOTS=$mode timeout 600 $TR bench.py --gpus $N --steps 100 --warmup 5 --no-extras > gpurun_out/r02_multi_${N}_$mode.json 2> gpurun_out/r02_multi_$N.err; echo "mode $mode rc=$?"
done
timeout 900 $TR bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r02_multi_${N}_full.json 2>> gpurun_out/r02_multi_$N.err; echo "full rc=$?"
timeout 300 $TR bench.py --impl reference --gpus $N --steps 20 --warmup 2 > gpurun_out/r02_multi_${N}_reference.json 2>/dev/null; echo "reference rc=$?"
