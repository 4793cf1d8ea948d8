#!/bin/bash
# One GPU session: bench line, ncu launch list of the same command, one full ncu capture of the sweep kernels,
# bench lines of the other workloads.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt
nproc > gpurun_out/nproc.txt
timeout 600 python bench.py --steps 200 --warmup 10 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/ncu_bench.log 2>&1; echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sweep_ -s 20 -c 4 -o gpurun_out/prof_sweep -f python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
for w in qap_5m grid_mrf_20m assignment_5m assignment_5m_split64; do
  timeout 600 python bench.py --steps 100 --warmup 5 --workload $w > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo "bench $w rc=$?"
  cut -c1-600 gpurun_out/bench_$w.json
done
timeout 300 python bench.py --impl reference --steps 20 --warmup 2 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "reference rc=$?"; cut -c1-700 gpurun_out/bench_reference.json
ls -la gpurun_out
