#!/bin/bash
# round 2 evidence session: full gpu test suite, the bench line, the reference arms, ncu launch list and one full capture of the sweep kernels
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt; nproc >> gpurun_out/smi.txt
timeout 2400 python -m pytest tests -q -m gpu 2>&1 | grep -v "cumulative\|bdd parallel" | tail -6
SECONDS=0; timeout 1200 python bench.py > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err; echo "bench rc=$? seconds=$SECONDS"
SECONDS=0; timeout 600 python bench.py --impl reference > gpurun_out/r02_bench_reference.json 2> /dev/null; echo "reference rc=$? seconds=$SECONDS"
timeout 600 python bench.py --impl reference_cuda --steps 100 --warmup 10 > gpurun_out/r02_bench_reference_cuda.json 2> /dev/null; echo "reference_cuda rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 300 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 20 --warmup 3 --no-cpu --no-extras > gpurun_out/ncu_bench.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sweep_lane -s 30 -c 2 -o gpurun_out/r02_prof_sweep -f python bench.py --steps 20 --warmup 3 --no-cpu --no-extras > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
for w in assignment_5m assignment_5m_split64; do timeout 600 python bench.py --steps 50 --warmup 5 --no-extras --workload $w > gpurun_out/r02_bench_$w.json 2> /dev/null; echo "bench $w rc=$?"; done
ls -la gpurun_out | head -40
