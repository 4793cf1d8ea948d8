#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_fullsize_gpu.py tests/test_lbfgs_gpu.py -x -q -m gpu 2>&1 | tail -6
timeout 300 python bench.py --impl reference_cuda --steps 50 --warmup 5 2>/dev/null | cut -c1-500
timeout 300 python bench.py --impl reference_cuda --steps 20 --warmup 3 --workload qap_5m 2>/dev/null | cut -c1-400
timeout 300 python bench.py --impl reference_cuda --steps 10 --warmup 2 --workload grid_mrf_20m 2>/dev/null | cut -c1-400
timeout 300 python bench.py --impl reference_cuda --steps 5 --warmup 1 --workload assignment_5m 2>/dev/null | cut -c1-400
