import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from bdd_b200.solver import lbfgs_cuda_mma, bdd_cuda_parallel_mma
col, costs, precision = bench.make_instance(1, "qap_5m")
s = bdd_cuda_parallel_mma(col, costs, precision=precision)
s.iterations(300); print("plain mma 300 its lb", s.lower_bound()); s.iterations(700); print("plain mma 1000 its lb", s.lower_bound())
del s
for init in (1e-6, 1e-4, 1e-2, 1.0):
    for hist in (5,):
        l = lbfgs_cuda_mma(col, costs, precision=precision, history_size=hist, init_step_size=init)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(300):
            l.iteration()
        lb = l.lower_bound(); dt = time.perf_counter() - t0
        print(f"init_step {init:g} history {hist}: 300 wrapper its in {dt*1e3:.1f} ms, lb {lb:.4f}, stats {l.lbfgs_stats()}")
        del l
