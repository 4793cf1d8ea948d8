"""Where an L-BFGS wrapper iteration spends its time (set_cover_1m or qap_5m): wall clock per iteration and kernel list."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from bdd_b200.solver import lbfgs_cuda_mma
w = sys.argv[1] if len(sys.argv) > 1 else "set_cover_1m"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 200
col, costs, precision = bench.make_instance(1, w)
l = lbfgs_cuda_mma(col, costs, precision=precision, history_size=5)
for _ in range(20):
    l.iteration()
torch.cuda.synchronize()
k0 = l.kernel_launches()
t0 = time.perf_counter()
for _ in range(n):
    l.iteration()
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print(f"{w}: {1e6 * dt / n:.1f} us per wrapper iteration, {(l.kernel_launches() - k0) / n:.1f} solver kernel launches per iteration, stats {l.lbfgs_stats()}, lb {l.lower_bound():.4f}")
