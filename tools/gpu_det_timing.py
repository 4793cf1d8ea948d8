"""b2b timing of the deterministic (segmented-sum) streaming path vs the atomic one, set_cover_1m float."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from bdd_b200 import instances
from bdd_b200.solver import bdd_cuda_parallel_mma
col, costs = instances.set_cover()
for det in (False, True):
    s = bdd_cuda_parallel_mma(col, costs, precision="float", deterministic=det)
    s.iterations(30); s.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for rep in range(3):
        e0.record(s.stream); s.iterations(300); e1.record(s.stream); s.synchronize()
        print(f"deterministic={det}: {e0.elapsed_time(e1) / 300 * 1e3:.2f} us per iteration (graph replay), lb {s.lower_bound():.4f}")
