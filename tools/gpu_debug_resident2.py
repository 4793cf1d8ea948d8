"""Debug: iterations(k) in one launch vs k single launches vs streaming."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from bdd_b200 import instances
from test_resident_gpu import make

for (m, n, k) in [(40, 30, 3), (500, 400, 7), (3000, 6000, 12)]:
    col, costs = instances.set_cover(m=m, n=n, k=k, seed=11)
    for K in (2, 3, 7):
        for first in (True, False):
            a = make(col, costs, "double", resident=True)
            b = make(col, costs, "double", resident=False)
            if not first:
                a.iteration(); b.iteration()
            a.iterations(K)
            for _ in range(K):
                b.forward_pass(0.5); b.backward_pass(0.5)
            da, db = a.get_delta().cpu().numpy(), b.get_delta().cpu().numpy()
            ca = [t.cpu().numpy() for t in a.get_solver_costs()]
            cb = [t.cpu().numpy() for t in b.get_solver_costs()]
            print(f"shape {(m, n, k)} K={K} first_call={first}: lb res {a.lower_bound():.6f} stream {b.lower_bound():.6f} | delta {np.abs(da - db).max():.3e}"
                  f" | lo {np.nanmax(np.abs(ca[0] - cb[0])):.3e} hi {np.nanmax(np.abs(ca[1] - cb[1])):.3e} mm {np.nanmax(np.abs(ca[2] - cb[2])):.3e}")
