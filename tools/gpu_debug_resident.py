"""Debug: resident vs streaming vs oracle, iteration by iteration."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import bindings as B
from bdd_b200 import instances
from test_resident_gpu import make

B.oracle_set_num_threads(1)
for (m, n, k) in [(40, 30, 3), (64, 60, 4), (100, 90, 5), (500, 400, 7), (3000, 6000, 12)]:
    col, costs = instances.set_cover(m=m, n=n, k=k, seed=11)
    for precision in ("double",):
        a = make(col, costs, precision, resident=True)
        b = make(col, costs, precision, resident=False)
        o = B.Oracle(col.instrs, col.delims, costs, precision)
        print(f"shape {(m, n, k)} {precision}: lb0 res {a.lower_bound():.6f} stream {b.lower_bound():.6f} oracle {o.lower_bound():.6f}")
        for it in range(4):
            a.iteration(); b.forward_pass(0.5); b.backward_pass(0.5); o.iteration()
            da, db, do = a.get_delta().cpu().numpy(), b.get_delta().cpu().numpy(), o.get_delta()
            la, lb, lo = a.lower_bound(), b.lower_bound(), o.lower_bound()
            ca = [t.cpu().numpy() for t in a.get_solver_costs()]
            cb = [t.cpu().numpy() for t in b.get_solver_costs()]
            print(f"   it {it}: lb res {la:.6f} stream {lb:.6f} oracle {lo:.6f} | max|delta res-stream| {np.abs(da - db).max():.3e} stream-oracle {np.abs(db - do).max():.3e}"
                  f" | costs lo {np.nanmax(np.abs(ca[0] - cb[0])):.3e} hi {np.nanmax(np.abs(ca[1] - cb[1])):.3e} mm {np.nanmax(np.abs(ca[2] - cb[2])):.3e} nan {int(np.isnan(ca[0]).sum())}")
            if abs(la - lb) > 1e-6:
                bad = np.nonzero(np.abs(da - db) > 1e-9)[0]
                print("      first differing delta entries", bad[:10], da[bad[:10]], db[bad[:10]])
                break
