"""GPU bring-up helper: runs tiny fixtures through the CUDA path and the oracle and prints the
first divergence in detail.  Not a test; used through gpurun while developing kernels."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import bindings as B
from bdd_b200.instances import BddCollection
from bdd_b200.solver import bdd_cuda_parallel_mma

INT_MAX = 2 ** 31 - 1


def check(name, precision, deterministic, lanes):
    g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    col = BddCollection(g["instrs"], g["delims"])
    B.oracle_set_num_threads(1)
    s = bdd_cuda_parallel_mma(col, g["costs"], precision=precision, deterministic=deterministic, lanes_per_bdd=lanes)
    o = B.Oracle(g["instrs"], g["delims"], g["costs"], precision)
    ii = np.nonzero(s.get_primal_variable_index() != INT_MAX)[0]
    print(f"== {name} {precision} det={deterministic} lanes={lanes}: lb0 gpu={s.lower_bound()!r} oracle={o.lower_bound()!r}")
    delta = torch.zeros(2 * s.nr_variables(), dtype=s.value_type, device="cuda")
    od = np.zeros(2 * o.n_vars, dtype=o.dtype)
    for it in range(3):
        for direction in ("fwd", "bwd"):
            if direction == "fwd":
                s.forward_mm(0.5, delta); o.forward_mm(0.5, od)
            else:
                s.backward_mm(0.5, delta); o.backward_mm(0.5, od)
            d = delta.cpu().numpy()
            lo, hi, mm = (t.cpu().numpy()[ii] for t in s.get_solver_costs())
            olo, ohi, omm = o.get_costs()
            err = [np.abs(d - od).max(), np.abs(lo - olo).max(), np.abs(hi - ohi).max(), np.abs(mm - omm).max()]
            print(f"  it {it} {direction}: max|delta|err={err[0]:.3e} lo={err[1]:.3e} hi={err[2]:.3e} mmd={err[3]:.3e}")
            if max(err) > 1e-4:
                bad = np.nonzero(np.abs(mm - omm) > 1e-4)[0][:10]
                print("   first bad layers:", bad, "vars", s.get_primal_variable_index()[ii][bad], "bdd", s.get_bdd_index()[ii][bad])
                print("   gpu mmd", mm[bad], "oracle", omm[bad])
                print("   gpu lo", lo[bad], "oracle", olo[bad], "gpu hi", hi[bad], "oracle", ohi[bad])
                return False
        print(f"  it {it}: lb gpu={s.lower_bound()!r} oracle={o.lower_bound()!r}")
    return True


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0))
    ok = True
    for name in ("matching_3x3", "short_mrf_chain", "mrf_grid_graph_3x3"):
        for precision in ("double", "float"):
            for lanes in (0, 2, 32):
                ok &= check(name, precision, True, lanes)
            ok &= check(name, precision, False, 0)
    print("ALL OK" if ok else "MISMATCH")
