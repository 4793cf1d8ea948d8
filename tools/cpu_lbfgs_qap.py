"""CPU-only check of what the L-BFGS wrapper does on QAP-shaped instances: the numpy restatement of the reference algorithm (oracle/lbfgs_oracle.py) on top of the
plain-C oracle solver, next to plain MMA iterations.  Test infrastructure; no GPU, nothing of bdd_b200/csrc involved.  Output: profiles/r02_lbfgs_qap_cpu_oracle.md."""
import sys, time
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/oracle')
import numpy as np
import bindings as B
from lbfgs_oracle import LbfgsOracle
from bdd_b200 import instances
B.oracle_set_num_threads(8)
for n in (16, 24, 32):
    col, costs = instances.qap(n=n, seed=2)
    o = B.Oracle(col.instrs, col.delims, costs, "double")
    l = LbfgsOracle(o, history_size=5, init_step_size=1e-6)
    p = B.Oracle(col.instrs, col.delims, costs, "double")
    t=time.time()
    marks={}
    for it in range(1,401):
        l.iteration(); p.iteration()
        if it in (100,200,400): marks[it]=(l.lbfgs_iterations, l.mma_iterations, l.lower_bound(), p.lower_bound(), l.step_size)
    print(f"qap n={n} nodes {col.nr_nodes}:", {k:(v[0],v[1],round(v[2],4),round(v[3],4),f'{v[4]:.2e}') for k,v in marks.items()}, f"({time.time()-t:.1f}s)", flush=True)
