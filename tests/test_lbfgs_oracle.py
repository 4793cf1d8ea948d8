"""The numpy L-BFGS restatement (oracle/lbfgs_oracle.py, test infrastructure) on the CPU oracle: the properties the
algorithm guarantees by construction.  (No reference vectors exist: the reference template is broken)"""
import numpy as np
import pytest

import bindings as B
from lbfgs_oracle import LbfgsOracle


@pytest.mark.parametrize("name", ["set_cover", "grid_mrf"])
def test_oracle_lbfgs_monotone_and_not_behind_mma(name):
    from bdd_b200 import instances
    col, costs = {"set_cover": lambda: instances.set_cover(m=400, n=700, k=8, seed=3),
                  "grid_mrf": lambda: instances.grid_mrf(7, 6, 3, seed=2)}[name]()
    B.oracle_set_num_threads(1)
    l = LbfgsOracle(B.Oracle(col.instrs, col.delims, costs, "double"), init_step_size=1e-3)
    plain = B.Oracle(col.instrs, col.delims, costs, "double")
    lbs = [l.lower_bound()]
    for _ in range(30):
        l.iteration(); plain.iteration()
        lbs.append(l.lower_bound())
        assert lbs[-1] >= lbs[-2] - 1e-9 * max(1.0, abs(lbs[-2]))
    assert l.lbfgs_iterations > 0 and l.mma_iterations >= 5
    assert l.lower_bound() >= plain.lower_bound() - 1e-6 * abs(plain.lower_bound())
    # the reparametrisation still sums to the objective
    l.o.distribute_delta()
    assert np.allclose(l.o.primal_objective(), costs, rtol=0, atol=1e-8)
