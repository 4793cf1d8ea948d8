"""GPU primal rounding (bddb200_rounding_perturb / bddb200_incremental_mm_agreement_rounding) against the numpy restatement of
src/bdd_solver/incremental_mm_agreement_rounding_cuda.cu (oracle/rounding_oracle.py) and against what a primal solution must
satisfy: every BDD accepts it, its cost is >= the lower bound and equals it where the relaxation is tight."""
import os

import numpy as np
import pytest

import bindings as B
import rounding_oracle as R
from conftest import GOLDEN

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

INT_MAX = 2 ** 31 - 1


def _col(g):
    from bdd_b200.instances import BddCollection
    return BddCollection(np.ascontiguousarray(g["instrs"], np.uint64), np.ascontiguousarray(g["delims"], np.uint64))


@pytest.mark.parametrize("precision", ["double", "float"])
@pytest.mark.parametrize("name", ["mrf_grid_graph_3x3", "long_mrf_chain", "matching_3x3_first_row"])
def test_perturbation_round_matches_restatement(name, precision):
    from bdd_b200.solver import bdd_cuda_parallel_mma
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    s = bdd_cuda_parallel_mma(_col(g), g["costs"], precision=precision, deterministic=True)
    for _ in range(3):
        s.iteration()
    s.distribute_delta()
    mm_vars, mm0, mm1 = (t.cpu().numpy() for t in s.min_marginals_cuda(True))
    obj_before = s.get_primal_objective_vector_host()
    types_o, s0, s1, mn = R.classify(mm_vars, mm0, mm1, s.nr_variables())
    delta, rnd = 0.37, 4
    sol, counts, types = s.rounding_perturb(delta, rnd)
    assert np.array_equal(types.cpu().numpy(), types_o)
    assert counts.tolist() == [int((types_o == k).sum()) for k in range(4)]
    if counts[0] + counts[1] == s.nr_variables():
        assert sol is not None and np.array_equal(sol, (types_o == R.ONE).astype(np.int8))
        return
    assert sol is None
    d0, d1 = R.perturbation(types_o, s0, s1, mn, delta, rnd, s.np_type)
    # update_costs(d0, d1) moved the objective (sum over BDDs of hi - lo) by d1 - d0
    obj_after = s.get_primal_objective_vector_host()
    assert np.allclose(obj_after - obj_before, d1.astype(np.float64) - d0.astype(np.float64), rtol=0, atol=1e-9 if precision == "double" else 1e-5)


@pytest.mark.parametrize("name,optimum", [("matching_3x3", -6.0), ("matching_3x3_first_row", -4.0), ("short_mrf_chain", 1.0), ("long_mrf_chain", -9.0), ("mrf_grid_graph_3x3", -8.0)])
def test_rounding_finds_the_optimum_on_tight_fixtures(name, optimum):
    """Known answers of test/test_bdd_cuda_parallel_mma.cu:197-247 / test_bdd_bipartite_matching_problem.cpp: the relaxation is tight,
    so the rounded solution is feasible and attains the bound."""
    from bdd_b200.solver import bdd_cuda_parallel_mma
    from bdd_b200.instances import bdds_accept
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    col = _col(g)
    s = bdd_cuda_parallel_mma(col, g["costs"], precision="double")
    s.run_solver(max_iter=200)
    sol, rounds = s.incremental_mm_agreement_rounding(init_delta=0.1, delta_growth_rate=1.1, num_itr_lb=100, num_rounds=100)
    assert sol is not None and rounds >= 1
    assert bdds_accept(col, sol).all()
    cost = float(np.dot(g["costs"], sol[: len(g["costs"])]))
    assert cost == pytest.approx(optimum, abs=1e-9)


@pytest.mark.parametrize("solver_kind", ["mma", "lbfgs"])
def test_rounding_set_cover_is_feasible(solver_kind):
    from bdd_b200 import instances
    from bdd_b200.solver import bdd_cuda_parallel_mma, lbfgs_cuda_mma
    col, costs = instances.set_cover(m=300, n=500, k=6, seed=9)
    s = (lbfgs_cuda_mma(col, costs, precision="double", init_step_size=1e-3) if solver_kind == "lbfgs"
         else bdd_cuda_parallel_mma(col, costs, precision="double"))
    lb = s.run_solver(max_iter=300)
    sol, rounds = s.incremental_mm_agreement_rounding(init_delta=0.1, delta_growth_rate=1.1, num_itr_lb=100, num_rounds=200)
    assert sol is not None
    assert instances.bdds_accept(col, sol).all()              # every row is covered
    cost = float(np.dot(costs, sol))
    assert cost >= lb - 1e-6 and cost <= 1.25 * lb            # integer costs in [1, 100]; the rounded cover stays near the bound


@pytest.mark.parametrize("precision", ["double", "float"])
@pytest.mark.parametrize("case", ["mrf_grid_graph_3x3", "long_mrf_chain", "matching_3x3", "cover_300x500"])
def test_perturbation_round_against_the_reference_decoder(case, precision):
    """The device rounding kernel against the reference's own CPU decoder (mm_primal_decoder, compiled into oracle/_ref): agreement type
    per variable, type statistics, and the side the perturbation is applied to for every variable whose side is not random
    (one / zero: by type; inconsistent: by the min-marginal sums, incremental_mm_agreement_rounding.hxx:100-140)."""
    if not B.ref_available():
        pytest.skip("oracle/_ref/libbdd_ref.so not built")
    from bdd_b200 import instances
    from bdd_b200.solver import bdd_cuda_parallel_mma
    if case == "cover_300x500":
        col, costs = instances.set_cover(m=300, n=500, k=6, seed=9)
    else:
        g = np.load(os.path.join(GOLDEN, case + ".npz"))
        col, costs = _col(g), g["costs"]
    s = bdd_cuda_parallel_mma(col, costs, precision=precision)
    for _ in range(4):
        s.iteration()
    s.distribute_delta()
    mm = s.min_marginals()
    types_r, sums_r, stats_r, sol_r = B.ref_mm_decode(mm)
    obj_before = s.get_primal_objective_vector_host()
    delta = 0.25
    sol, counts, types = s.rounding_perturb(delta, 2)
    types = types.cpu().numpy()
    assert np.array_equal(types, types_r)
    assert {"zero": int(counts[0]), "one": int(counts[1]), "equal": int(counts[2]), "inconsistent": int(counts[3])} == stats_r
    if sol_r is not None:
        assert sol is not None and np.array_equal(sol, sol_r)
        return
    change = s.get_primal_objective_vector_host() - obj_before          # d1 - d0 per variable
    tol = 1e-9 if precision == "double" else 1e-5
    for v in range(len(types_r)):
        if types_r[v] == 1:
            assert abs(change[v] + delta) <= tol
        elif types_r[v] == 0:
            assert abs(change[v] - delta) <= tol
        elif types_r[v] == 3 and abs(sums_r[v, 0] - sums_r[v, 1]) > 1e-3:
            assert (change[v] > 0) == (sums_r[v, 0] < sums_r[v, 1]) and 0 < abs(change[v]) <= delta * delta + tol
        else:
            assert 0 <= abs(change[v]) <= delta * delta + tol
