"""The reference's own GPU test protocol (test/test_cuda_parallel_mma.cu, test/test_bdd_cuda_parallel_mma.cu),
its driver loop run_solver and its std::variant handling, compiled against the drop-in C++ class
bdd_b200/csrc/host/bdd_solver/bdd_cuda_parallel_mma.h with the reference's headers and CPU objects
(tests/cpp/test_dropin.cu, built by `make -C oracle dropin` where /root/reference exists)."""
import os
import subprocess

import pytest

from conftest import ROOT

BIN = os.path.join(ROOT, "oracle", "_ref", "test_dropin")


@pytest.mark.gpu
def test_reference_protocol_against_dropin_class():
    if not os.path.exists(BIN):
        pytest.skip("oracle/_ref/test_dropin not built (needs /root/reference at build time)")
    r = subprocess.run([BIN], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    tail = "\n".join(r.stdout.splitlines()[-40:])
    assert r.returncode == 0 and "ALL OK" in r.stdout, tail


def test_dropin_header_declares_the_reference_surface():
    """Every member SURVEY 8b lists for the callers of the class is declared by the drop-in header."""
    src = open(os.path.join(ROOT, "bdd_b200", "csrc", "host", "bdd_solver", "bdd_cuda_parallel_mma.h")).read()
    for member in ["iteration(", "lower_bound()", "forward_mm(", "backward_mm(", "normalize_delta(", "net_solver_costs()",
                   "gradient_step(", "bdds_solution_vec()", "make_dual_feasible(", "update_costs(", "nr_layers()",
                   "min_marginals_cuda(", "min_marginals()", "distribute_delta()", "nr_variables()", "nr_bdds()",
                   "get_num_bdds_per_var()", "set_cost(", "get_primal_objective_vector_host()", "nr_hops()",
                   "get_solver_costs(", "set_solver_costs(", "lower_bound_per_bdd(", "bdds_solution()", "using value_type",
                   "flush_forward_states()", "flush_backward_states()", "forward_run()", "backward_run("]:
        assert member in src, member
