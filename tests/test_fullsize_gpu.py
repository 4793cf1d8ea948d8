"""Full-size parity of BASELINE configs 3 and 4 and the long-chain stress instance against the reference's OWN CPU `parallel mma`
object code (oracle/_ref/libbdd_ref.so, all host threads; the plain-C oracle where that library is absent).  These sizes reach the
launch planner's many-wave branch, the 24-warp float build and the 14-warp double plan that the small shape tests do not.
Tolerances: double 1e-6 * max(1, |LB|), float 1e-4 relative."""
import os

import numpy as np
import pytest

import bindings as B

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _gpu():
    if not torch.cuda.is_available():
        pytest.fail("GPU tests selected but no CUDA device is visible")
    yield
    B.oracle_set_num_threads(1)


def cpu_reference(col, costs, precision):
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    if B.ref_available():
        B.ref_set_num_threads(max(n, B.ref_max_threads()))
        return B.RefSolver(B.RefCollection.from_arrays(col.instrs, col.delims), costs, precision)
    B.oracle_set_num_threads(max(n, B.oracle_max_threads()))
    return B.Oracle(col.instrs, col.delims, costs, precision)


def gpu_solver(col, costs, precision):
    from bdd_b200.solver import bdd_cuda_parallel_mma
    return bdd_cuda_parallel_mma(col, costs, precision=precision, device=0)


def follow(s, o, iterations, rel):
    lb0 = o.lower_bound()
    assert abs(s.lower_bound() - lb0) <= rel * max(1.0, abs(lb0))
    for it in range(iterations):
        s.iteration(); o.iteration()
        a, b = s.lower_bound(), o.lower_bound()
        assert abs(a - b) <= rel * max(1.0, abs(b)), (it, a, b)


def test_qap_5m_double_full_size():
    """BASELINE config 3: QAP-shaped assignment ILP, n = 40, 5.06 M nodes, double, 10 iterations."""
    from bdd_b200 import instances
    col, costs = instances.qap(n=40, seed=2)
    assert abs(col.nr_nodes - 5.0e6) < 0.5e6
    follow(gpu_solver(col, costs, "double"), cpu_reference(col, costs, "double"), 10, 1e-6)


def test_grid_mrf_20m_float_full_size():
    """BASELINE config 4 on one GPU: 283 x 283 grid MRF with 4 labels, 20.0 M nodes, float, 6 iterations."""
    from bdd_b200 import instances
    col, costs = instances.grid_mrf(283, 283, 4, seed=4)
    assert abs(col.nr_nodes - 20.0e6) < 1.0e6
    follow(gpu_solver(col, costs, "float"), cpu_reference(col, costs, "float"), 6, 1e-4)


@pytest.mark.parametrize("precision,rel", [("float", 1e-4), ("double", 1e-6)])
def test_assignment_1118_long_chains(precision, rel):
    """Config 3b (stress): 2 x 1118 simplex BDDs of 1118 variables, H = 1118; unsplit, and cut by split_long_bdds at the length
    compute_split_length picks (the split instance has the same optimum; its bound is compared with ITS OWN CPU solve)."""
    from bdd_b200 import instances
    from bdd_b200.split import compute_split_length, split_long_bdds
    col, costs = instances.assignment(1118, seed=3)
    follow(gpu_solver(col, costs, precision), cpu_reference(col, costs, precision), 5, rel)
    scol, n_all = split_long_bdds(col, compute_split_length(col))
    scosts = np.concatenate([costs, np.zeros(n_all - len(costs))])
    follow(gpu_solver(scol, scosts, precision), cpu_reference(scol, scosts, precision), 5, rel)


def test_reference_cuda_solver_agrees():
    """The reference's own CUDA solver (oracle/_ref/libbdd_ref_cuda.so, unmodified sources built for sm_100a) and this repository's
    kernels follow each other on the 1 M-node instance (float, 1e-4 relative) and on a QAP-shaped one (double, 1e-6)."""
    if not B.ref_cuda_available():
        pytest.skip("oracle/_ref/libbdd_ref_cuda.so not built")
    from bdd_b200 import instances
    for (col, costs), precision, rel in ((instances.set_cover(), "float", 1e-4), (instances.qap(n=14, seed=2), "double", 1e-6)):
        s = gpu_solver(col, costs, precision)
        r = B.RefCudaSolver(col.instrs, col.delims, costs, precision)
        assert abs(s.lower_bound() - r.lower_bound()) <= rel * max(1.0, abs(r.lower_bound()))
        for it in range(8):
            s.iteration(); r.iteration()
            a, b = s.lower_bound(), r.lower_bound()
            assert abs(a - b) <= rel * max(1.0, abs(b)), (precision, it, a, b)
