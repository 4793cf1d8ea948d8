"""The on-chip ("resident") iteration kernel (bdd_b200/csrc/resident.cuh) against the streaming per-pass kernels and the CPU
oracle.  With BDDB200_RESIDENT=1, iteration() / iterations(n) of a solver whose bundles are all lane class and fit one wave run
as ONE cooperative launch; forward_pass() / backward_pass() and all other solvers use the streaming kernels."""
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
    B.oracle_set_num_threads(1)
    yield


def make(col, costs, precision, resident=True, **kw):
    from bdd_b200.solver import bdd_cuda_parallel_mma
    old = os.environ.pop("BDDB200_RESIDENT", None)
    if resident:
        os.environ["BDDB200_RESIDENT"] = "1"
    try:
        return bdd_cuda_parallel_mma(col, costs, precision=precision, device=0, **kw)
    finally:
        os.environ.pop("BDDB200_RESIDENT", None)
        if old is not None:
            os.environ["BDDB200_RESIDENT"] = old


def tol(precision, scale=1.0):
    return (1e-9 if precision == "double" else 2e-4) * max(1.0, abs(scale))


def same_state(a, b, precision):
    la, lb = a.lower_bound(), b.lower_bound()
    assert abs(la - lb) <= tol(precision, lb), (la, lb)
    da, db = a.get_delta().cpu().numpy(), b.get_delta().cpu().numpy()
    assert np.allclose(da, db, rtol=0, atol=tol(precision, np.abs(db).max()))
    for x, y in zip(a.get_solver_costs(), b.get_solver_costs()):
        x, y = x.cpu().numpy(), y.cpu().numpy()
        assert np.allclose(x, y, rtol=0, atol=tol(precision, np.abs(y).max()))


@pytest.mark.parametrize("precision", ["float", "double"])
@pytest.mark.parametrize("shape", [(3000, 6000, 12), (500, 400, 7), (40, 30, 3)])
def test_resident_equals_streaming(shape, precision):
    from bdd_b200 import instances
    m, n, k = shape
    col, costs = instances.set_cover(m=m, n=n, k=k, seed=11)
    a = make(col, costs, precision, resident=True)
    b = make(col, costs, precision, resident=False)
    l0 = a.kernel_launches()
    a.iterations(7)
    assert a.kernel_launches() - l0 <= 2, "iterations(n) of an eligible solver is one cooperative launch (plus the hand-over of the pending sums)"
    l1 = a.kernel_launches()
    a.iterations(2)
    assert a.kernel_launches() - l1 == 1
    b.iterations(2)
    for _ in range(7):
        b.forward_pass(0.5)
        b.backward_pass(0.5)
    same_state(a, b, precision)
    # mixing the two forms on one solver: single iteration, explicit passes, several iterations
    a.iteration()
    a.forward_pass(0.5); a.backward_pass(0.5)
    a.iterations(3)
    for _ in range(5):
        b.iteration()
    same_state(a, b, precision)
    # the min-marginals and per-BDD bounds read the state the on-chip kernel wrote back
    _, lo_a, hi_a = a.min_marginals_cuda(True)
    _, lo_b, hi_b = b.min_marginals_cuda(True)
    n_inner = a.nr_layers() - a.nr_bdds()
    assert np.allclose(lo_a.cpu().numpy()[:n_inner], lo_b.cpu().numpy()[:n_inner], rtol=0, atol=tol(precision, 1e3))
    assert np.allclose(hi_a.cpu().numpy()[:n_inner], hi_b.cpu().numpy()[:n_inner], rtol=0, atol=tol(precision, 1e3))


@pytest.mark.parametrize("precision", ["float", "double"])
def test_resident_against_oracle_with_cost_updates(precision):
    """update_costs invalidates cost_from_terminal: the next launch recomputes it on chip (backward_run) before iterating."""
    from bdd_b200 import instances
    col, costs = instances.set_cover(m=2000, n=4000, k=12, seed=7)
    s = make(col, costs, precision)
    o = B.Oracle(col.instrs, col.delims, costs, precision)
    rng = np.random.default_rng(3)
    for it in range(6):
        if it % 2 == 1:
            pert = rng.integers(-2, 3, size=len(costs)).astype(np.float64)
            s.update_costs(None, pert)
            o.update_costs(None, pert)
        s.iteration()
        o.iteration()
        assert abs(s.lower_bound() - o.lower_bound()) <= tol(precision, o.lower_bound()), it
    assert np.allclose(s.get_delta().cpu().numpy(), o.get_delta(), rtol=0, atol=tol(precision, np.abs(o.get_delta()).max()) * 10)


@pytest.mark.parametrize("precision", ["float", "double"])
def test_resident_mixed_widths_and_lengths(precision):
    """Bundles of different J and BDDs of different length in one launch (simplex rows of several sizes + covers)."""
    from bdd_b200 import instances
    from bdd_b200.lp import EQ, GE
    rng = np.random.default_rng(5)
    n = 300
    batches = []
    for k in (3, 5, 9, 14):
        idx = np.sort(np.stack([rng.choice(n, size=k, replace=False) for _ in range(40)]), axis=1)
        batches.append(instances.ConstraintBatch([1] * k, EQ, 1, idx))
        batches.append(instances.ConstraintBatch([1] * k, GE, 1, idx[::-1].copy()))
    batches.append(instances.ConstraintBatch([1] * 2, GE, 1, np.sort(np.stack([np.arange(n), (np.arange(n) + 1) % n], axis=1), axis=1)))   # every variable covered
    col = instances.from_batches(batches)
    costs = rng.integers(-10, 10, size=n).astype(np.float64)
    s = make(col, costs, precision)
    o = B.Oracle(col.instrs, col.delims, costs, precision)
    for it in range(12):
        s.iteration(); o.iteration()
        assert abs(s.lower_bound() - o.lower_bound()) <= tol(precision, o.lower_bound()), it
    s.iterations(9)
    for _ in range(9):
        o.iteration()
    assert abs(s.lower_bound() - o.lower_bound()) <= tol(precision, o.lower_bound())


def test_full_size_set_cover_resident_vs_streaming():
    """BASELINE config 2 (1.025 M nodes, float): 20 iterations in one launch == 20 x 2 streaming passes (1e-4 relative)."""
    from bdd_b200 import instances
    col, costs = instances.set_cover()
    a = make(col, costs, "float", resident=True)
    b = make(col, costs, "float", resident=False)
    a.iterations(20)
    b.iterations(20)
    la, lb = a.lower_bound(), b.lower_bound()
    assert abs(la - lb) <= 1e-4 * abs(lb), (la, lb)
    da, db = a.get_delta().cpu().numpy(), b.get_delta().cpu().numpy()
    assert np.allclose(da, db, rtol=0, atol=1e-3 * max(1.0, np.abs(db).max()))


@pytest.mark.parametrize("precision", ["float", "double"])
@pytest.mark.parametrize("as_real", [True, False])
def test_fused_host_step_equals_three_calls(precision, as_real):
    """bddb200_step_host == update_costs(host) + iteration() + lower_bound(), step after step (graph replay included), and the
    solver can be used with the separate calls in between."""
    from bdd_b200 import instances
    col, costs = instances.set_cover(m=1500, n=3000, k=10, seed=4)
    a = make(col, costs, precision, resident=False)
    b = make(col, costs, precision, resident=False)
    rng = np.random.default_rng(8)
    dt = a.np_type if as_real else np.float64
    n = len(costs)
    for k in range(9):
        lo = rng.integers(0, 2, size=n).astype(dt)
        hi = rng.integers(-2, 3, size=n).astype(dt)
        if k == 4:                       # a plain iteration in between on both
            a.iteration(); b.iteration()
        la = a.step(lo, hi)
        b.update_costs(lo, hi)
        b.iteration()
        lb = b.lower_bound()
        assert abs(la - lb) <= tol(precision, lb), (k, la, lb)
    same_state(a, b, precision)
    # only hi costs, and no costs at all (a bare iteration + bound)
    la = a.step(None, rng.integers(-1, 2, size=n).astype(dt) * 0 + 1.0)
    b.update_costs(None, np.ones(n, dtype=dt)); b.iteration()
    assert abs(la - b.lower_bound()) <= tol(precision, la)
    la = a.step(None, None)
    b.iteration()
    assert abs(la - b.lower_bound()) <= tol(precision, la)
