"""The plain-C oracle against the reference's own CPU solver (oracle/_ref/libbdd_ref.so, built
from /root/reference) on instances beyond the committed fixtures.  CPU only; skipped where the
reference library was not built."""
import numpy as np
import pytest

import bindings as B
from bdd_b200 import instances

pytestmark = pytest.mark.skipif(not B.ref_available(), reason="oracle/_ref/libbdd_ref.so not built (needs /root/reference)")


def ref_solver(col, costs, precision):
    rc = B.RefCollection.from_arrays(col.instrs, col.delims)
    return B.RefSolver(rc, costs, precision)


@pytest.mark.parametrize("precision", ["double", "float"])
@pytest.mark.parametrize("gen", ["cover", "qap", "mrf", "assignment", "random"])
def test_oracle_equals_reference_bitwise(gen, precision):
    col, costs = {
        "cover": lambda: instances.set_cover(m=300, n=500, k=9, seed=5),
        "qap": lambda: instances.qap(n=6, seed=2),
        "mrf": lambda: instances.grid_mrf(6, 5, 3, seed=4),
        "assignment": lambda: instances.assignment(12, seed=3),
        "random": lambda: instances.random_inequalities(80, 50, max_len=9, max_coeff=4, seed=9),
    }[gen]()
    B.ref_set_num_threads(1); B.oracle_set_num_threads(1)
    r = ref_solver(col, costs, precision)
    o = B.Oracle(col.instrs, col.delims, costs, precision)
    assert r.n_vars == o.n_vars and r.n_bdds == o.n_bdds and r.n_layers == o.n_layers
    assert r.lower_bound() == o.lower_bound()
    for it in range(30):
        r.iteration(); o.iteration()
        assert r.lower_bound() == o.lower_bound(), f"iteration {it}"
    assert np.array_equal(r.bdds_solution(), o.bdds_solution())


def test_reference_converter_matches_own_qbdd_builder():
    """Own QBDD builder vs the reference's lineq_bdd -> bdd_mgr -> reorder -> make_qbdd chain:
    same node count per constraint, and the same lower-bound trajectory."""
    rng = np.random.default_rng(3)
    rc = B.RefCollection()
    batches = []
    nv = 30
    for _ in range(40):
        ln = int(rng.integers(2, 9))
        vs = np.sort(rng.choice(nv, size=ln, replace=False))
        co = (rng.integers(1, 5, size=ln) * rng.choice([-1, 1], size=ln)).tolist()
        ineq = int(rng.integers(0, 3))
        rhs = int(rng.integers(min(0, sum(c for c in co if c < 0)), sum(c for c in co if c > 0) + 1))
        try:
            t = instances.qbdd_template(co, ineq, rhs)
        except ValueError:
            continue
        if t is None or not instances._both_values_feasible(t):
            continue
        nr = rc.add_constraint(co, vs, ineq, rhs)
        assert nr >= 0
        batches.append(instances.ConstraintBatch(co, ineq, rhs, vs[None, :]))
    col = instances.from_batches(batches)
    instrs, delims = rc.export()
    assert np.array_equal(delims, col.delims)
    costs = np.zeros(nv)
    costs[:] = rng.integers(-5, 6, size=nv)
    used = np.unique(col.instrs[col.instrs[:, 2] < instances.BOTSINK, 2]).astype(np.int64)
    mask = np.zeros(nv, bool); mask[used] = True
    costs[~mask] = 0
    B.oracle_set_num_threads(1)
    a = B.Oracle(instrs, delims, costs[: int(used.max()) + 1], "double")
    b = B.Oracle(col.instrs, col.delims, costs[: int(used.max()) + 1], "double")
    for it in range(20):
        a.iteration(); b.iteration()
        assert abs(a.lower_bound() - b.lower_bound()) < 1e-9
