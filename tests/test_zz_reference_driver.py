"""The reference's OWN command-line driver -- src/bdd_solver/bdd_solver_cl.cpp + bdd_solver.cpp with every solver, converter and input
source they pull in, unmodified -- linked against the drop-in GPU class and libbdd_b200.so (oracle/Makefile: ref_driver; the .lp reader
and Eigen / tsl are stood in for by test infrastructure, see there).  Here `"relaxation solver": "cuda parallel mma"` is the reference's
construct_solver building LPMP::bdd_cuda_parallel_mma<REAL> into its std::variant, its run_solver loop and its log lines -- with this
repository's kernels underneath.

CPU part: the executable runs the reference's CPU solvers to the published answers (so the shimmed input path is right) and, without a
device, fails in the drop-in constructor with the library's message.  The GPU part could not be run on hardware in the round it was
written (the GPU budget was spent): it is marked xfail(strict=False) so that it reports without gating the suite, and it is the last
file pytest collects."""
import json
import os
import re
import subprocess

import pytest

from conftest import GOLDEN, ROOT

BIN = os.path.join(ROOT, "oracle", "_ref", "ref_bdd_solver_cl")
EXPECTED = {"matching_3x3": -6.0, "short_chain_shuffled": 1.0, "long_chain": -9.0, "grid_graph_3x3": -8.0}
needs_bin = pytest.mark.skipif(not os.path.exists(BIN), reason="oracle/_ref/ref_bdd_solver_cl not built (needs /root/reference at build time)")


def run(name, solver, precision="double", extra=None, timeout=120):
    cfg = {"input": os.path.join(GOLDEN, name + ".lp"), "relaxation solver": solver, "precision": precision,
           "termination criteria": {"maximum iterations": 300, "improvement slope": 0.0, "minimum improvement": 0.0, "time limit": 1e10}}
    cfg.update(extra or {})
    r = subprocess.run([BIN, json.dumps(cfg)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=timeout)
    m = re.findall(r"final lower bound = (\S+)", r.stdout)
    return r, (float(m[-1]) if m else None)


@needs_bin
@pytest.mark.parametrize("name", sorted(EXPECTED))
@pytest.mark.parametrize("solver", ["parallel mma", "sequential mma"])
def test_reference_driver_with_its_cpu_solvers(name, solver):
    r, lb = run(name, solver)
    assert r.returncode == 0 and lb is not None, r.stdout[-2000:]
    assert abs(lb - EXPECTED[name]) <= 1e-5 * max(1.0, abs(EXPECTED[name])), r.stdout[-500:]


@needs_bin
@pytest.mark.parametrize("name", ["matching_3x3", "long_mrf_chain", "mrf_grid_graph_3x3"])
def test_files_written_by_the_reference_are_read_by_both_readers(tmp_path, name):
    """"export lp" of the reference's driver is its own ILP_input::write_lp (include/ILP/ILP_input.h:229-303): the file it writes reads
    back, through the Python reader and through the C++ reader behind the C ABI, as the problem that went in"""
    from bdd_b200 import lp
    from bdd_b200.collection import ilp_input
    out = tmp_path / "exported.lp"
    r, lb = run(name, "parallel mma", extra={"export lp": str(out), "termination criteria": {"maximum iterations": 1}})
    assert r.returncode == 0 and out.exists(), r.stdout[-1000:]
    want = lp.parse_lp(open(os.path.join(GOLDEN, name + ".lp")).read())
    got = lp.parse_lp(out.read_text())
    assert got.var_names == want.var_names and got.objective == want.objective
    assert [(k.variables, k.coefficients, k.ineq, k.rhs) for k in got.constraints] == [(k.variables, k.coefficients, k.ineq, k.rhs) for k in want.constraints]
    native = ilp_input(str(out))
    assert native.variable_names() == want.var_names and native.objective()[0].tolist() == want.objective
    assert [native.constraint(c) for c in range(native.nr_constraints())] == [(k.variables, k.coefficients, k.ineq, k.rhs) for k in want.constraints]


@needs_bin
@pytest.mark.parametrize("name", ["matching_3x3", "short_chain_shuffled", "grid_graph_3x3"])
def test_relaxation_exported_by_the_reference_equals_ours(tmp_path, name):
    """"export bdd lp" of the reference's driver writes the relaxation over the BDDs ITS preprocessor built (BDD manager and all); this
    build's driver writes it over the BDDs of the direct builder.  Node order inside a layer may differ, the relaxation may not: same
    numbers of variables and rows, same optimum under an LP solver."""
    from test_collection import _relaxation_optimum
    from bdd_b200 import bdd_solver as drv, lp
    theirs, ours = tmp_path / "theirs.lp", tmp_path / "ours.lp"
    r, _ = run(name, "parallel mma", extra={"export bdd lp": str(theirs), "termination criteria": {"maximum iterations": 1}})
    assert r.returncode == 0 and theirs.exists(), r.stdout[-1000:]
    try:
        drv.bdd_solver({"input": os.path.join(GOLDEN, name + ".lp"), "relaxation solver": "cuda parallel mma", "export bdd lp": str(ours),
                        "termination criteria": {"maximum iterations": 1}}, log=lambda *a: None)
    except RuntimeError:
        pass                                  # no GPU here: the export happens before the solver is constructed
    a, b = lp.parse_lp(theirs.read_text()), lp.parse_lp(ours.read_text())
    assert len(a.var_names) == len(b.var_names)
    assert abs(_relaxation_optimum(ours.read_text()) - EXPECTED[name]) <= 1e-7
    if name != "short_chain_shuffled":        # BDDs over shuffled variables: the reference ties their first layer to the wrong variable (bdd_collection.h:800) and adds empty rows
        assert len(a.constraints) == len(b.constraints)
        assert abs(_relaxation_optimum(theirs.read_text()) - _relaxation_optimum(ours.read_text())) <= 1e-9


@needs_bin
def test_reference_cpu_lbfgs_equals_its_plain_mma(tmp_path):
    """BASELINE config 5 compares with the reference's `lbfgs parallel mma`.  Run through the reference's own driver it never accepts a
    step at this commit (the stored alpha is uninitialised): its bounds are those of plain `parallel mma`, iteration by
    iteration, while the algorithm its comments describe (oracle/lbfgs_oracle.py, which the GPU implementation follows) does better on
    the same instance.  profiles/r02_reference_cpu_lbfgs.md has the numbers."""
    import sys
    import numpy as np
    import bindings as B
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from lbfgs_oracle import LbfgsOracle
    from bdd_b200 import instances, lp
    rng = np.random.default_rng(1)
    m, n, k = 300, 500, 8
    rows = []
    for r in range(m):
        cols = {r % n}
        while len(cols) < k:
            cols.add(int(rng.integers(0, n)))
        rows.append(sorted(cols))
    c = rng.integers(1, 101, size=n)
    path = tmp_path / "cover.lp"
    path.write_text("Minimize\n " + " + ".join(f"{int(c[j])} x{j}" for j in range(n)) + "\nSubject To\n" + "\n".join(" " + " + ".join(f"x{j}" for j in row) + " >= 1" for row in rows) + "\nEnd\n")
    traces = {}
    for solver in ("parallel mma", "lbfgs parallel mma"):
        cfg = {"input": str(path), "relaxation solver": solver, "precision": "double", "lbfgs": {"history size": 5},
               "termination criteria": {"maximum iterations": 120, "improvement slope": 0.0, "minimum improvement": 0.0, "time limit": 1e10}}
        r = subprocess.run([BIN, json.dumps(cfg)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
        assert r.returncode == 0, r.stdout[-1000:]
        traces[solver] = [float(x) for x in re.findall(r"iteration \d+, lower bound = (\S+),", r.stdout)]
        if solver.startswith("lbfgs"):
            assert r.stdout.count("step size selection unsuccessful") >= 6
    assert len(traces["parallel mma"]) == len(traces["lbfgs parallel mma"]) == 120
    assert np.allclose(traces["parallel mma"], traces["lbfgs parallel mma"], rtol=2e-5, atol=0)      # rejected trial steps are undone: a last printed digit may differ
    col, costs = instances.from_ilp(lp.parse_lp(path.read_text()))
    B.oracle_set_num_threads(1)
    plain, wrapped = B.Oracle(col.instrs, col.delims, costs, "double"), LbfgsOracle(B.Oracle(col.instrs, col.delims, costs, "double"))
    for _ in range(120):
        plain.iteration(); wrapped.iteration()
    assert abs(plain.lower_bound() - traces["parallel mma"][-1]) <= 1e-3 * abs(plain.lower_bound())       # the driver prints six digits
    assert wrapped.lbfgs_iterations > 0 and wrapped.lower_bound() > plain.lower_bound() + 1.0


@needs_bin
def test_reference_driver_reaches_the_dropin_constructor():
    torch = pytest.importorskip("torch")
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    for solver in ("cuda parallel mma", "lbfgs cuda mma", "cuda lbfgs parallel mma"):
        r, _ = run("matching_3x3", solver)
        assert r.returncode != 0 and "no CUDA device available: libbdd_b200 has no CPU fallback" in r.stdout, r.stdout[-1000:]


@needs_bin
@pytest.mark.gpu
@pytest.mark.xfail(strict=False, reason="not yet run on hardware (written after the round's GPU budget was spent); reports without gating")
@pytest.mark.parametrize("precision", ["double", "float"])
@pytest.mark.parametrize("solver", ["cuda parallel mma", "lbfgs cuda mma", "cuda lbfgs parallel mma"])
def test_reference_driver_on_the_dropin_class(solver, precision):
    for name, want in EXPECTED.items():
        r, lb = run(name, solver, precision, timeout=45)            # a run takes a second; the first failure ends the test
        assert r.returncode == 0 and lb is not None, r.stdout[-2000:]
        assert abs(lb - want) <= 1e-3 * max(1.0, abs(want)), (name, lb)       # the reference builds the float solver for "double" and vice versa (bdd_solver.cpp:167-174)
