"""The reference's outer surface on the GPU path: JSON config -> bdd_solver (bdd_b200/bdd_solver.py mirrors src/bdd_solver/bdd_solver.cpp)
on the .lp fixtures of test/test_problems.h, with the known lower bounds of test/test_bdd_cuda_parallel_mma.cu:197-247 and
test/test_bdd_bipartite_matching_problem.cpp:26-58.  Every GPU solver string of the reference (and the README spelling) is accepted."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import GOLDEN, ROOT

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

EXPECTED = json.load(open(os.path.join(GOLDEN, "expected.json")))


@pytest.mark.parametrize("solver", ["cuda parallel mma", "lbfgs cuda mma", "cuda lbfgs parallel mma", "lbfgs cuda parallel mma"])
@pytest.mark.parametrize("name", ["matching_3x3", "short_chain_shuffled", "long_chain", "grid_graph_3x3", "matching_3x3_first_row"])
def test_config_solves_fixture(name, solver):
    from bdd_b200.bdd_solver import bdd_solver
    cfg = {"input": os.path.join(GOLDEN, name + ".lp"), "precision": "double", "relaxation solver": solver,
           "termination criteria": {"maximum iterations": 300, "minimum improvement": 1e-12, "improvement slope": 0.0},
           "lbfgs": {"history size": 5, "initial step size": 1e-4},
           "perturbation rounding": {"initial perturbation": 0.1, "perturbation growth rate": 1.1, "inner iterations": 100, "outer iterations": 100}}
    s = bdd_solver(cfg, log=lambda *a: None)
    want = EXPECTED[name]["lb"]
    assert s.solution is not None
    n = s.ilp.nr_variables()
    obj = float(np.dot(s.costs[:n], s.solution[:n])) + s.ilp.constant
    assert obj == pytest.approx(want, abs=1e-6)          # the relaxation is tight on these fixtures: rounded optimum == bound
    from bdd_b200.instances import bdds_accept
    assert bdds_accept(s.bdd_col, s.solution).all()
    mms = s.min_marginals()
    assert len(mms) == n and all(m.shape[1] == 2 for m in mms)


def test_inline_lp_string_and_cli():
    lp_text = open(os.path.join(GOLDEN, "matching_3x3.lp")).read()
    cfg = {"input": lp_text, "precision": "float", "relaxation solver": "cuda parallel mma",
           "termination criteria": {"maximum iterations": 200}, "perturbation rounding": {}}
    r = subprocess.run([sys.executable, "-m", "bdd_b200.bdd_solver", json.dumps(cfg)], cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    out = json.loads(r.stdout.strip().splitlines()[-1])
    assert out["lower_bound"] == pytest.approx(-6.0, abs=1e-3)
    assert out["objective"] == pytest.approx(-6.0, abs=1e-6)
    assert sum(out["solution"].values()) == 3


def test_config_errors():
    from bdd_b200.bdd_solver import bdd_solver
    with pytest.raises(RuntimeError, match="no input"):
        bdd_solver({"relaxation solver": "cuda parallel mma"}, log=lambda *a: None)
    with pytest.raises(RuntimeError, match="unknown"):
        bdd_solver({"input": os.path.join(GOLDEN, "matching_3x3.lp"), "relaxation solver": "sequential mma"}, log=lambda *a: None)
