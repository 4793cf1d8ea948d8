"""The C++ host driver (bdd_b200/csrc/host/bdd_solver_native.hpp, bdd_solver_cl): its .lp reader and BDD builder against the Python
ones (which tests/test_host.py pins against the reference's converter) on every fixture -- no GPU needed -- and, on the GPU box, the
command line solving the fixtures with all four GPU config strings (known answers of test/test_bdd_cuda_parallel_mma.cu:197-247)."""
import glob
import json
import os
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "bdd_b200", "bdd_solver_cl")


def _need_cli():
    if not os.path.exists(CLI):
        pytest.skip("bdd_b200/bdd_solver_cl not built")


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "*.lp"))))
def test_cpp_reader_and_bdd_builder_equal_the_python_ones(path):
    _need_cli()
    from bdd_b200 import instances, lp
    r = subprocess.run([CLI, "--lp-to-bdds", path], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=60)
    assert r.returncode == 0, r.stderr
    out = json.loads(r.stdout)
    ilp = lp.parse_lp(open(path).read())
    col, costs = instances.from_ilp(ilp)
    assert out["var_names"] == ilp.var_names
    assert np.array_equal(np.asarray(out["objective"]), costs) and out["constant"] == ilp.constant
    assert np.array_equal(np.asarray(out["delims"], dtype=np.uint64), col.delims)
    assert np.array_equal(np.asarray(out["instrs"], dtype=np.uint64).reshape(-1, 3), col.instrs)


def test_cpp_reader_handles_the_lp_subset(tmp_path):
    """coefficients with '*', constants on the left, multi-line constraints, identifiers, comment lines, =< / =>, a Bounds tail"""
    _need_cli()
    from bdd_b200 import instances, lp
    text = ("\\ a comment\nMinimize\n 2 x_1 - 3 * y(2) + z + 4\nSubject To\n c1: x_1 + 2 y(2)\n   - z >= 1\n x_1 + y(2) + z + w =< 2\n"
            " r3: 3 w - 2 x_1 + 1 => 0\n w + z = 1\nBounds\n x_1 <= 1\nBinaries\n x_1 z\nEnd\n")
    p = tmp_path / "t.lp"
    p.write_text(text)
    r = subprocess.run([CLI, "--lp-to-bdds", str(p)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=60)
    assert r.returncode == 0, r.stderr
    out = json.loads(r.stdout)
    ilp = lp.parse_lp(text)
    col, costs = instances.from_ilp(ilp)
    assert out["var_names"] == ilp.var_names and out["constant"] == ilp.constant == 4.0
    assert np.array_equal(np.asarray(out["objective"]), costs)
    assert np.array_equal(np.asarray(out["instrs"], dtype=np.uint64).reshape(-1, 3), col.instrs)


def _write_assignment_lp(path, n, seed=3):
    """n x n assignment problem as an .lp file: 2n simplex constraints over n variables each (long, thin BDDs)"""
    rng = np.random.default_rng(seed)
    c = rng.integers(0, 100, size=(n, n))
    lines = ["Minimize", " " + " + ".join(f"{int(c[i, j])} x_{i}_{j}" for i in range(n) for j in range(n)), "Subject To"]
    for i in range(n):
        lines.append(" " + " + ".join(f"x_{i}_{j}" for j in range(n)) + " = 1")
    for j in range(n):
        lines.append(" " + " + ".join(f"x_{i}_{j}" for i in range(n)) + " = 1")
    lines.append("End")
    path.write_text("\n".join(lines) + "\n")


@pytest.mark.parametrize("length", [3, 4, 7, 16])
def test_cpp_splitter_equals_the_python_one(tmp_path, length):
    """bdd_b200/csrc/host/split.hpp (the C++ driver's "split bdds") against bdd_b200/split.py, whose instruction arrays are pinned
    bit for bit to the reference's bdd_collection::split_qbdd (tests/test_split.py): same chunks, same auxiliary variables, same order."""
    _need_cli()
    from bdd_b200 import instances, lp
    from bdd_b200.split import split_long_bdds
    paths = sorted(glob.glob(os.path.join(GOLDEN, "*.lp")))
    big = tmp_path / "assignment_24.lp"
    _write_assignment_lp(big, 24)
    for path in paths + [str(big)]:
        r = subprocess.run([CLI, "--split", str(length), path], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=60)
        assert r.returncode == 0, r.stderr
        out = json.loads(r.stdout)
        ilp = lp.parse_lp(open(path).read())
        col, costs = instances.from_ilp(ilp)
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            want, n_vars = split_long_bdds(col, length, nr_variables=len(costs))
        assert out["nr_variables"] == n_vars, path
        assert np.array_equal(np.asarray(out["delims"], dtype=np.uint64), want.delims), path
        assert np.array_equal(np.asarray(out["instrs"], dtype=np.uint64).reshape(-1, 3), want.instrs), path


@pytest.mark.parametrize("length", [2, 3])
def test_cpp_splitter_with_implication_bdd_equals_the_library(tmp_path, length):
    """bdd_solver_cl --split ... --implication-bdd (the header-only class inside the C++ driver) against the same class behind the C ABI
    (bdd_b200/collection.py), which tests/test_collection.py pins to the reference's split_qbdd(..., with_implication_bdd = true)"""
    _need_cli()
    from bdd_b200 import instances, lp
    from bdd_b200.collection import bdd_collection
    rng = np.random.default_rng(4)
    lines = ["Minimize", " " + " + ".join(f"{int(rng.integers(1, 9))} x{i}" for i in range(20)), "Subject To"]
    for r in range(6):
        vs = sorted(rng.choice(20, size=9, replace=False).tolist())
        lines.append(" " + " + ".join(f"x{v}" for v in vs) + (" = 3" if r % 2 else " <= 4"))
    lines.append("End")
    path = tmp_path / "cardinality.lp"
    path.write_text("\n".join(lines) + "\n")
    r = subprocess.run([CLI, "--split", str(length), str(path), "--implication-bdd"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=60)
    assert r.returncode == 0, r.stderr
    out = json.loads(r.stdout)
    col, costs = instances.from_ilp(lp.parse_lp(path.read_text()))
    c = bdd_collection(col)
    n_split, n_vars = c.split_long_bdds(length, len(costs), True)
    want = c.export()
    assert out["n_split"] == n_split == 6 and out["nr_variables"] == n_vars
    assert want.nr_bdds > 6 * -(-9 // length), "no implication BDD in this instance"
    assert np.array_equal(np.asarray(out["delims"], dtype=np.uint64), want.delims)
    assert np.array_equal(np.asarray(out["instrs"], dtype=np.uint64).reshape(-1, 3), want.instrs)


def test_cpp_split_length_rule_equals_the_python_one(tmp_path):
    _need_cli()
    from bdd_b200 import instances, lp
    from bdd_b200.split import compute_split_length
    big = tmp_path / "assignment_40.lp"
    _write_assignment_lp(big, 40)
    r = subprocess.run([CLI, "--split", "auto", str(big)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=60)
    assert r.returncode == 0, r.stderr
    out = json.loads(r.stdout)
    col, _ = instances.from_ilp(lp.parse_lp(open(big).read()))
    want = compute_split_length(col, n_sms=148)
    assert out["split_length"] == (want if want < 2 ** 62 else -1)
    assert out["n_split"] == 80      # 80 BDDs of 40 variables, fewer than one wave of bundles: all are cut at the minimum length


EXPECTED = {"matching_3x3": -6.0, "short_chain_shuffled": 1.0, "long_chain": -9.0, "grid_graph_3x3": -8.0}


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["cuda parallel mma", "lbfgs cuda mma", "cuda lbfgs parallel mma", "lbfgs cuda parallel mma"])
@pytest.mark.parametrize("precision", ["double", "float"])
def test_cli_solves_the_fixtures_on_the_gpu(kind, precision):
    _need_cli()
    for name, lb in EXPECTED.items():
        path = os.path.join(GOLDEN, name + ".lp")
        if not os.path.exists(path):
            continue
        cfg = {"input": path, "precision": precision, "relaxation solver": kind,
               "termination criteria": {"maximum iterations": 300, "minimum improvement": 1e-12, "improvement slope": 0.0},
               "perturbation rounding": {"initial perturbation": 0.1, "perturbation growth rate": 1.1, "inner iterations": 100, "outer iterations": 100}}
        r = subprocess.run([CLI, json.dumps(cfg)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        out = json.loads(r.stdout.strip().splitlines()[-1])
        tol = 1e-6 if precision == "double" else 1e-3
        assert abs(out["lower_bound"] - lb) <= tol * max(1.0, abs(lb)) + (0.0 if precision == "double" else 0.5), (name, out["lower_bound"])
        assert "solution" in out and abs(out["objective"] - lb) <= 1e-6, (name, out.get("objective"))
        assert "[bdd solver] iteration 0, lower bound" in r.stderr


@pytest.mark.gpu
def test_cli_rejects_what_it_does_not_provide():
    _need_cli()
    path = os.path.join(GOLDEN, "long_chain.lp")
    for cfg in ({"input": path, "relaxation solver": "parallel mma"}, {"input": path, "precision": "half"}, {"relaxation solver": "cuda parallel mma"}):
        r = subprocess.run([CLI, json.dumps(cfg)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=120)
        assert r.returncode == 1 and "bdd_solver_cl:" in r.stderr


@pytest.mark.gpu
def test_cli_split_bdds_solves_the_assignment_problem(tmp_path):
    """"split bdds" in the C++ driver (bdd_solver.cpp:105-123 -> bdd_preprocessor.cpp:372-415): the split relaxation of an n x n assignment
    problem reaches the same optimum (the LP is integral), the rounded solution is feasible and optimal, and the Python driver agrees."""
    _need_cli()
    from scipy.optimize import linear_sum_assignment
    n = 24
    big = tmp_path / "assignment_24.lp"
    _write_assignment_lp(big, n)
    c = np.random.default_rng(3).integers(0, 100, size=(n, n))
    rows, cols = linear_sum_assignment(c)
    opt = float(c[rows, cols].sum())
    base = {"input": str(big), "precision": "double", "relaxation solver": "cuda parallel mma",
            "termination criteria": {"maximum iterations": 3000, "minimum improvement": 1e-12, "improvement slope": 0.0},
            "perturbation rounding": {"initial perturbation": 0.1, "perturbation growth rate": 1.1, "inner iterations": 100, "outer iterations": 100}}
    for split in (None, {"split length": 8}, {}):
        cfg = dict(base)
        if split is not None:
            cfg["split bdds"] = split
        r = subprocess.run([CLI, json.dumps(cfg)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        out = json.loads(r.stdout.strip().splitlines()[-1])
        assert out["lower_bound"] <= opt + 1e-6 and out["lower_bound"] >= opt - 0.05 * abs(opt), (split, out["lower_bound"], opt)
        assert "solution" in out and abs(out["objective"] - opt) <= 1e-6, (split, out.get("objective"), opt)
        if split is not None:
            assert "[bdd preprocessor] split 48 BDDs" in r.stderr
    # the Python driver on the same configuration (dual only: rounding perturbs the costs the bound is read from)
    from bdd_b200.bdd_solver import bdd_solver
    cfg = {k: v for k, v in base.items() if k != "perturbation rounding"}
    s = bdd_solver(log=lambda *a: None)
    s.solve({**cfg, "split bdds": {"split length": 8}})
    r = subprocess.run([CLI, json.dumps({**cfg, "split bdds": {"split length": 8}})], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    assert abs(s.lower_bound() - json.loads(r.stdout.strip().splitlines()[-1])["lower_bound"]) <= 1e-6 * max(1.0, abs(opt))
