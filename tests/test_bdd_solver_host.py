"""Host side of the JSON-config driver (bdd_b200/bdd_solver.py; reference: src/bdd_solver/bdd_solver.cpp:44-123, 468-475): configuration
reading, ILP parsing, BDD construction and splitting run without a GPU; constructing a solver without one fails loudly."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN
from bdd_b200 import bdd_solver as drv
from bdd_b200 import instances


def quiet(*a):
    pass


def test_read_config_file_and_inline(tmp_path):
    cfg = {"input": os.path.join(GOLDEN, "matching_3x3.lp"), "relaxation solver": "cuda parallel mma"}
    f = tmp_path / "config.json"
    f.write_text(json.dumps(cfg))
    assert drv.read_config(str(f)) == cfg
    assert drv.read_config(json.dumps(cfg)) == cfg


@pytest.mark.parametrize("name", ["matching_3x3", "long_mrf_chain", "mrf_grid_graph_3x3"])
def test_ilp_to_bdds_matches_fixture(name):
    s = drv.bdd_solver(log=quiet)
    s.ilp = s.read_ILP({"input": os.path.join(GOLDEN, name + ".lp")})
    col, costs = s.transform_to_BDDs({})
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    assert np.array_equal(col.delims, g["delims"]) and np.array_equal(costs, g["costs"])
    # the same from an inline LP string
    s2 = drv.bdd_solver(log=quiet)
    s2.ilp = s2.read_ILP({"input": open(os.path.join(GOLDEN, name + ".lp")).read()})
    assert s2.ilp.var_names == s.ilp.var_names


def test_split_key():
    s = drv.bdd_solver(log=quiet)
    s.ilp = s.read_ILP({"input": os.path.join(GOLDEN, "long_mrf_chain.lp")})
    whole, costs = s.transform_to_BDDs({})
    longest = int(np.diff(whole.delims.astype(np.int64)).max())
    split, costs2 = s.transform_to_BDDs({"split bdds": {"split length": 2}})
    assert np.array_equal(costs, costs2)                                   # auxiliary variables carry no cost
    if longest > 2 * 2 + 2:
        assert split.nr_bdds > whole.nr_bdds and split.nr_variables() > whole.nr_variables()
    # with the implication BDD: the same chunks in the same order, plus at most one BDD over auxiliary variables per cut BDD
    with_imp, _ = s.transform_to_BDDs({"split bdds": {"split length": 2, "implication bdd": True}})
    assert split.nr_bdds <= with_imp.nr_bdds and with_imp.nr_variables() == split.nr_variables()
    n_orig = whole.nr_variables()
    extra = 0
    for b in range(with_imp.nr_bdds):
        idx = with_imp.instrs[int(with_imp.delims[b]):int(with_imp.delims[b + 1]) - 2, 2]
        extra += bool((idx >= n_orig).all())
    assert extra == with_imp.nr_bdds - split.nr_bdds
    # a constraint long enough for three cuts whose paths exclude combinations of cut nodes: one implication BDD
    s = drv.bdd_solver(log=quiet)
    s.ilp = s.read_ILP({"input": "Minimize\n " + " + ".join(f"{i + 1} x{i}" for i in range(8)) + "\nSubject To\n " + " + ".join(f"x{i}" for i in range(8)) + " = 3\nEnd\n"})
    split, _ = s.transform_to_BDDs({"split bdds": {"split length": 2}})
    for key in ("implication bdd", "implication"):
        with_imp, _ = s.transform_to_BDDs({"split bdds": {"split length": 2, key: True}})
        assert (split.nr_bdds, with_imp.nr_bdds) == (4, 5)
        assert np.array_equal(with_imp.instrs[: split.nr_nodes, 2], split.instrs[:, 2])
        assert (with_imp.instrs[split.nr_nodes:-2, 2] >= 8).all()


def test_unsupported_keys_and_missing_gpu():
    import torch
    with pytest.raises(RuntimeError, match="no input"):
        drv.bdd_solver({"relaxation solver": "cuda parallel mma"}, log=quiet)
    with pytest.raises(RuntimeError, match="scope"):
        drv.bdd_solver({"input": os.path.join(GOLDEN, "matching_3x3.lp"), "variable order": "bfs"}, log=quiet)
    with pytest.raises(RuntimeError, match="scope"):
        drv.bdd_solver({"input": os.path.join(GOLDEN, "matching_3x3.lp"), "export bdd graph": "x.dot"}, log=quiet)
    with pytest.raises(RuntimeError, match="file extension"):
        drv.bdd_solver({"input": os.path.join(GOLDEN, "matching_3x3.lp"), "export lp": "x.mps"}, log=quiet)
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):                                  # no CPU fallback
            drv.bdd_solver({"input": os.path.join(GOLDEN, "matching_3x3.lp"), "relaxation solver": "cuda parallel mma"}, log=quiet)


@pytest.mark.parametrize("name", ["matching_3x3", "long_mrf_chain"])
def test_export_lp_key_writes_the_problem_back(tmp_path, name):
    """"export lp" (bdd_solver.cpp:412-430) in the Python driver and in the C++ command line: the exported file reads back as the same ILP
    (the export happens before a solver is constructed, so it works without a GPU; the solve itself then fails loudly here)"""
    import subprocess
    import torch
    from bdd_b200 import lp
    src = os.path.join(GOLDEN, name + ".lp")
    want = lp.parse_lp(open(src).read())

    def same(path):
        got = lp.parse_lp(open(path).read())
        assert got.var_names == want.var_names and got.objective == want.objective and got.constant == want.constant
        assert [(k.variables, k.coefficients, k.ineq, k.rhs) for k in got.constraints] == [(k.variables, k.coefficients, k.ineq, k.rhs) for k in want.constraints]

    out, bdd_out = tmp_path / "py.lp", tmp_path / "py_bdd.lp"
    cfg = {"input": src, "relaxation solver": "cuda parallel mma", "export lp": str(out), "export bdd lp": str(bdd_out),
           "termination criteria": {"maximum iterations": 2}}
    try:
        drv.bdd_solver(cfg, log=quiet)
    except RuntimeError:
        assert not torch.cuda.is_available()
    same(out)
    py_bdd_lp = bdd_out.read_text()
    assert py_bdd_lp.startswith("Minimize\n") and "\nR_0: " in py_bdd_lp and py_bdd_lp.endswith("End\n")
    cli = os.path.join(os.path.dirname(GOLDEN), "..", "bdd_b200", "bdd_solver_cl")
    if os.path.exists(cli):
        out = tmp_path / "cpp.lp"
        cfg["export lp"], cfg["export bdd lp"] = str(out), str(tmp_path / "cpp_bdd.lp")
        r = subprocess.run([cli, json.dumps(cfg)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=120)
        assert r.returncode == 0 or not torch.cuda.is_available()
        same(out)
        assert (tmp_path / "cpp_bdd.lp").read_text() == py_bdd_lp          # both drivers export the same relaxation (tests/test_collection.py solves it)
