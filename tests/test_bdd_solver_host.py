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
    with pytest.raises(RuntimeError, match="implication"):
        s.transform_to_BDDs({"split bdds": {"split length": 2, "implication bdd": True}})


def test_unsupported_keys_and_missing_gpu():
    import torch
    with pytest.raises(RuntimeError, match="no input"):
        drv.bdd_solver({"relaxation solver": "cuda parallel mma"}, log=quiet)
    with pytest.raises(RuntimeError, match="scope"):
        drv.bdd_solver({"input": os.path.join(GOLDEN, "matching_3x3.lp"), "variable order": "bfs"}, log=quiet)
    with pytest.raises(RuntimeError, match="scope"):
        drv.bdd_solver({"input": os.path.join(GOLDEN, "matching_3x3.lp"), "export lp": "x.lp"}, log=quiet)
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):                                  # no CPU fallback
            drv.bdd_solver({"input": os.path.join(GOLDEN, "matching_3x3.lp"), "relaxation solver": "cuda parallel mma"}, log=quiet)
