"""Host-side logic and the C-ABI library surface.  CPU only (no compute calls)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, golden_names
from bdd_b200 import instances, lp
from bdd_b200 import _lib


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    for header_name, symbols in (("bdd_b200.h", _lib.SYMBOLS), ("bdd_b200_collection.h", _lib.COLLECTION_SYMBOLS)):
        header = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", header_name)).read(), flags=re.S)      # declarations only
        declared = set(re.findall(r"\b(bddb200_[a-z_0-9]+)\s*\(", header))
        declared -= {"bddb200_solver", "bddb200_instruction", "bddb200_status", "bddb200_precision", "bddb200_options"}
        assert declared == set(symbols), declared ^ set(symbols)
        for name in declared:
            assert hasattr(lib, name), name
    assert sorted(os.listdir(os.path.join(ROOT, "include"))) == ["bdd_b200.h", "bdd_b200_collection.h"]
    assert b"sm_100a" in lib.bddb200_version()


def test_every_entry_point_refuses_a_null_handle():
    """all 90-odd entry points that take a solver / L-BFGS / collection / ILP handle, called with NULL for every pointer and 0 for every
    number (in a child process: a crash would take the test runner with it): an error status or a neutral value, never a fault"""
    import subprocess
    import sys
    code = r'''
import ctypes as C, sys
sys.path.insert(0, %r)
from bdd_b200 import _lib
lib = _lib.load()
skip = {"bddb200_default_options", "bddb200_last_error", "bddb200_version", "bddb200_create", "bddb200_plan_shard", "bddb200_create_shard", "bddb200_load",
        "bddb200_layout_stats", "bddb200_delta_exchange", "bddb200_delta_exchange_two_shot", "bddb200_collection_create", "bddb200_ilp_read"}
neutral = {"bddb200_precision_of", "bddb200_device_of", "bddb200_collection_destroy", "bddb200_ilp_destroy"}
n = 0
for name in _lib.SYMBOLS + _lib.COLLECTION_SYMBOLS:
    if name in skip:
        continue
    f = getattr(lib, name)
    args = [None if (t in (C.c_void_p, C.c_char_p) or (isinstance(t, type) and issubclass(t, C._Pointer))) else (0.5 if t is C.c_double else 0) for t in f.argtypes]
    r = f(*args)
    n += 1
    if f.restype is C.c_int and name not in neutral and r == 0:
        sys.exit("accepted a null handle: " + name)
    if f.restype is C.c_size_t and r != 0:
        sys.exit("non-zero size for a null handle: " + name)
print("ok", n)
''' % ROOT
    r = subprocess.run([sys.executable, "-c", code], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.startswith("ok "), r.stdout
    assert int(r.stdout.split()[1]) >= 90


def test_public_headers_are_plain_c(tmp_path):
    """the drop-in boundary is a C ABI: both headers compile as C99 with nothing but the standard library, and a C program links every
    host-side entry point it needs to go from an .lp text to a split BDD collection (no GPU involved)"""
    import subprocess
    src = tmp_path / "use.c"
    src.write_text('''#include <stdio.h>
#include "bdd_b200_collection.h"
int main(void)
{
    bddb200_ilp* ilp; bddb200_collection* col; size_t n_bdds, n_vars, n_split, n_all;
    if(bddb200_ilp_read("Minimize\\n x1 + 2 x2 + x3 + x4 + 3 x5\\nSubject To\\n x1 + x2 + x3 + x4 + x5 = 2\\n x1 + x5 <= 1\\nEnd\\n", &ilp)) { puts(bddb200_last_error()); return 1; }
    if(bddb200_ilp_to_bdds(ilp, &col) || bddb200_ilp_nr_variables(ilp, &n_vars)) { puts(bddb200_last_error()); return 1; }
    if(bddb200_collection_split_long_bdds(col, 2, n_vars, 1, &n_split, &n_all) || bddb200_collection_nr_bdds(col, &n_bdds)) { puts(bddb200_last_error()); return 1; }
    printf("%zu %zu %zu %zu\\n", n_vars, n_split, n_bdds, n_all);
    {   /* the solver straight from the collection: works on a GPU box, fails loudly without a device (no CPU fallback) */
        double costs[5]; bddb200_solver* s = NULL; int rc;
        bddb200_ilp_objective(ilp, costs, NULL);
        rc = bddb200_create_from_collection(col, costs, 5, BDDB200_DOUBLE, NULL, &s);
        if(rc == BDDB200_OK) { double lb; bddb200_iterations(s, 0.5, 50); bddb200_lower_bound(s, &lb); if(lb > 2.0 + 1e-9) return 2; bddb200_destroy(s); }
        else if(rc != BDDB200_ERR_NO_DEVICE && rc != BDDB200_ERR_CUDA) { puts(bddb200_last_error()); return 3; }
    }
    bddb200_collection_destroy(col); bddb200_ilp_destroy(ilp);
    return 0;
}
''')
    exe = tmp_path / "use"
    lib_dir = os.path.dirname(_lib.LIB_PATH)
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I" + os.path.join(ROOT, "include"), "-o", str(exe), str(src),
                        "-L" + lib_dir, "-lbdd_b200", "-Wl,-rpath," + lib_dir], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    r = subprocess.run([str(exe)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=60)
    assert r.returncode == 0, r.stdout
    n_vars, n_split, n_bdds, n_all = map(int, r.stdout.split())
    assert (n_vars, n_split) == (5, 1) and n_bdds == 1 + 3 + 1 and n_all > n_vars          # three chunks and their implication BDD, the short BDD stays


def test_cuda_library_is_sm100a_only():
    """The product library carries sm_100a SASS and no other architecture."""
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], stdout=subprocess.PIPE, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_create_without_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    lib = _lib.load()
    g = np.load(os.path.join(GOLDEN, "matching_3x3.npz"))
    instrs = np.ascontiguousarray(g["instrs"]); delims = np.ascontiguousarray(g["delims"])
    h = C.c_void_p()
    rc = lib.bddb200_create(instrs.ctypes.data, instrs.shape[0], delims.ctypes.data, delims.shape[0] - 1, None, 0, 1, None, C.byref(h))
    assert rc != 0 and not h.value
    assert b"no CPU fallback" in lib.bddb200_last_error() or rc == 2
    with pytest.raises(RuntimeError):
        from bdd_b200.solver import bdd_cuda_parallel_mma
        bdd_cuda_parallel_mma(instances.BddCollection(instrs, delims), g["costs"])


def layout_stats(col, lanes=0):
    lib = _lib.load()
    out = np.zeros(13, dtype=np.uint64)
    instrs = np.ascontiguousarray(col.instrs); delims = np.ascontiguousarray(col.delims)
    rc = lib.bddb200_layout_stats(instrs.ctypes.data, instrs.shape[0], delims.ctypes.data, delims.shape[0] - 1, lanes, out.ctypes.data, 13)
    if rc != 0:
        raise RuntimeError(lib.bddb200_last_error().decode())
    return dict(zip(["slots", "layer_entries", "bundles", "real_nodes", "max_hops", "max_tile", "small_bundles", "ext_layers",
                     "chunks", "stage_small", "stage_large", "lane_bundles", "topo_words"], out.tolist()))


def test_layout_set_cover_is_tight():
    col, _ = instances.set_cover(m=2048, n=4096, k=20, seed=3)
    st = layout_stats(col)
    assert st["real_nodes"] == 2048 * 39
    # width-2 BDDs: one lane per BDD, lane-local class, uniform tiles of 2 rows
    assert st["bundles"] == 2048 // 32 and st["lane_bundles"] == st["bundles"] and st["small_bundles"] == 0
    assert st["max_hops"] == 21
    assert st["slots"] == st["bundles"] * 32 * 42          # 21 hops x 2 rows (root and terminal hop padded)
    assert st["topo_words"] == st["bundles"] * 32 * 21     # one packed topology word per (hop, lane)
    assert st["chunks"] == 0 and st["stage_small"] == 0 and st["stage_large"] == 0   # chunking is a launch parameter there
    assert st["ext_layers"] == 2048 * 21
    # forcing more lanes per BDD trades padding for parallelism (generic class: chunked, per-slot topology)
    st2 = layout_stats(col, lanes=2)
    assert st2["lane_bundles"] == 0 and st2["bundles"] == 2048 // 16 and st2["slots"] == st2["bundles"] * 32 * 21
    assert st2["small_bundles"] == st2["bundles"] and st2["topo_words"] == st2["slots"]
    assert 0 < st2["stage_small"] <= 12 * 1024 and st2["stage_large"] == 0


def test_layout_mixed_lengths_and_widths():
    col, _ = instances.random_inequalities(200, 80, max_len=12, max_coeff=6, seed=11)
    for lanes in (0, 1, 2, 4, 8, 16, 32):
        st = layout_stats(col, lanes)
        assert st["real_nodes"] == col.nr_nodes - 2 * col.nr_bdds
        assert st["slots"] >= st["real_nodes"] + col.nr_bdds
        assert st["ext_layers"] == sum(1 for _ in range(col.nr_bdds)) + len(set()) + st["ext_layers"] - col.nr_bdds  # tautology guard
    with pytest.raises(RuntimeError):
        layout_stats(col, lanes=3)


def test_layout_rejects_non_qbdd():
    g = np.load(os.path.join(GOLDEN, "matching_3x3.npz"))
    bad = g["instrs"].copy()
    bad[0, 0] += np.uint64(2)
    with pytest.raises(RuntimeError, match="QBDD"):
        layout_stats(instances.BddCollection(bad, g["delims"]))


@pytest.mark.parametrize("name", golden_names())
def test_lp_reader_and_qbdd_builder_match_reference(name):
    """The LP reader + QBDD builder reproduce the reference converter's collection for every
    fixture: same variables, same number of BDDs and nodes per BDD, same objective."""
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    ilp = lp.parse_lp(open(os.path.join(GOLDEN, name + ".lp")).read())
    col, costs = instances.from_ilp(ilp)
    assert np.array_equal(costs, g["costs"])
    assert np.array_equal(col.delims, g["delims"])
    # same variable per instruction (node order inside a layer may differ)
    assert np.array_equal(col.instrs[:, 2], g["instrs"][:, 2])
    # round trip through the writer
    ilp2 = lp.parse_lp(lp.write_lp(ilp))
    assert ilp2.objective == ilp.objective and len(ilp2.constraints) == len(ilp.constraints)
    for a, b in zip(ilp.constraints, ilp2.constraints):
        assert (a.variables, a.coefficients, a.ineq, a.rhs) == (b.variables, b.coefficients, b.ineq, b.rhs)


def test_lp_reader_details():
    txt = """\\ comment
Minimize
 2 x1 - x2 + 3.5 y_3 - z + 4
Subject To
 c1: x1 + x2 + y_3 >= 1
 - x1 + 2 z <= 1
 named[2]: x2 - z = 0
Bounds
 x1 <= 1
Binaries
 x1 x2
End
"""
    ilp = lp.parse_lp(txt)
    assert ilp.var_names == ["x1", "x2", "y_3", "z"]
    assert ilp.objective == [2.0, -1.0, 3.5, -1.0] and ilp.constant == 4.0
    assert [(c.identifier, c.variables, c.coefficients, c.ineq, c.rhs) for c in ilp.constraints] == [
        ("c1", [0, 1, 2], [1, 1, 1], lp.GE, 1), ("", [0, 3], [-1, 2], lp.LE, 1), ("named[2]", [1, 3], [1, -1], lp.EQ, 0)]


REF_PARSER_CASES = [      # test/test_ILP_parser.cpp:8-27 of the reference: ILP_example, ILP_example_hash
    ("Minimize\nx1 + 2*x2 + 1.5 * x3 - 0.5*x4 - x5\nSubject To\nx1 + 2*x2 + 3 * x3 - 5*x4 - x5 >= 1\nBounds\n x1 <= 1\n x2 >= 0\nEnd",
     ["x1", "x2", "x3", "x4", "x5"], ""),
    ("Minimize\nx1 + 2*x2 + 1.5 * x#3 - 0.5*x#4 - x3\nSubject To\n b_cuta_;0;14_1;@41d: x1 + 2*x2 + 3 * x#3 - 5*x#4 - x3 >= 1\nBounds\n x1 <= 1\n x2 >= 0\nEnd",
     ["x1", "x2", "x#3", "x#4", "x3"], "b_cuta_;0;14_1;@41d"),
]


@pytest.mark.parametrize("text, names, identifier", REF_PARSER_CASES)
def test_readers_pass_the_reference_parser_known_answers(text, names, identifier):
    """the reference's own parser test (test/test_ILP_parser.cpp:38-79): five variables with these names and objective coefficients, one
    constraint -- through the Python reader, the C++ reader behind the C ABI and the pybind11 module"""
    import glob
    import sys
    from bdd_b200.collection import ilp_input
    want_constraint = ([0, 1, 2, 3, 4], [1, 2, 3, -5, -1], lp.GE, 1)
    ilp = lp.parse_lp(text)
    assert ilp.var_names == names and ilp.objective == [1.0, 2.0, 1.5, -0.5, -1.0] and len(ilp.constraints) == 1
    c = ilp.constraints[0]
    assert (c.identifier, c.variables, c.coefficients, c.ineq, c.rhs) == (identifier,) + want_constraint
    native = ilp_input(text)
    assert native.variable_names() == names and native.objective()[0].tolist() == [1.0, 2.0, 1.5, -0.5, -1.0]
    assert native.nr_constraints() == 1 and native.constraint(0) == want_constraint
    pkg = os.path.join(ROOT, "bdd_b200")
    if glob.glob(os.path.join(pkg, "ILP_instance_py*.so")):
        sys.path.insert(0, pkg)
        import ILP_instance_py as ip
        bound = ip.parse_ILP(text)
        assert [bound.get_var_name(v) for v in range(bound.nr_variables())] == names and bound.objective() == [1.0, 2.0, 1.5, -0.5, -1.0]
        assert bound.nr_constraints() == 1 and bound.constraint(0)[0] == identifier and bound.constraint(0)[1:3] == want_constraint[:2]


def test_generator_sizes():
    t = instances.qbdd_template([1] * 20, lp.GE, 1)
    assert len(t.layer) + 2 == 41                       # 2k-1 inner nodes + 2 terminals
    t = instances.qbdd_template([1] * 5, lp.EQ, 1)
    assert len(t.layer) + 2 == 2 * 5 + 1                # simplex_constraint, bdd_collection.cpp:2059
    assert instances.qbdd_template([1, 1], lp.LE, 5) is None
    with pytest.raises(ValueError):
        instances.qbdd_template([1, 1], lp.GE, 3)
    col, costs = instances.qap(n=5)
    assert col.nr_bdds == 2 * 5 + 5 * 4 * 5 and costs.shape[0] == 25 + 10 * 20
    col, costs = instances.grid_mrf(4, 3, 3)
    ne = 3 * 3 + 4 * 2
    assert col.nr_bdds == 12 + ne + 2 * 3 * ne and costs.shape[0] == 12 * 3 + ne * 9
    col, costs = instances.assignment(7)
    assert col.nr_bdds == 14 and col.nr_nodes == 14 * 15
    sub = col.select([3, 9])
    assert sub.nr_bdds == 2 and sub.nr_nodes == 30 and int(sub.instrs[:13, :2].max()) < 15
