"""The host-side BDD collection (include/bdd_b200_collection.h, bdd_b200/csrc/host/bdd_collection.hpp, bdd_b200/collection.py) against the
reference's own BDD::bdd_collection compiled into oracle/_ref: direct generators, relabelling, reorder, make_qbdd, bdd_and, remove and
split_qbdd with the implication BDD -- instruction arrays bit for bit -- and against the meaning of each operation by enumeration.
CPU only."""
import itertools
import os

import numpy as np
import pytest

import bindings as B
from conftest import GOLDEN, golden_names
from bdd_b200 import instances
from bdd_b200.collection import bdd_collection
from bdd_b200.instances import BddCollection, bdds_accept

needs_ref = pytest.mark.skipif(not B.ref_available(), reason="oracle/_ref not built")


def from_rows(rows) -> BddCollection:
    """rows of (coefficients, variables, relation 0 '<=' / 1 '>=' / 2 '=', right-hand side)"""
    from types import SimpleNamespace
    return instances.from_constraints([SimpleNamespace(coefficients=c, variables=v, ineq=i, rhs=r) for c, v, i, r in rows])


def assert_same(mine: bdd_collection, ref: "B.RefCollection"):
    m = mine.export()
    r_instrs, r_delims = ref.export()
    assert np.array_equal(m.delims, r_delims)
    assert np.array_equal(m.instrs[:, 2], r_instrs[:, 2])
    inner = m.instrs[:, 2] < instances.BOTSINK
    assert np.array_equal(m.instrs[inner], r_instrs[inner])              # lo / hi of the sink instructions carry no meaning


def both():
    return bdd_collection(), B.RefCollection()


# ------------------------------------------------------------------------------------------------ generators
@needs_ref
def test_generators_match_the_reference_bit_for_bit():
    mine, ref = both()
    for n in (1, 2, 3, 7, 40):
        assert mine.simplex_constraint(n) == ref.simplex_constraint(n)
        assert mine.not_all_false_constraint(n) == ref.not_all_false_constraint(n)
    for n in (2, 3, 7, 40):
        assert mine.all_equal_constraint(n) == ref.all_equal_constraint(n)
    for n, k in [(2, 0), (2, 1), (2, 2), (5, 0), (5, 1), (5, 2), (5, 3), (5, 5), (9, 4), (30, 7), (30, 29)]:
        assert mine.cardinality_constraint(n, k) == ref.cardinality_constraint(n, k)
    assert_same(mine, ref)


@pytest.mark.parametrize("n", [1, 2, 3, 6])
def test_generators_mean_what_they_say(n):
    col = bdd_collection()
    kinds = [("simplex", col.simplex_constraint(n), lambda s: s == 1), ("some", col.not_all_false_constraint(n), lambda s: s >= 1)]
    if n > 1:
        kinds.append(("equal", col.all_equal_constraint(n), lambda s: s in (0, n)))
        for k in range(n + 1):
            kinds.append((f"card{k}", col.cardinality_constraint(n, k), lambda s, k=k: s == k))
    for x in itertools.product((0, 1), repeat=n):
        for name, b, want in kinds:
            assert col.evaluate(b, x) == want(sum(x)), (name, x)
    for name, b, _ in kinds:
        assert col.is_qbdd(b) == (name != "some" or n == 1), name          # arcs may skip layers only into the bot sink


def test_generator_arguments_are_checked():
    from bdd_b200._lib import BddB200Error
    col = bdd_collection()
    for call in (lambda: col.simplex_constraint(0), lambda: col.all_equal_constraint(1), lambda: col.cardinality_constraint(3, 4),
                 lambda: col.negate(0), lambda: col.bdd_and([0])):
        with pytest.raises(BddB200Error):
            call()
    assert col.nr_bdds() == 0


@needs_ref
@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_add_linear_constraint_builds_the_bdd_of_the_reference_converter(seed):
    """the direct builder against the reference's lineq_bdd -> bdd_mgr -> add_bdd -> reorder -> make_qbdd -> rebase chain
    (bdd_preprocessor.cpp:175-228): same BDDs emitted or skipped, same layers, same widths, same function; node order inside a layer
    is the one thing the BDD manager decides differently.  And bit for bit against the Python builder (bdd_b200/instances.py)."""
    rng = np.random.default_rng(seed)
    mine, ref, rows = bdd_collection(), B.RefCollection(), []
    for _ in range(40):
        n = int(rng.integers(1, 11))
        variables = sorted(rng.choice(30, size=n, replace=False).tolist())
        coeffs = [int(c) for c in rng.integers(-4, 6, size=n) if True]
        coeffs = [c if c != 0 else 1 for c in coeffs]
        ineq = int(rng.integers(0, 3))
        lo, hi = sum(c for c in coeffs if c < 0), sum(c for c in coeffs if c > 0)
        rhs = int(rng.integers(lo, hi + 1))
        try:
            want = ref.add_constraint(coeffs, variables, ineq, rhs)
        except RuntimeError:
            with pytest.raises(Exception):
                mine.add_linear_constraint(coeffs, variables, ineq, rhs)
            continue
        got = mine.add_linear_constraint(coeffs, variables, ineq, rhs)
        assert (got is None) == (want == -1), (coeffs, variables, ineq, rhs)
        if got is not None:
            assert got == want
            rows.append((coeffs, variables, ineq, rhs))
    m = mine.export()
    r_instrs, r_delims = ref.export()
    sinks_alike = lambda idx: np.where(idx >= instances.BOTSINK, instances.TOPSINK, idx)       # the reference's add_bdd / make_qbdd put the top sink first
    assert m.nr_bdds > 10 and np.array_equal(m.delims, r_delims) and np.array_equal(sinks_alike(m.instrs[:, 2]), sinks_alike(r_instrs[:, 2]))
    r = BddCollection(r_instrs, r_delims)
    for _ in range(200):
        x = rng.integers(0, 2, size=30)
        assert np.array_equal(bdds_accept(m, x), bdds_accept(r, x))
        assert np.array_equal(bdds_accept(m, x), [(sum(c * x[v] for c, v in zip(co, va)) <= rhs) if i == 0 else (sum(c * x[v] for c, v in zip(co, va)) >= rhs) if i == 1
                                                  else (sum(c * x[v] for c, v in zip(co, va)) == rhs) for co, va, i, rhs in rows])
    py = from_rows(rows)
    assert np.array_equal(py.delims, m.delims) and np.array_equal(py.instrs, m.instrs)


@pytest.mark.parametrize("name", golden_names())
def test_library_lp_reader_and_builder_equal_the_python_ones(name):
    """bddb200_ilp_* (the C++ reader behind the C ABI) on every fixture, from the file and from its text: variables, objective, every
    constraint and the BDD collection equal those of bdd_b200/lp.py + instances.py, which tests/test_host.py pins against the
    reference's own converter"""
    from bdd_b200 import lp
    from bdd_b200.collection import ilp_input
    path = os.path.join(GOLDEN, name + ".lp")
    want = lp.parse_lp(open(path).read())
    col, costs = instances.from_ilp(want)
    for source in (path, open(path).read()):
        got = ilp_input(source)
        assert got.nr_variables() == len(want.var_names) and got.variable_names() == want.var_names
        obj, const = got.objective()
        assert np.array_equal(obj, costs) and const == want.constant
        assert got.nr_constraints() == len(want.constraints)
        for c, k in enumerate(want.constraints):
            assert got.constraint(c) == (list(k.variables), list(k.coefficients), k.ineq, k.rhs)
        flat = got.to_bdds().export()
        assert np.array_equal(flat.delims, col.delims) and np.array_equal(flat.instrs, col.instrs)


def test_knapsack_rows_beyond_the_direct_builder_are_refused_not_attempted():
    """distinct powers of two make every subset sum different: the builder stops at 2^20 states per layer with a clear message instead
    of running out of memory (the reference's test/hard_ineqs.h rows are of this kind and go through its BDD manager)"""
    import time
    from bdd_b200._lib import BddB200Error
    coeffs, variables = [2 ** k for k in range(40)], list(range(40))
    t = time.time()
    with pytest.raises(BddB200Error, match="partial sums"):
        bdd_collection().add_linear_constraint(coeffs, variables, 0, 2 ** 39)
    with pytest.raises(ValueError, match="partial sums"):
        instances.qbdd_template(coeffs, 0, 2 ** 39)
    assert time.time() - t < 60
    assert bdd_collection().add_linear_constraint(coeffs[:12], variables[:12], 0, 2 ** 11) is not None


@pytest.mark.parametrize("seed", [5, 11])
def test_the_two_lp_readers_agree_on_mutated_files(seed):
    """differential test: every fixture mutated a few characters at a time goes through the Python reader (bdd_b200/lp.py) and the C++
    reader (bddb200_ilp_*); both reject it or both read the same ILP"""
    import random
    from bdd_b200 import lp
    from bdd_b200.collection import ilp_input
    rnd = random.Random(seed)
    alphabet = " \n\t+-*<=>:0123456789.xyzXe_()[]\\/"
    n_same = n_rejected = 0
    for name in golden_names():
        base = open(os.path.join(GOLDEN, name + ".lp")).read()
        for _ in range(60):
            t = base
            for _ in range(rnd.randint(1, 4)):
                pos, kind = rnd.randrange(len(t)), rnd.randrange(4)
                if kind == 0:
                    t = t[:pos] + t[pos + rnd.randint(1, 8):]
                elif kind == 1:
                    t = t[:pos] + rnd.choice(alphabet) + t[pos:]
                elif kind == 2:
                    t = t[:pos] + rnd.choice(alphabet) + t[pos + 1:]
                else:
                    q = rnd.randrange(len(t))
                    t = t[:pos] + t[q:q + rnd.randint(0, 20)] + t[pos:]
            try:
                a = lp.parse_lp(t)
                got_py = (a.var_names, a.objective, a.constant, [(c.variables, c.coefficients, c.ineq, c.rhs) for c in a.constraints])
            except Exception:
                got_py = None
            try:
                b = ilp_input(t)
                obj, const = b.objective()
                got_cpp = (b.variable_names(), obj.tolist(), const, [b.constraint(c) for c in range(b.nr_constraints())])
            except Exception:
                got_cpp = None
            assert (got_py is None) == (got_cpp is None), t
            if got_py is None:
                n_rejected += 1
            else:
                assert got_py == got_cpp, t
                n_same += 1
    assert n_same >= 100 and n_rejected >= 100


def test_library_lp_reader_rejects_what_it_cannot_read():
    from bdd_b200._lib import BddB200Error
    from bdd_b200.collection import ilp_input
    for text in ("Minimize\n x +\nSubject To\n x + y >= \nEnd\n", "Subject To\n c: x + y = 1\nEnd\n", "Minimize\n x\nSubject To\n x + y >= 3\nEnd\n"):
        with pytest.raises(BddB200Error):
            ilp_input(text).to_bdds()


# ------------------------------------------------------------------------------------------------ relabelling, reorder, make_qbdd, remove
@needs_ref
def test_rebase_negate_invert_remove_match_the_reference():
    mine, ref = both()
    for c in (mine, ref):
        a = c.simplex_constraint(5)
        c.rebase(a, [3, 8, 9, 20, 21])
        b = c.not_all_false_constraint(4)
        c.rebase(b, [1, 2, 5, 7])
        c.invert(b, 5)
        d = c.cardinality_constraint(6, 3)
        c.negate(d)
        e = c.all_equal_constraint(3)
        c.invert(e, 0)
    assert_same(mine, ref)
    assert np.array_equal(mine.variables(1), ref.variables(1)) and list(mine.variables(1)) == [1, 2, 5, 7]
    for c in (mine, ref):
        c.remove([0, 2])
    assert_same(mine, ref)
    assert mine.nr_bdds() == 2


@needs_ref
def test_reorder_and_variable_order_match_the_reference():
    """bdd_and writes its result in reverse post-order, which is not layered: reorder has something to do there; a rebase onto
    non-monotone variables makes `variables` fall back to the topological order of the layers (bdd_collection.cpp:1232-1304)"""
    mine, ref = both()
    for c in (mine, ref):
        ops = _operands(c, range(6))
        made = [c.bdd_and(ops[:3]), c.bdd_and(ops[2:]), c.bdd_and(ops)]
        s = c.simplex_constraint(5)
        c.rebase(s, [7, 3, 9, 1, 4])
        k = c.cardinality_constraint(6, 2)
        c.rebase(k, [5, 4, 30, 2, 1, 0])
        x = c.not_all_false_constraint(4)
        c.rebase(x, [9, 2, 6, 3])
        made += [s, k, x, c.make_qbdd(x)]
    moved = 0
    for b in made:
        assert np.array_equal(mine.variables(b), ref.variables(b))
        assert mine.is_reordered(b) == ref.is_reordered(b) and mine.is_qbdd(b) == ref.is_qbdd(b)
        moved += not mine.is_reordered(b)
        mine.reorder(b)
        ref.reorder(b)
        assert mine.is_reordered(b)
    assert moved >= 3
    assert list(mine.variables(made[3])) == [7, 3, 9, 1, 4] and list(mine.variables(made[4])) == [5, 4, 30, 2, 1, 0]
    assert_same(mine, ref)


@needs_ref
def test_make_qbdd_matches_the_reference():
    mine, ref = both()
    for c in (mine, ref):
        made = []
        for n in (1, 2, 5, 9):
            made.append(c.not_all_false_constraint(n))
        for n in (2, 3, 8):
            made.append(c.all_equal_constraint(n))
        b = c.cardinality_constraint(7, 0)
        made.append(b)
        x = c.not_all_false_constraint(6)
        c.rebase(x, [2, 3, 11, 12, 40, 41])
        c.invert(x, 11)
        made.append(x)
        made.append(c.simplex_constraint(4))                            # quasi-reduced already: make_qbdd copies it
        ops = _operands(c, range(6))
        made.append(c.bdd_and(ops[:3]))                                 # reduced conjunctions skip layers in many places
        made.append(c.bdd_and(ops[1:]))
        for b in made:
            q = c.make_qbdd(b)
            assert c.is_qbdd(q)
    assert_same(mine, ref)


@pytest.mark.parametrize("n", [2, 4, 6])
def test_make_qbdd_keeps_the_function(n):
    col = bdd_collection()
    for b in (col.not_all_false_constraint(n), col.all_equal_constraint(n), col.cardinality_constraint(n, 0)):
        q = col.make_qbdd(b)
        assert col.is_qbdd(q) and col.is_reordered(q)
        for x in itertools.product((0, 1), repeat=n):
            assert col.evaluate(q, x) == col.evaluate(b, x)


# ------------------------------------------------------------------------------------------------ bdd_and
def _operands(c, which):
    """a handful of BDDs over overlapping ascending variables"""
    made = []
    def add(b, variables, inverted=()):
        c.rebase(b, variables)
        for v in inverted:
            c.invert(b, v)
        made.append(b)
    add(c.simplex_constraint(4), [0, 2, 4, 6])
    add(c.not_all_false_constraint(3), [1, 2, 3], inverted=[2])
    add(c.cardinality_constraint(5, 2), [3, 4, 5, 7, 8])
    add(c.all_equal_constraint(3), [0, 5, 9])
    add(c.not_all_false_constraint(4), [4, 6, 8, 9], inverted=[4])
    add(c.simplex_constraint(3), [7, 8, 9])
    return [made[i] for i in which]


@needs_ref
@pytest.mark.parametrize("which", [(0, 1), (2, 3), (0, 1, 2), (1, 3, 4), (0, 1, 2, 3), (0, 1, 2, 3, 4, 5), (5, 0, 3)])
def test_bdd_and_matches_the_reference(which):
    mine, ref = both()
    nrs = [_operands(c, which) for c in (mine, ref)]
    assert nrs[0] == nrs[1]
    assert mine.bdd_and(nrs[0]) == ref.bdd_and(nrs[1])
    assert_same(mine, ref)


@needs_ref
def test_bdd_and_of_more_operands_than_the_reference_takes_at_once():
    """above 49 operands the reference sorts and batches (bdd_collection.h:506-546); the result is the same canonical BDD"""
    mine, ref = both()
    nrs = []
    for c in (mine, ref):
        mine_nrs = []
        for i in range(60):
            b = c.not_all_false_constraint(3)
            c.rebase(b, [i % 7, 7 + i % 5, 12 + i % 3])
            if i % 2:
                c.invert(b, 7 + i % 5)
            mine_nrs.append(b)
        b = c.simplex_constraint(4)
        c.rebase(b, [0, 3, 9, 14])
        mine_nrs.append(b)
        nrs.append(mine_nrs)
    a, b = mine.bdd_and(nrs[0]), ref.bdd_and(nrs[1])
    m, (r_instrs, r_delims) = mine.export(), ref.export()
    got = m.instrs[int(m.delims[a]):int(m.delims[a + 1])].astype(np.int64) - int(m.delims[a])
    want = r_instrs[int(r_delims[b]):int(r_delims[b + 1])].astype(np.int64) - int(r_delims[b])
    assert np.array_equal(got[:-2, :2], want[:-2, :2]) and np.array_equal(m.instrs[int(m.delims[a]):int(m.delims[a + 1]), 2], r_instrs[int(r_delims[b]):int(r_delims[b + 1]), 2])


def test_bdd_and_is_the_conjunction():
    col = bdd_collection()
    nrs = _operands(col, range(6))
    for which in [(0, 1), (1, 2, 3), (0, 1, 2, 3, 4), (2, 4, 5)]:
        a = col.bdd_and([nrs[i] for i in which])
        for x in itertools.product((0, 1), repeat=10):
            assert col.evaluate(a, x) == all(col.evaluate(nrs[i], x) for i in which)


# ------------------------------------------------------------------------------------------------ split_qbdd with the implication BDD
def _long_bdds():
    """quasi-reduced BDDs whose paths rule out combinations of cut nodes (so that the implication BDD is not trivial) and some that do not"""
    cols = [instances.random_inequalities(10, 16, max_len=14, max_coeff=5, seed=s)[0] for s in (3, 11)]
    cols.append(from_rows([([1] * 9, list(range(9)), 2, 4), ([1] * 10, list(range(10)), 0, 3), ([2, 3, 1, 4, 2, 3, 1, 2, 5, 1], list(range(10)), 1, 9)]))
    cols.append(instances.assignment(9, seed=1)[0])
    return cols


@needs_ref
@pytest.mark.parametrize("chunk", [2, 3, 4])
def test_split_with_implication_bdd_matches_the_reference(chunk):
    n_implication = 0
    for col in _long_bdds():
        mine, ref = bdd_collection(col), B.RefCollection.from_arrays(col.instrs, col.delims)
        aux = aux_ref = col.nr_variables()
        for b in range(col.nr_bdds):
            layers = np.unique(col.instrs[int(col.delims[b]):int(col.delims[b + 1]) - 2, 2]).shape[0]
            widths = np.unique(col.instrs[int(col.delims[b]):int(col.delims[b + 1]) - 2, 2], return_counts=True)[1]
            if layers <= chunk or any(widths[c] <= 1 for c in range(chunk, layers, chunk)):
                continue                      # short, or a cut in front of a width-1 layer (the reference asserts there, bdd_collection.cpp:598)
            new_nrs, aux = mine.split_qbdd(b, chunk, aux, True)
            n_new, aux_ref = ref.split_qbdd_implication(b, chunk, aux_ref)
            assert aux == aux_ref and len(new_nrs) == n_new
            nr_chunks = -(-layers // chunk)
            n_implication += n_new == nr_chunks + 1
        assert_same(mine, ref)
        for b in range(mine.nr_bdds()):
            assert mine.is_qbdd(b)
    assert n_implication > 0 or chunk == 4


@needs_ref
def test_split_with_implication_bdd_matches_the_reference_on_long_bdds():
    """BASELINE config 3b's shape: simplex rows over 1118 variables cut every 64 (17 cuts per BDD, ~290 constraints in the conjunction)"""
    col = instances.assignment(1118, seed=3)[0].select([0, 1, 1118, 2235])
    mine, ref = bdd_collection(col), B.RefCollection.from_arrays(col.instrs, col.delims)
    aux = aux_ref = 1118 * 1118
    for b in range(col.nr_bdds):
        new_nrs, aux = mine.split_qbdd(b, 64, aux, True)
        n_new, aux_ref = ref.split_qbdd_implication(b, 64, aux_ref)
        assert aux == aux_ref and len(new_nrs) == n_new == 18 + 1
    assert_same(mine, ref)


@pytest.mark.parametrize("rows, chunk", [
    ([([1] * 6, list(range(6)), 2, 2)], 2),                             # exactly 2 of 6: three chunks
    ([([1] * 8, list(range(8)), 2, 1)], 2),                             # a simplex over 8: four chunks, three cuts
    ([([2, 1, 3, 1, 2, 1], list(range(6)), 0, 5), ([1, 2, 1, 2, 1, 1, 2], list(range(7)), 1, 4)], 2),
])
def test_implication_bdd_accepts_exactly_the_cut_patterns_of_the_paths(rows, chunk):
    """every assignment of the original variables that the BDD accepts extends in exactly one way to the auxiliary variables such that
    all chunks accept, and the implication BDD accepts that extension; the implication BDD only looks at auxiliary variables"""
    col = from_rows(rows)
    n_orig = col.nr_variables()
    c = bdd_collection(col)
    aux = n_orig
    for b in range(col.nr_bdds):
        n_layers = len(c.variables(b))
        first = c.nr_bdds()
        new_nrs, aux_next = c.split_qbdd(b, chunk, aux, True)
        assert len(new_nrs) == -(-n_layers // chunk) + 1, "these BDDs have non-trivial implications"
        imp = new_nrs[-1]
        assert c.is_qbdd(imp) and all(aux <= v < aux_next for v in c.variables(imp))
        n_aux = aux_next - aux
        assert n_aux <= 12
        for x in itertools.product((0, 1), repeat=n_orig):
            full = np.zeros(aux_next, dtype=np.int8)
            full[:n_orig] = x
            accepted = c.evaluate(b, full)
            extensions = 0
            for y in itertools.product((0, 1), repeat=n_aux):
                full[aux:aux_next] = y
                if all(c.evaluate(k, full) for k in new_nrs[:-1]):
                    extensions += 1
                    assert c.evaluate(imp, full)
            assert extensions == (1 if accepted else 0), x
        aux = aux_next
        assert first + len(new_nrs) == c.nr_bdds()


def test_split_long_bdds_equals_the_python_splitter_without_implication_bdd():
    from bdd_b200.split import split_long_bdds
    for col in _long_bdds():
        for length in (3, 5):
            want, n_vars_want = split_long_bdds(col, length)
            c = bdd_collection(col)
            n_split, n_vars = c.split_long_bdds(length, col.nr_variables(), False)
            got = c.export()
            assert n_vars == n_vars_want and n_split > 0
            assert np.array_equal(got.delims, want.delims) and np.array_equal(got.instrs[:, 2], want.instrs[:, 2])
            inner = got.instrs[:, 2] < instances.BOTSINK
            assert np.array_equal(got.instrs[inner], want.instrs[inner])


def test_split_long_bdds_with_implication_bdd_keeps_the_feasible_set():
    """the split problem has exactly one solution per solution of the original one (the auxiliary variables are functions of the path)"""
    col = from_rows([([1] * 6, list(range(6)), 2, 2), ([1, 2, 1, 2, 1, 2], list(range(6)), 0, 5), ([1, 1], [0, 5], 0, 1)])
    c = bdd_collection(col)
    n_split, n_vars = c.split_long_bdds(2, col.nr_variables(), True)
    assert n_split == 2
    flat = c.export()
    assert flat.nr_bdds == 1 + 2 * (3 + 1)                                # the short BDD, then per long BDD three chunks and the implication BDD
    blocks = []                                                          # auxiliary variables of each group of four BDDs
    for g in range(2):
        vs = np.unique(np.concatenate([c.variables(1 + 4 * g + k) for k in range(4)]))
        blocks.append([int(v) for v in vs if v >= 6])
    assert not set(blocks[0]) & set(blocks[1]) and len(blocks[0]) + len(blocks[1]) == n_vars - 6 and max(map(len, blocks)) <= 8
    for x in itertools.product((0, 1), repeat=6):
        want = bool(bdds_accept(col, x).all())
        full = np.zeros(n_vars, dtype=np.int8)
        full[:6] = x
        n_ext = int(c.evaluate(0, full))
        for g in range(2):
            count = 0
            for y in itertools.product((0, 1), repeat=len(blocks[g])):
                full[blocks[g]] = y
                count += all(c.evaluate(1 + 4 * g + k, full) for k in range(4))
            n_ext *= count
        assert n_ext == (1 if want else 0), x


# ------------------------------------------------------------------------------------------------ solving what the splitter produces
def _cardinality_problem(seed=5, n_vars=24, n_rows=10, row_len=9, k=3):
    """exactly k of row_len variables per row: long BDDs of width up to k + 1 whose cuts have non-trivial implications"""
    rng = np.random.default_rng(seed)
    rows = [([1] * row_len, sorted(rng.choice(n_vars, size=row_len, replace=False).tolist()), 2, k) for _ in range(n_rows)]
    return from_rows(rows), rng.normal(size=n_vars)


def _split_with_implication(col, costs, length):
    c = bdd_collection(col)
    n_split, n_all = c.split_long_bdds(length, len(costs), True)
    flat = c.export()
    assert n_split == col.nr_bdds and flat.nr_bdds == col.nr_bdds * (-(-9 // length) + 1)         # every row: its chunks and one implication BDD
    return flat, np.concatenate([costs, np.zeros(n_all - len(costs))])


def test_oracle_solves_a_collection_with_implication_bdds():
    """the split relaxation with implication BDDs is a relaxation of the same problem: its bound stays below the whole problem's bound
    -- which both reach from below -- and is at least as good as random guessing, i.e. finite and increasing"""
    col, costs = _cardinality_problem()
    flat, c = _split_with_implication(col, costs, 3)
    B.oracle_set_num_threads(1)
    whole, split = B.Oracle(col.instrs, col.delims, costs, "double"), B.Oracle(flat.instrs, flat.delims, c, "double")
    lbs = [split.lower_bound()]
    for _ in range(200):
        whole.iteration(); split.iteration()
        lbs.append(split.lower_bound())
    assert np.isfinite(lbs).all() and lbs[-1] >= lbs[0] and all(b >= a - 1e-9 for a, b in zip(lbs, lbs[1:]))
    assert lbs[-1] <= whole.lower_bound() + 1e-6 * max(1.0, abs(whole.lower_bound()))


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["double", "float"])
@pytest.mark.parametrize("length", [2, 3])
def test_collection_with_implication_bdds_solves_like_the_oracle_on_gpu(precision, length):
    """chunks plus implication BDDs (pass-through nodes, top sink before bot sink) go through the sweep kernels like any other
    collection: iteration-by-iteration parity with the CPU oracle"""
    pytest.importorskip("torch")
    from bdd_b200.solver import bdd_cuda_parallel_mma
    col, costs = _cardinality_problem()
    flat, c = _split_with_implication(col, costs, length)
    B.oracle_set_num_threads(1)
    s = bdd_cuda_parallel_mma(flat, c, precision=precision, deterministic=(precision == "double"))
    o = B.Oracle(flat.instrs, flat.delims, c, precision)
    tol = 1e-9 if precision == "double" else 1e-4
    assert abs(s.lower_bound() - o.lower_bound()) <= tol * max(1.0, abs(o.lower_bound()))
    for _ in range(15):
        s.iteration(); o.iteration()
        assert abs(s.lower_bound() - o.lower_bound()) <= tol * max(1.0, abs(o.lower_bound()))


# ------------------------------------------------------------------------------------------------ the reference's own collection tests
def test_reference_collection_known_answers():
    """test/bdd/test_bdd_collection_layer_widths.cpp, _layer_offsets.cpp, _remove.cpp, _qbdd.cpp of the reference, as they stand"""
    col = bdd_collection()
    for i in range(2, 15):
        b = col.simplex_constraint(i)
        flat = col.export()
        first, last = int(flat.delims[b]), int(flat.delims[b + 1]) - 2
        idx = flat.instrs[first:last, 2]
        starts = np.flatnonzero(np.concatenate([[True], idx[1:] != idx[:-1]]))
        widths = np.diff(np.concatenate([starts, [last - first]]))
        assert len(widths) == i and (widths[1:] == 2).all()                       # layer_widths
        assert starts[0] == 0 and all(starts[v] == 2 * v - 1 for v in range(1, i))   # layer_offsets relative to offset(b)
    col = bdd_collection()
    s3, s4 = col.simplex_constraint(3), col.simplex_constraint(4)
    sizes = np.diff(col.export().delims.astype(np.int64))
    assert sizes[s3] == 2 * 3 - 1 + 2 and sizes[s4] == 2 * 4 - 1 + 2 and col.nr_bdds() == 2
    col.remove(0)
    assert col.nr_bdds() == 1 and np.diff(col.export().delims.astype(np.int64))[0] == 2 * 4 - 1 + 2
    col = bdd_collection()
    for i in range(2, 17):
        b = col.not_all_false_constraint(i)
        q = col.make_qbdd(b)
        sizes = np.diff(col.export().delims.astype(np.int64))
        assert col.is_qbdd(q) and not col.is_qbdd(b) and sizes[b] < sizes[q]


@pytest.mark.parametrize("with_implication_bdd", [False, True])
def test_reference_split_protocol_on_cardinality_constraints(with_implication_bdd):
    """test/bdd/test_bdd_collection_split_qbdd.cpp:14-56: a cardinality constraint cut into chunks (with and without the implication BDD)
    is solved to the sum of its k smallest costs by 200 MMA iterations, like the constraint as a whole; for k = 1 the auxiliary variables
    are two per cut.  The solver is the CPU oracle (pinned to the reference's CPU solver, tests/test_oracle_vs_ref.py)."""
    rng = np.random.default_rng(0)
    B.oracle_set_num_threads(1)
    n_cases = 0
    for i in range(4, 17, 3):
        for k in sorted({1, 2, i // 2, i - 2} - {0}):
            if k >= i - 1:
                continue
            for chunk in sorted({2, 3, i // 2, i - 2}):
                if chunk < 2 or chunk + 1 >= i:
                    continue
                col = bdd_collection()
                card = col.cardinality_constraint(i, k)
                whole = col.export()
                new_nrs, next_aux = col.split_qbdd(card, chunk, i, with_implication_bdd)
                nr_chunks = -(-i // chunk)
                assert len(new_nrs) in (nr_chunks, nr_chunks + 1) and len(new_nrs) > 1
                assert with_implication_bdd or len(new_nrs) == nr_chunks
                if k == 1:
                    assert next_aux == i + 2 * (nr_chunks - 1)
                col.remove(card)
                split = col.export()
                costs = rng.uniform(-10, 10, size=i)
                want = np.sort(costs)[:k].sum()
                o_whole = B.Oracle(whole.instrs, whole.delims, costs, "double")
                o_split = B.Oracle(split.instrs, split.delims, np.concatenate([costs, np.zeros(next_aux - i)]), "double")
                for _ in range(200):
                    o_split.iteration()
                assert abs(o_whole.lower_bound() - want) <= 1e-6, (i, k, chunk)
                assert abs(o_split.lower_bound() - want) <= 1e-6, (i, k, chunk, o_split.lower_bound(), want)
                n_cases += 1
    assert n_cases >= 30


@pytest.mark.parametrize("with_implication_bdd", [False, True])
@pytest.mark.parametrize("name", golden_names())
def test_reference_split_protocol_on_the_test_problems(name, with_implication_bdd):
    """test/bdd/test_bdd_collection_split_qbdd.cpp:58-84: every BDD of a test problem cut at length 2; after 300 iterations the bound
    of the split collection equals the bound of the original one"""
    from bdd_b200 import lp
    ilp = lp.parse_lp(open(os.path.join(GOLDEN, name + ".lp")).read())
    flat, costs = instances.from_ilp(ilp)
    col = bdd_collection(flat)
    aux, removed = len(costs), []
    for b in range(flat.nr_bdds):
        try:
            new_nrs, nxt = col.split_qbdd(b, 2, aux, with_implication_bdd)
        except Exception:
            continue                                   # cut in front of a layer of width 1: the reference asserts (bdd_collection.cpp:598)
        if len(new_nrs) > 1:
            removed.append(b)
            aux = nxt
    col.remove(removed)
    split = col.export()
    B.oracle_set_num_threads(1)
    o_whole = B.Oracle(flat.instrs, flat.delims, costs, "double")
    o_split = B.Oracle(split.instrs, split.delims, np.concatenate([costs, np.zeros(aux - len(costs))]), "double")
    for _ in range(300):
        o_whole.iteration(); o_split.iteration()
    assert abs(o_whole.lower_bound() - o_split.lower_bound()) <= 1e-6 * max(1.0, abs(o_whole.lower_bound())), (len(removed), o_whole.lower_bound(), o_split.lower_bound())


def test_reference_variables_protocol():
    """test/bdd/test_bdd_collection_variables.cpp: variables() of simplex and not-all-false BDDs, also after a rebase onto reversed and
    onto shuffled variables (the order of the layers, not of the indices)"""
    col = bdd_collection()
    for i in range(2, 29):
        order = list(range(i))
        s, n = col.simplex_constraint(i), col.not_all_false_constraint(i)
        assert list(col.variables(s)) == order and list(col.variables(n)) == order
        col.rebase(s, order[::-1]); col.rebase(n, order[::-1])
        assert list(col.variables(s)) == order[::-1] and list(col.variables(n)) == order[::-1]
    rng = np.random.default_rng(1)
    shuffled = np.cumsum(rng.integers(1, 43, size=42))
    rng.shuffle(shuffled)
    b = col.simplex_constraint(42)
    assert list(col.variables(b)) == list(range(42))
    col.rebase(b, shuffled)
    assert list(col.variables(b)) == shuffled.tolist() and col.is_qbdd(b) and col.is_reordered(b)


def test_reference_bdd_and_protocol():
    """test/bdd/test_bdd_collection_and.cpp: the conjunction of the three covering constraints over (0,1,3), (0,2,4), (1,2,5) is the
    function the BDD manager computes for them (here: checked on all 64 assignments, and reduced: no node with lo == hi, no duplicates)"""
    col = bdd_collection()
    parts = []
    for variables in ([0, 1, 3], [0, 2, 4], [1, 2, 5]):
        b = col.not_all_false_constraint(3)
        col.rebase(b, variables)
        parts.append(b)
    both = col.bdd_and(parts)
    for x in itertools.product((0, 1), repeat=6):
        assert col.evaluate(both, x) == ((x[0] or x[1] or x[3]) and (x[0] or x[2] or x[4]) and (x[1] or x[2] or x[5]))
    flat = col.export()
    nodes = flat.instrs[int(flat.delims[both]):int(flat.delims[both + 1]) - 2]
    assert (nodes[:, 0] != nodes[:, 1]).all() and len({tuple(r) for r in nodes.tolist()}) == len(nodes)


# ------------------------------------------------------------------------------------------------ write_bdd_lp
def _relaxation_optimum(lp_text: str) -> float:
    """the exported .lp read back with the package's own reader and solved as a linear programme over [0, 1] by HiGHS (scipy)"""
    from scipy.optimize import linprog
    from scipy.sparse import lil_matrix
    from bdd_b200 import lp
    ilp = lp.parse_lp(lp_text)
    n = len(ilp.var_names)
    A = lil_matrix((len(ilp.constraints), n))
    b = np.zeros(len(ilp.constraints))
    for r, k in enumerate(ilp.constraints):
        assert k.ineq == lp.EQ
        for v, a in zip(k.variables, k.coefficients):
            A[r, v] = a
        b[r] = k.rhs
    res = linprog(np.asarray(ilp.objective), A_eq=A.tocsr(), b_eq=b, bounds=(0, 1), method="highs")
    assert res.status == 0, res.message
    return float(res.fun)


@needs_ref
def test_write_bdd_lp_matches_the_reference_byte_for_byte(tmp_path):
    rng = np.random.default_rng(2)
    cases = []
    for name in golden_names():
        g = np.load(os.path.join(GOLDEN, name + ".npz"))
        cases.append((BddCollection(g["instrs"], g["delims"]), g["costs"]))
    col, _ = instances.random_inequalities(15, 20, max_len=9, max_coeff=4, seed=8)
    cases.append((col, np.concatenate([[1.5, -1.0 / 3.0, 1e-7, 12345678.9, -0.0, 0.0], rng.normal(size=14) * 100])))
    n_equal = 0
    for k, (col, costs) in enumerate(cases):
        path = tmp_path / f"bdd_{k}.lp"
        bdd_collection(col).write_bdd_lp(path, costs)
        want = B.RefCollection.from_arrays(col.instrs, col.delims).write_bdd_lp(costs)
        c = bdd_collection(col)
        ascending = all(np.all(np.diff(c.variables(b).astype(np.int64)) > 0) for b in range(col.nr_bdds))
        got = path.read_text()
        if got == want:
            n_equal += 1
            continue
        # the reference ties the first layer of a BDD to the BDD's smallest variable (header :800), which is wrong when the root branches
        # on another one: only then may the texts differ, and only in the rows that link the layers to the shared variables
        assert not ascending, k
        first_diff = next(i for i, (x, y) in enumerate(zip(got, want)) if x != y)
        last_flow_row = max(want.rindex("\nFC_"), want.rindex("\nR_"))
        assert first_diff > want.index("\n", last_flow_row + 1), k
    assert n_equal >= 5


@pytest.mark.parametrize("name, known", [("matching_3x3", -6.0), ("short_chain_shuffled", 1.0), ("long_chain", -9.0), ("grid_graph_3x3", -8.0)])
def test_exported_relaxation_has_the_published_optimum_and_bounds_the_dual_solver(tmp_path, name, known):
    """the linear programme write_bdd_lp exports is the relaxation itself: solved by an LP solver that knows nothing of BDD sweeps it
    gives the reference's published answers (test/test_bdd_cuda_parallel_mma.cu:197-247), and the MMA bound approaches it from below"""
    if not os.path.exists(os.path.join(GOLDEN, name + ".npz")):
        pytest.skip("fixture not present")
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    col = BddCollection(g["instrs"], g["delims"])
    path = tmp_path / "relaxation.lp"
    bdd_collection(col).write_bdd_lp(path, g["costs"])
    optimum = _relaxation_optimum(path.read_text())
    assert abs(optimum - known) <= 1e-7
    B.oracle_set_num_threads(1)
    o = B.Oracle(col.instrs, col.delims, g["costs"], "double")
    for _ in range(300):
        o.iteration()
        assert o.lower_bound() <= optimum + 1e-9
    assert o.lower_bound() >= optimum - 1e-6


def test_split_relaxation_with_implication_bdds_is_as_tight_as_the_whole():
    """cutting a BDD adds auxiliary variables; with the chunks (and the implication BDD) the LP bound of a single cardinality constraint
    stays the sum of its k smallest costs"""
    import tempfile
    rng = np.random.default_rng(3)
    for n, k, chunk in ((8, 3, 2), (9, 2, 3), (10, 5, 3)):
        costs = np.round(rng.uniform(-10, 10, size=n), 2)          # the file carries six significant digits
        for with_implication_bdd in (False, True):
            c = bdd_collection()
            card = c.cardinality_constraint(n, k)
            _, n_all = c.split_qbdd(card, chunk, n, with_implication_bdd)
            c.remove(card)
            with tempfile.TemporaryDirectory() as d:
                path = os.path.join(d, "split.lp")
                c.write_bdd_lp(path, np.concatenate([costs, np.zeros(n_all - n)]))
                assert abs(_relaxation_optimum(open(path).read()) - np.sort(costs)[:k].sum()) <= 1e-7


@pytest.mark.parametrize("make", [lambda: instances.set_cover(m=25, n=40, k=6, seed=3), lambda: instances.grid_mrf(4, 3, 3, seed=2), lambda: instances.qap(n=4, seed=5),
                                  lambda: instances.assignment(6, seed=1)], ids=["set_cover", "grid_mrf", "qap", "assignment"])
def test_dual_bounds_never_exceed_the_lp_optimum_of_the_relaxation(tmp_path, make):
    """every iterate of the CPU oracle and of its L-BFGS wrapper is a feasible dual point: its bound stays below the optimum of the
    exported linear programme (HiGHS), on the four benchmark shapes in small"""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(B.__file__))))
    from lbfgs_oracle import LbfgsOracle
    col, costs = make()
    costs = np.round(costs, 3)                       # the exported file carries six significant digits
    path = tmp_path / "relaxation.lp"
    bdd_collection(col).write_bdd_lp(path, costs)
    optimum = _relaxation_optimum(path.read_text())
    B.oracle_set_num_threads(1)
    plain = B.Oracle(col.instrs, col.delims, costs, "double")
    wrapped = LbfgsOracle(B.Oracle(col.instrs, col.delims, costs, "double"), init_step_size=1e-3)
    slack = 1e-7 * max(1.0, abs(optimum))
    for _ in range(150):
        plain.iteration(); wrapped.iteration()
        assert plain.lower_bound() <= optimum + slack and wrapped.lower_bound() <= optimum + slack
    assert plain.lower_bound() >= optimum - 0.1 * max(1.0, abs(optimum))          # MMA stops at a fixed point, which need not be the optimum


def test_reference_loose_covering_problem(tmp_path):
    """test/test_loose_covering_problem.cpp:8-59: three covering rows over six unit-cost variables have relaxation value 1.5 (the
    reference reaches it with 200 iterations of its sequential solver, tolerance 1e-4); with the cut x1 + ... + x6 >= 2 it is 2.
    Here: the LP optimum of the exported relaxation, and the bound the parallel-MMA oracle reaches."""
    from bdd_b200 import lp
    base = "Minimize\nx1 + x2 + x3 + x4 + x5 + x6\nSubject To\nx1 + x2 + x4 >= 1\nx1 + x3 + x5 >= 1\nx2 + x3 + x6 >= 1\n"
    for extra, want in (("", 1.5), ("x1 + x2 + x3 + x4 + x5 + x6 >= 2\n", 2.0)):
        col, costs = instances.from_ilp(lp.parse_lp(base + extra + "Bounds\nBinaries\nx1\nx2\nx3\nx4\nx5\nx6\nEnd"))
        path = tmp_path / "cover.lp"
        bdd_collection(col).write_bdd_lp(path, costs)
        assert abs(_relaxation_optimum(path.read_text()) - want) <= 1e-9
        B.oracle_set_num_threads(1)
        o = B.Oracle(col.instrs, col.delims, costs, "double")
        for _ in range(200):
            o.iteration()
        assert o.lower_bound() <= want + 1e-9
        if not extra:
            assert abs(o.lower_bound() - want) <= 1e-4           # with the cut, min-marginal averaging stops at a fixed point below the optimum (5/3)


@pytest.mark.parametrize("n", list(range(3, 14)))
def test_reference_single_inequality_protocol(n):
    """test/test_random_inequality_to_bdd.cpp:14-47: one random inequality as one BDD -- the bound of the solver is then the minimum
    over the feasible assignments (enumeration), to 1e-8.  BDD from the direct builder, bound from the CPU oracle."""
    rng = np.random.default_rng(100 + n)
    tested = 0
    for attempt in range(12):
        coeffs = [int(c) if c != 0 else 1 for c in rng.integers(-5, 6, size=n)]
        ineq = int(rng.integers(0, 3))
        rhs = int(rng.integers(sum(c for c in coeffs if c < 0), sum(c for c in coeffs if c > 0) + 1))
        costs = rng.normal(size=n)
        best = min((float(np.dot(costs, x)) for x in itertools.product((0, 1), repeat=n)
                    if (np.dot(coeffs, x) <= rhs if ineq == 0 else np.dot(coeffs, x) >= rhs if ineq == 1 else np.dot(coeffs, x) == rhs)), default=None)
        c = bdd_collection()
        try:
            nr = c.add_linear_constraint(coeffs, list(range(n)), ineq, rhs)
        except Exception:
            assert best is None                      # infeasible
            continue
        if nr is None:                               # always satisfied: every variable is free
            assert abs(best - np.minimum(costs, 0).sum()) <= 1e-8
            continue
        flat = c.export()
        # variables the constraint does not depend on have no layer: they are free as well
        used = set(flat.instrs[:-2, 2].astype(np.int64).tolist())
        free = sum(min(costs[v], 0.0) for v in range(n) if v not in used)
        o = B.Oracle(flat.instrs, flat.delims, None, "double")
        for v in used:
            o.set_cost(float(costs[v]), v)
        assert abs(o.lower_bound() + free - best) <= 1e-8, (coeffs, ineq, rhs)
        tested += 1
    assert tested >= 4


@pytest.mark.parametrize("n", [1, 2, 5, 8])
def test_reference_utility_protocol(n):
    """test/bdd/test_bdd_collection_utility.cpp: each generator encodes its function and negate turns it into the complement (the
    reference compares with its BDD manager's simplex / all_false / all_equal / cardinality; here: all 2^n assignments)"""
    col = bdd_collection()
    made = [(col.simplex_constraint(n), lambda s: s == 1), (col.not_all_false_constraint(n), lambda s: s >= 1)]
    if n >= 2:
        made.append((col.all_equal_constraint(n), lambda s: s in (0, n)))
        made += [(col.cardinality_constraint(n, k), lambda s, k=k: s == k) for k in range(1, n + 1)]
    for x in itertools.product((0, 1), repeat=n):
        assert all(col.evaluate(b, x) == want(sum(x)) for b, want in made), x
    for b, _ in made:
        col.negate(b)
    for x in itertools.product((0, 1), repeat=n):
        assert all(col.evaluate(b, x) == (not want(sum(x))) for b, want in made), x
