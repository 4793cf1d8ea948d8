"""Long-BDD splitting (bdd_b200/split.py) against the reference's own bdd_collection::split_qbdd (oracle/_ref, bit-exact instruction
arrays) and against its meaning: an assignment satisfies the original BDD iff exactly one assignment of the auxiliary variables
makes every chunk accept."""
import itertools

import numpy as np
import pytest

import bindings as B
from bdd_b200 import instances
from bdd_b200.instances import BddCollection, bdds_accept
from bdd_b200.split import split_long_bdds, split_qbdd


def _cases():
    return {
        "assignment": lambda: instances.assignment(12, seed=3)[0],
        "set_cover": lambda: instances.set_cover(m=20, n=40, k=11, seed=2)[0],
        "random": lambda: instances.random_inequalities(25, 30, max_len=12, max_coeff=5, seed=7)[0],
    }


@pytest.mark.skipif(not B.ref_available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("chunk", [2, 3, 5])
@pytest.mark.parametrize("name", ["assignment", "set_cover", "random"])
def test_split_matches_reference_bit_for_bit(name, chunk):
    col = _cases()[name]()
    rc = B.RefCollection.from_arrays(col.instrs, col.delims)
    instrs = np.ascontiguousarray(col.instrs, dtype=np.uint64)
    delims = col.delims.astype(np.int64)
    aux = aux_ref = col.nr_variables()
    base = int(delims[-1])
    ours = []
    n_split = 0
    for b in range(col.nr_bdds):
        try:
            chunks, aux_new = split_qbdd(instrs, int(delims[b]), int(delims[b + 1]), chunk, aux, base)
        except ValueError:
            continue                      # cut in front of a width-1 layer: the reference asserts there (bdd_collection.cpp:598)
        n_new, aux_ref = rc.split_qbdd(b, chunk, aux_ref)
        aux = aux_new
        assert aux == aux_ref
        assert n_new == max(len(chunks), 1)
        ours.extend(chunks)
        base += sum(c.shape[0] for c in chunks)
        n_split += len(chunks) > 1
    assert n_split > 0
    ref_instrs, ref_delims = rc.export()
    mine = np.concatenate([instrs] + ours, axis=0)
    assert ref_instrs.shape == mine.shape
    inner = mine[:, 2] < instances.BOTSINK
    assert np.array_equal(ref_instrs[:, 2], mine[:, 2])
    assert np.array_equal(ref_instrs[inner], mine[inner])          # (lo / hi of sink instructions are not meaningful)
    assert np.array_equal(ref_delims[: col.nr_bdds + 1].astype(np.int64), delims)
    assert int(ref_delims[-1]) == mine.shape[0]


@pytest.mark.parametrize("chunk", [2, 3, 4])
def test_split_preserves_the_constraint(chunk):
    col, _ = instances.random_inequalities(6, 9, max_len=9, max_coeff=4, seed=5)
    n = col.nr_variables()
    for b in range(col.nr_bdds):
        one = col.select(np.array([b]))
        try:
            split, n_all = split_long_bdds(one, chunk, nr_variables=n)
        except ValueError:
            continue
        if split.nr_bdds == 1:
            continue
        n_aux = n_all - n
        if n_aux > 10:
            continue
        used = np.unique(one.instrs[one.instrs[:, 2] < instances.BOTSINK, 2].astype(np.int64))
        for bits in itertools.product((0, 1), repeat=len(used)):
            x = np.zeros(n_all, dtype=np.int8)
            x[used] = bits
            want = bool(bdds_accept(one, x)[0])
            hits = 0
            for aux in itertools.product((0, 1), repeat=n_aux):
                x[n:] = aux
                hits += bool(bdds_accept(split, x).all())
            assert hits == (1 if want else 0)


def test_split_long_bdds_keeps_short_ones():
    col, _ = instances.set_cover(m=30, n=50, k=5, seed=1)
    out, n_all = split_long_bdds(col, 8)
    assert out is col and n_all == col.nr_variables()
    out, n_all = split_long_bdds(col, 3)
    assert out.nr_bdds == 2 * col.nr_bdds and n_all == col.nr_variables() + 2 * col.nr_bdds


def test_compute_split_length_fills_the_gpu():
    import sys
    from bdd_b200.split import compute_split_length
    col, _ = instances.assignment(300, seed=3)                     # 600 BDDs of 300 variables
    length = compute_split_length(col)
    assert 16 <= length < 300
    out, _ = split_long_bdds(col, length)
    assert out.nr_bdds >= 32 * 148 * 16 or length == 16            # enough bundles for 16 warps on every SM (or the floor)
    if length > 16:
        assert split_long_bdds(col, length + 1)[0].nr_bdds < 32 * 148 * 16     # and it is the largest such length
    small, _ = instances.set_cover(m=30, n=50, k=5, seed=1)
    assert compute_split_length(small) == sys.maxsize             # short BDDs are left alone


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["double", "float"])
def test_split_collection_solves_like_the_oracle_on_gpu(precision):
    """The chunk BDDs (head / tail gadgets of varying width) go through the sweep kernels like any other collection: pass-by-pass
    parity with the CPU oracle, and the bound of the split relaxation stays below the optimum of the assignment problem."""
    torch = pytest.importorskip("torch")
    from bdd_b200.solver import bdd_cuda_parallel_mma
    col, costs = instances.assignment(24, seed=3)
    split, n_all = split_long_bdds(col, 7)
    assert split.nr_bdds > col.nr_bdds
    c = np.concatenate([costs, np.zeros(n_all - len(costs))])
    B.oracle_set_num_threads(1)
    s = bdd_cuda_parallel_mma(split, c, precision=precision, deterministic=(precision == "double"))
    o = B.Oracle(split.instrs, split.delims, c, precision)
    tol = 1e-9 if precision == "double" else 1e-4
    for _ in range(15):
        s.iteration(); o.iteration()
        assert abs(s.lower_bound() - o.lower_bound()) <= tol * max(1.0, abs(o.lower_bound()))
    whole = bdd_cuda_parallel_mma(col, costs, precision=precision)
    whole.iterations(600)
    s.iterations(600)
    opt = whole.lower_bound()                      # the assignment relaxation is tight
    assert s.lower_bound() <= opt + 1e-3 * abs(opt)
    assert s.lower_bound() >= opt - 0.05 * abs(opt)
