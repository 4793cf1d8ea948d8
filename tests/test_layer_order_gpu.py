"""reference_layer_order(): the permutation between this package's BDD-major per-layer vectors and the reference's hop-sorted ones
(bdd_cuda_base.cu:146-188, :240-285).  The pure function is checked on the CPU against a hand-worked example and an independent
node-level restatement of the reference's sort; on the GPU against the reference's own CUDA solver (oracle/_ref/libbdd_ref_cuda.so):
its get_primal_variable_index / get_bdd_index and its get_solver_costs before and after iterations."""
import numpy as np
import pytest

import bindings as B
from bdd_b200 import instances
from bdd_b200.solver import reference_layer_order

INT_MAX = 2 ** 31 - 1


def bdd_major_indices(col):
    """per-layer (variable, BDD) in BDD-major order with one terminal layer per BDD (include/bdd_b200.h, "Layer order")"""
    primal, bdd = [], []
    for b in range(col.nr_bdds):
        idx = col.instrs[int(col.delims[b]):int(col.delims[b + 1]) - 2, 2].astype(np.int64)
        layers = idx[np.concatenate([[True], idx[1:] != idx[:-1]])]
        primal += layers.tolist() + [INT_MAX]
        bdd += [b] * (len(layers) + 1)
    return np.asarray(primal), np.asarray(bdd)


def node_level_reference_order(col):
    """bdd_cuda_base.cu:95-135 + :146-172 + :240-262 spelled out on nodes: hop distance per node, sort of the nodes by
    (hop, variable, BDD), unique keys = layers"""
    keys = []
    for b in range(col.nr_bdds):
        first, last = int(col.delims[b]), int(col.delims[b + 1])
        hop, prev = 0, int(col.instrs[first, 2])
        for i in range(first, last):
            var = int(col.instrs[i, 2])
            if var != prev:
                prev = var
                if i - first <= last - first - 2:
                    hop += 1
            keys.append((hop, var if var < int(instances.BOTSINK) else INT_MAX, b))
    return sorted(set(keys))


def test_hand_worked_example():
    # BDD 0 over variables (3, 5), BDD 1 over (5,), BDD 2 over (0, 5, 7): BDD-major layers with terminal entries
    primal = np.array([3, 5, INT_MAX, 5, INT_MAX, 0, 5, 7, INT_MAX])
    bdd = np.array([0, 0, 0, 1, 1, 2, 2, 2, 2])
    perm = reference_layer_order(primal, bdd)
    # hop 0: (0, bdd 2), (3, bdd 0), (5, bdd 1); hop 1: (5, bdd 0), (5, bdd 2), terminal of bdd 1; hop 2: (7, bdd 2), terminal of bdd 0; hop 3: terminal of bdd 2
    assert perm.tolist() == [5, 0, 3, 1, 6, 4, 7, 2, 8]
    assert reference_layer_order(np.zeros(0), np.zeros(0)).shape == (0,)


@pytest.mark.parametrize("make", [lambda: instances.set_cover(m=40, n=60, k=7, seed=1)[0], lambda: instances.assignment(6, seed=2)[0],
                                  lambda: instances.random_inequalities(30, 25, max_len=9, max_coeff=4, seed=3)[0]])
def test_permutation_equals_a_node_level_restatement_of_the_reference_sort(make):
    col = make()
    primal, bdd = bdd_major_indices(col)
    perm = reference_layer_order(primal, bdd)
    assert sorted(perm.tolist()) == list(range(len(primal)))
    want = node_level_reference_order(col)
    assert [(int(primal[k]), int(bdd[k])) for k in perm] == [(v, b) for _, v, b in want]


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["double", "float"])
def test_permutation_maps_onto_the_reference_cuda_solver(precision):
    if not B.ref_cuda_available():
        pytest.skip("oracle/_ref/libbdd_ref_cuda.so not built")
    pytest.importorskip("torch")
    from bdd_b200.solver import bdd_cuda_parallel_mma
    for col, costs in (instances.set_cover(m=300, n=500, k=9, seed=4), instances.random_inequalities(200, 120, max_len=10, max_coeff=4, seed=6),
                       instances.assignment(12, seed=5)):
        s = bdd_cuda_parallel_mma(col, costs, precision=precision, deterministic=True)
        r = B.RefCudaSolver(col.instrs, col.delims, costs, precision)
        perm = s.reference_layer_order()
        ref_primal, ref_bdd = r.layer_indices()
        assert r.nr_layers() == s.nr_layers() == perm.shape[0]
        assert np.array_equal(s.get_primal_variable_index()[perm], ref_primal)
        assert np.array_equal(s.get_bdd_index()[perm], ref_bdd)
        tol = 1e-9 if precision == "double" else 1e-4
        for n_iter in (0, 3):
            for _ in range(n_iter):
                s.iteration(); r.iteration()
            mine = [t.cpu().numpy().astype(np.float64)[perm] for t in s.get_solver_costs()]
            theirs = r.get_solver_costs()
            inner = ref_primal != INT_MAX                                  # terminal entries carry no cost (bdd_cuda_base.cu:1270-1271)
            for a, b in zip(mine, theirs):
                assert np.allclose(a[inner], b[inner], rtol=tol, atol=tol * max(1.0, float(np.abs(b[inner]).max())))
