"""The numpy restatement of the rounding classification (oracle/rounding_oracle.py) against the reference's own CPU decoder
(include/mm_primal_decoder.h, src/bdd_solver/mm_primal_decoder.cpp compiled into oracle/_ref/libbdd_ref.so): agreement type of every
variable, min-marginal sums, type statistics and the reconstructed solution.  Runs without a GPU."""
import numpy as np
import pytest

import bindings as B
import rounding_oracle as R

INT_MAX = 2 ** 31 - 1
pytestmark = pytest.mark.skipif(not B.ref_available(), reason="oracle/_ref/libbdd_ref.so not built")


def random_mms(rng, n_vars, decided):
    """min-marginals per variable: integer-valued gaps (as the integer-cost fixtures give), a mix of all four agreement types"""
    out = []
    for v in range(n_vars):
        k = int(rng.integers(1, 7))
        base = rng.integers(-20, 20, size=(k, 1)).astype(np.float64)
        kind = rng.integers(0, 2) if decided else rng.integers(0, 4)
        if kind == 0:
            gap = rng.integers(1, 9, size=(k, 1)).astype(np.float64)             # zero: mm0 < mm1 everywhere
        elif kind == 1:
            gap = -rng.integers(1, 9, size=(k, 1)).astype(np.float64)            # one
        elif kind == 2:
            gap = np.zeros((k, 1))                                                # equal
        else:
            gap = rng.integers(-5, 6, size=(k, 1)).astype(np.float64)            # whatever comes: mostly inconsistent
        out.append(np.concatenate([base, base + gap], axis=1))
    return out


@pytest.mark.parametrize("decided", [False, True])
@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_classification_equals_the_reference_decoder(seed, decided):
    rng = np.random.default_rng(seed)
    n_vars = 400
    mms = random_mms(rng, n_vars, decided)
    types_r, sums_r, stats_r, sol_r = B.ref_mm_decode(mms)
    # the sorted (variable, BDD) layout min_marginals_cuda(true) returns, terminal entries last
    mm_vars = np.concatenate([np.full(m.shape[0], v) for v, m in enumerate(mms)] + [np.full(3, INT_MAX)])
    mm0 = np.concatenate([m[:, 0] for m in mms] + [np.zeros(3)])
    mm1 = np.concatenate([m[:, 1] for m in mms] + [np.zeros(3)])
    types, s0, s1, mn = R.classify(mm_vars, mm0, mm1, n_vars)
    assert np.array_equal(types, types_r)
    assert np.array_equal(s0, sums_r[:, 0]) and np.array_equal(s1, sums_r[:, 1])
    assert stats_r == {"one": int((types == R.ONE).sum()), "zero": int((types == R.ZERO).sum()),
                       "equal": int((types == R.EQUAL).sum()), "inconsistent": int((types == R.INCONSISTENT).sum())}
    if decided:
        assert sol_r is not None and np.array_equal(sol_r, (types == R.ONE).astype(np.int8))
    # which side the perturbation goes to is a function of type and sums for everything but the `equal` type
    d0, d1 = R.perturbation(types, s0, s1, mn, 0.5, 3, np.float64)
    for v in range(n_vars):
        if types_r[v] == 1:
            assert d0[v] == 0.5 and d1[v] == 0
        elif types_r[v] == 0:
            assert d0[v] == 0 and d1[v] == 0.5
        elif types_r[v] == 3:
            hi_side = sums_r[v, 0] < sums_r[v, 1]      # incremental_mm_agreement_rounding.hxx:131-140
            assert (d1[v] > 0 and d0[v] == 0) if hi_side else (d0[v] > 0 and d1[v] == 0)
        else:
            assert (d0[v] > 0) != (d1[v] > 0)
