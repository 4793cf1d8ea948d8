"""TEST INFRASTRUCTURE ONLY -- numpy restatement of one perturbation round of the reference's GPU primal rounding
(src/bdd_solver/incremental_mm_agreement_rounding_cuda.cu): mm_diff_direction_func :28-42, compute_mm_types :79-110
(fill_mm_type_func :44-67), compute_mm_sums :124-146, mm_types_transform :148-212.  Input: min-marginals sorted by
variable (min_marginals_cuda(true)).  The random number of a variable is minstd_rand (thrust::default_random_engine)
discarded by (variable + round) -- the reference discards by (thread id + round), which is an implementation detail of
thrust's launch shape; the distribution is the same.  Parity of the deterministic part (types, sums, which side gets the
perturbation) is exact; no reference test or golden vector covers this file ("next" row of SURVEY 8f)."""
import numpy as np

ZERO, ONE, EQUAL, INCONSISTENT = 0, 1, 2, 3
INT_MAX = 2 ** 31 - 1


def classify(mm_vars, mm0, mm1, nr_vars):
    """per-variable type and (sum mm0, sum mm1)"""
    keep = mm_vars != INT_MAX
    v, m0, m1 = mm_vars[keep].astype(np.int64), mm0[keep], mm1[keep]
    direction = np.where(m0.astype(np.float64) + 1e-6 <= m1.astype(np.float64), -1, np.where(m1.astype(np.float64) + 1e-6 <= m0.astype(np.float64), 1, 0))
    mn = np.full(nr_vars, 2); mx = np.full(nr_vars, -2)
    np.minimum.at(mn, v, direction); np.maximum.at(mx, v, direction)
    types = np.full(nr_vars, INCONSISTENT)
    types[(mx == 0) & (mn == 0)] = EQUAL
    types[mx < 0] = ZERO
    types[mn > 0] = ONE
    types[mn == 2] = ZERO
    s0 = np.zeros(nr_vars, dtype=mm0.dtype); s1 = np.zeros(nr_vars, dtype=mm0.dtype)
    for i in range(v.shape[0]):          # sequential, BDD order: the summation order of the device kernel
        s0[v[i]] += m0[i]; s1[v[i]] += m1[i]
    return types, s0, s1, mn


def minstd_after(n):
    return pow(48271, int(n) + 1, 2147483647)


def perturbation(types, s0, s1, mn, delta, round_index, dtype):
    n = types.shape[0]
    d0 = np.zeros(n, dtype=dtype); d1 = np.zeros(n, dtype=dtype)
    for v in range(n):
        t = types[v]
        if t == ONE:
            d0[v] = delta
        elif t == ZERO:
            if mn[v] != 2:
                d1[v] = delta
        else:
            x = minstd_after(v + round_index)
            u = np.float32(np.float32(x - 1) / np.float32(2147483646.0))
            r = np.float32(np.float32(-delta) + u * np.float32(2.0 * delta))
            amount = dtype(np.float64(np.float32(abs(r))) * np.float64(delta))      # float |r| times double delta, rounded to REAL
            if t == EQUAL:
                if r < 0:
                    d0[v] = amount
                else:
                    d1[v] = amount
            else:
                if s0[v] < s1[v]:
                    d1[v] = amount
                else:
                    d0[v] = amount
    return d0, d1
