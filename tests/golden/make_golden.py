"""Generate the golden vectors under tests/golden/ from the REFERENCE ITSELF.

Run in the build container only (needs /root/reference and oracle/_ref/libbdd_ref.so):

    make -C oracle ref && python tests/golden/make_golden.py

For every LP instance embedded in the reference's tests
  test/test_problems.h:4-195                       (short/long MRF chains, 3x3 grid)
  test/test_bdd_bipartite_matching_problem.cpp:8-52 (3x3 matchings)
  test/test_bdd_cuda_parallel_mma.cu:9-195          (matching_3x3, chains, grid)
the script
  1. parses the LP text with bdd_b200.lp (the LP strings are read from the reference tree at
     generation time and stored as fixture input),
  2. converts every constraint to a QBDD with the reference's own converter
     (bdd_converter / bdd_collection through oracle/ref_wrap.cpp) and stores the flat
     bdd_collection arrays,
  3. runs the reference's CPU solver bdd_parallel_mma_base<bdd_branch_instruction<double,
     uint16_t>> single-threaded and records
       - the lower bound before any iteration and after each of 200 iterations,
       - following test/test_cuda_parallel_mma.cu:68-101, the per-variable delta after every
         forward_mm and every backward_mm of 10 iterations (un-normalised, fed back as is),
         with the lower bound after each backward pass,
       - the min-marginals of the initial state.
The known answers of the reference tests are stored next to them in expected.json.
"""
import json
import os
import re
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import bindings as B  # noqa: E402
from bdd_b200.lp import parse_lp  # noqa: E402

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))

# known answers published by the reference's tests
EXPECTED = {
    # test/test_bdd_cuda_parallel_mma.cu:241-246, 200 iterations + distribute_delta, tol 1e-12
    "matching_3x3": {"lb": -6.0, "tol": 1e-12, "source": "test/test_bdd_cuda_parallel_mma.cu:241"},
    "short_chain_shuffled": {"lb": 1.0, "tol": 1e-12, "source": "test/test_bdd_cuda_parallel_mma.cu:243"},
    "long_chain": {"lb": -9.0, "tol": 1e-12, "source": "test/test_bdd_cuda_parallel_mma.cu:245"},
    "grid_graph_3x3": {"lb": -8.0, "tol": 1e-12, "source": "test/test_bdd_cuda_parallel_mma.cu:247"},
    # test/test_bdd_bipartite_matching_problem.cpp:40-58 (sequential mma, 20 iterations, 1e-6)
    "matching_3x3_diag": {"lb": -6.0, "tol": 1e-6, "source": "test/test_bdd_bipartite_matching_problem.cpp:38"},
    "matching_3x3_first_row": {"lb": -4.0, "tol": 1e-6, "source": "test/test_bdd_bipartite_matching_problem.cpp:58"},
    # test/test_bdd_small_binary_mrfs.cpp:10-65 (parallel mma, within 0.1)
    "short_mrf_chain": {"lb": 1.0, "tol": 0.1, "source": "test/test_bdd_small_binary_mrfs.cpp"},
    "short_mrf_chain_shuffled": {"lb": 1.0, "tol": 0.1, "source": "test/test_bdd_small_binary_mrfs.cpp"},
    "long_mrf_chain": {"lb": -9.0, "tol": 0.1, "source": "test/test_bdd_small_binary_mrfs.cpp"},
    "mrf_grid_graph_3x3": {"lb": -8.0, "tol": 0.1, "source": "test/test_bdd_small_binary_mrfs.cpp"},
}


def ref_strings(path):
    s = open(path).read()
    return dict(re.findall(r'const\s+(?:std::string|char\s*\*)\s+(\w+)\s*=\s*\n?R"\((.*?)\)";', s, re.S))


def main():
    probs = {}
    probs.update(ref_strings(f"{REF}/test/test_problems.h"))
    probs.update(ref_strings(f"{REF}/test/test_bdd_bipartite_matching_problem.cpp"))
    probs.update(ref_strings(f"{REF}/test/test_bdd_cuda_parallel_mma.cu"))
    B.ref_set_num_threads(1)
    for name, txt in sorted(probs.items()):
        ilp = parse_lp(txt)
        rc = B.RefCollection()
        for c in ilp.constraints:
            rc.add_constraint(c.coefficients, c.variables, c.ineq, c.rhs)
        instrs, delims = rc.export()
        costs = np.asarray(ilp.objective, dtype=np.float64)

        # (a) plain iterations
        rs = B.RefSolver(rc, costs, "double")
        lbs = [rs.lower_bound()]
        for _ in range(200):
            rs.iteration()
            lbs.append(rs.lower_bound())
        rs.distribute_delta()
        lb_after_distribute = rs.lower_bound()

        # (b) pass-by-pass protocol of test/test_cuda_parallel_mma.cu:68-101
        rs2 = B.RefSolver(rc, costs, "double")
        delta = np.zeros(2 * rs2.n_vars, dtype=np.float64)
        deltas_fwd, deltas_bwd, lbs_pass = [], [], []
        for _ in range(10):
            rs2.forward_mm(0.5, delta)
            deltas_fwd.append(delta.copy())
            rs2.backward_mm(0.5, delta)
            deltas_bwd.append(delta.copy())
            lbs_pass.append(rs2.lower_bound())

        # (c) min-marginals of the initial state, (variable, k-th BDD) order
        rs3 = B.RefSolver(rc, costs, "double")
        mm = rs3.min_marginals()
        nbpv = np.asarray([rs3.nr_bdds_of_var(v) for v in range(rs3.n_vars)], dtype=np.int64)

        # (d) float trajectory
        rsf = B.RefSolver(rc, costs, "float")
        lbs_f = [rsf.lower_bound()]
        for _ in range(50):
            rsf.iteration()
            lbs_f.append(rsf.lower_bound())

        np.savez_compressed(
            os.path.join(OUT, name + ".npz"),
            instrs=instrs, delims=delims, costs=costs,
            lbs=np.asarray(lbs), lb_after_distribute_cpu=np.asarray(lb_after_distribute),
            deltas_fwd=np.asarray(deltas_fwd), deltas_bwd=np.asarray(deltas_bwd), lbs_pass=np.asarray(lbs_pass),
            min_marginals=mm, nr_bdds_per_var=nbpv, lbs_float=np.asarray(lbs_f),
        )
        with open(os.path.join(OUT, name + ".lp"), "w") as f:
            f.write(txt.strip() + "\n")
        print(f"{name}: {rs.n_vars} vars, {rs.n_bdds} bdds, {instrs.shape[0]} nodes, lb0={lbs[0]}, lb200={lbs[-1]}, after distribute={lb_after_distribute}")
    with open(os.path.join(OUT, "expected.json"), "w") as f:
        json.dump(EXPECTED, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
