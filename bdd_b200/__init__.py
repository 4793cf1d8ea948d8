"""bdd_b200 -- B200-native deferred min-marginal averaging over BDD collections.

Drop-in for the reference's `cuda parallel mma` hot path
(LPMP::bdd_cuda_parallel_mma<REAL>, include/bdd_solver/bdd_cuda_parallel_mma.h).
The sweep is hand-written sm_100a CUDA behind the C-ABI in include/bdd_b200.h
(libbdd_b200.so); this package holds the host-side mirror of the reference interface
and the input plumbing around it:

    solver      bdd_cuda_parallel_mma, lbfgs_cuda_mma, run_solver     (ctypes over include/bdd_b200.h)
    bdd_solver  the JSON-config driver (src/bdd_solver/bdd_solver.cpp), also `python -m bdd_b200.bdd_solver`
    collection  bdd_collection, ilp_input                             (ctypes over include/bdd_b200_collection.h)
    dist        constraint-sharded multi-GPU driver, one process per GPU
    lp, instances, split   .lp reader / writer, BDD builders and generators, long-BDD splitting in numpy
  There is no CPU fallback: importing
`bdd_b200.solver` fails loudly if the CUDA library has not been built.
"""
__all__ = ["lp", "instances", "split", "collection", "solver", "bdd_solver", "dist"]
