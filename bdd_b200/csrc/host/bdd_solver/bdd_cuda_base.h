// Replaces include/bdd_solver/bdd_cuda_base.h of the reference: LPMP::bdd_cuda_base<REAL> is defined
// together with its subclass in bdd_cuda_parallel_mma.h (both are thin wrappers over libbdd_b200.so).
#pragma once
#include "bdd_solver/bdd_cuda_parallel_mma.h"
