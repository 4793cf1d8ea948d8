"""INTEGRATION.md section 1 as a build: the reference's include tree with its class headers replaced by the drop-in ones, and the
reference's own callers of the class -- the JSON driver bdd_solver.cpp with its std::variant of solver types, the GPU primal rounding,
the lbfgs<> instantiations, the pimpl wrapper bdd_cuda<REAL>, the hybrid CPU + GPU solver -- and the reference's own eight test sources
of the class compiled against it unmodified with nvcc -DWITH_CUDA for sm_100a (oracle/Makefile: integration).  Needs /root/reference; no GPU."""
import os
import subprocess

import pytest

from conftest import ROOT


@pytest.mark.skipif(not os.path.isdir("/root/reference/src/bdd_solver"), reason="the reference sources are not here")
def test_reference_callers_compile_against_the_dropin_headers():
    r = subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "-j8", "integration"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=1200)
    assert r.returncode == 0 and "integration build check: OK (5 reference sources and 8 reference tests" in r.stdout, r.stdout[-3000:]
    overlay = os.path.join(ROOT, "oracle", "_ref", "overlay", "include", "bdd_solver")
    for name, target in (("bdd_cuda_parallel_mma.h", "bdd_b200/csrc/host/bdd_solver/bdd_cuda_parallel_mma.h"), ("bdd_cuda_base.h", "bdd_b200/csrc/host/bdd_solver/bdd_cuda_base.h"),
                         ("lbfgs.h", "bdd_b200/csrc/host/bdd_solver/lbfgs.h"), ("lbfgs_generic.h", "/root/reference/include/bdd_solver/lbfgs.h"),
                         ("bdd_solver.h", "/root/reference/include/bdd_solver/bdd_solver.h")):
        assert os.path.realpath(os.path.join(overlay, name)) == os.path.realpath(os.path.join(ROOT, target)), name
