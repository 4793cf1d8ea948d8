#!/bin/bash
# last GPU call of round 2 (budget: under two minutes): the library as rebuilt from two translation units, on hardware --
# smoke() against the oracle, the new GPU tests of the implication-BDD collections, one pass of the parity file
mkdir -p gpurun_out
( timeout 45 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -v "cumulative\|bdd parallel" | tail -4
  timeout 40 python -m pytest tests/test_collection.py tests/test_split.py -m gpu -q -x 2>&1 | grep -v "cumulative\|bdd parallel" | tail -4 ) | tee gpurun_out/r02_verify.log
