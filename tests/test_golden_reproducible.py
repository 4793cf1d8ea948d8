"""The committed golden vectors are what the generating script produces from the reference TODAY: tests/golden/make_golden.py run into
a scratch directory (it needs /root/reference for the LP strings of the reference's tests and oracle/_ref for the reference's solver)
gives byte-identical arrays and LP texts.  CPU only; skipped where the reference is not present."""
import glob
import os
import sys

import numpy as np
import pytest

import bindings as B
from conftest import GOLDEN


@pytest.mark.skipif(not (os.path.isdir("/root/reference/test") and B.ref_available()), reason="needs /root/reference and oracle/_ref")
def test_golden_vectors_regenerate_identically(tmp_path, capsys):
    sys.path.insert(0, GOLDEN)
    import make_golden
    old = make_golden.OUT
    make_golden.OUT = str(tmp_path)
    try:
        make_golden.main()
    finally:
        make_golden.OUT = old
    capsys.readouterr()
    made = sorted(os.path.basename(f) for f in glob.glob(os.path.join(str(tmp_path), "*")))
    kept = sorted(os.path.basename(f) for f in glob.glob(os.path.join(GOLDEN, "*")) if f.endswith((".npz", ".lp", ".json")))
    assert made == kept
    for name in made:
        a, b = os.path.join(str(tmp_path), name), os.path.join(GOLDEN, name)
        if name.endswith(".npz"):
            x, y = np.load(a), np.load(b)
            assert set(x.files) == set(y.files), name
            for k in x.files:
                assert np.array_equal(x[k], y[k]), (name, k)
        else:
            assert open(a).read() == open(b).read(), name
