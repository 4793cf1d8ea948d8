import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """Tests marked gpu are skipped (not failed) on a host without a CUDA device."""
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Make sure both shared libraries exist (the oracle is test infrastructure; the CUDA
    library cross-compiles without a GPU)."""
    if not os.path.exists(os.path.join(ROOT, "oracle", "liboracle_mma.so")) or not os.path.exists(
            os.path.join(ROOT, "bdd_b200", "libbdd_b200.so")):
        sys.path.insert(0, ROOT)
        import __graft_entry__
        __graft_entry__.build()
    yield


def golden_names():
    return sorted(f[:-4] for f in os.listdir(GOLDEN) if f.endswith(".npz"))
