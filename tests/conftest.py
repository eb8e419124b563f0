import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "pinns-tf2.0_b200")
for p in (ROOT, os.path.join(PKG, "utils")):
    if p not in sys.path:
        sys.path.insert(0, p)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


@pytest.fixture(scope="session")
def lib_built():
    """Build libpinn_b200.so if it is stale/missing (nvcc cross-compiles without a GPU)."""
    sys.path.insert(0, PKG)
    import build as pinn_build
    return pinn_build.build()
