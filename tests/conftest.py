import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def iris():
    return dict(np.load(os.path.join(GOLDEN, "iris_ppca.npz")))


@pytest.fixture(scope="session")
def gp_golden():
    return dict(np.load(os.path.join(GOLDEN, "gp_example.npz")))


@pytest.fixture(scope="session")
def lib():
    """The CUDA library; GPU tests must run on it -- no fallback."""
    import gplvmhaskell_b200 as g
    return g.load_library()


def rel(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(1e-300, float(np.max(np.abs(b)))))
