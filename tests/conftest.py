import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def insel_sift():
    return dict(np.load(os.path.join(GOLDEN, "insel_sift.npz")))


@pytest.fixture(scope="session")
def insel_orb():
    return dict(np.load(os.path.join(GOLDEN, "insel_orb.npz")))


@pytest.fixture(scope="session")
def synthetic_cv2():
    return dict(np.load(os.path.join(GOLDEN, "synthetic_cv2.npz")))


@pytest.fixture(scope="session")
def sfm():
    """The product: ctypes binding over the C-ABI library (fails loudly if it is not built)."""
    import __graft_entry__ as ge
    return ge.load_package()
