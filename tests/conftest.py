import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    return np.load(os.path.join(GOLDEN_DIR, "golden_exact.npz"))


@pytest.fixture(scope="session")
def golden_rs_extra():
    return np.load(os.path.join(GOLDEN_DIR, "golden_rs_extra.npz"))


@pytest.fixture(scope="session")
def golden_stats():
    return np.load(os.path.join(GOLDEN_DIR, "golden_stats.npz"))
