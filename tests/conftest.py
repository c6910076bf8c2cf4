import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name))


@pytest.fixture(scope="session")
def bunny():
    return load_golden("bunny.npz")


@pytest.fixture(scope="session")
def syn1500():
    return load_golden("synthetic1500.npz")


@pytest.fixture(scope="session")
def nonrigid_golden():
    return load_golden("nonrigid.npz")
