import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_random():
    return np.load(os.path.join(GOLDEN_DIR, "reference_random.npz"))


@pytest.fixture(scope="session")
def golden_c1():
    return np.load(os.path.join(GOLDEN_DIR, "reference_c1.npz"))


@pytest.fixture(scope="session")
def golden_hand():
    return np.load(os.path.join(GOLDEN_DIR, "reference_handvectors.npz"))
