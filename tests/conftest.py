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


class Coeffs:
    """Deterministic source of the random coefficient vectors.  The oracle and
    the GPU path each get their own instance with the same seed, so both see
    identical values in the reference's draw order."""

    def __init__(self, seed=20260101):
        self.rng = np.random.default_rng(seed)
        self.draws = []

    def __call__(self, n):
        r = self.rng.random(int(n))
        self.draws.append(int(n))
        return r


@pytest.fixture
def coeffs():
    return Coeffs


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
