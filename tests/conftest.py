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
    config.addinivalue_line("markers", "slow: full-size configuration that takes minutes (N = 32768)")


def _cuda_device_present() -> bool:
    """True when this machine has an NVIDIA device node.  Deliberately NOT a call into the engine: on
    a GPU box a missing / unloadable libsdpsr_cuda.so must fail the GPU tests loudly, not skip them."""
    return os.path.exists("/dev/nvidiactl") or os.path.exists("/dev/nvidia0")


def pytest_collection_modifyitems(config, items):
    if _cuda_device_present():
        return
    skip = pytest.mark.skip(reason="no CUDA device on this machine (the engine has no CPU fallback)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


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
