import os
import sys
from contextlib import contextmanager

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture
def temp_np_seed():
    """Save/restore NumPy's global RNG around a seeded block (the reference's determinism device, tests/conftest.py:12-26)."""

    @contextmanager
    def _seed(seed: int):
        state = np.random.get_state()
        np.random.seed(seed)
        try:
            yield
        finally:
            np.random.set_state(state)

    return _seed


@pytest.fixture(scope="session")
def golden():
    def _load(name: str):
        return np.load(os.path.join(GOLDEN, name + ".npz"))

    return _load
