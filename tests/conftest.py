import os
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)
GOLDEN = os.path.join(REPO, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_pol():
    return np.load(os.path.join(GOLDEN, "polarisation.npz"))


@pytest.fixture(scope="session")
def golden_gw():
    return np.load(os.path.join(GOLDEN, "gridworld.npz"))


@pytest.fixture(scope="session")
def golden_dbg():
    return np.load(os.path.join(GOLDEN, "debug.npz"))


def detab(idx, n_cells, radix):
    """Little-endian digits of idx, as int8 [n_cells][len(idx)] (cell-major batch layout)."""
    idx = np.asarray(idx, np.int64)
    out = np.zeros((n_cells, idx.shape[0]), np.int8)
    for c in range(n_cells):
        out[c] = idx % radix
        idx = idx // radix
    return out
