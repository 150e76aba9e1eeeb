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


def pytest_collection_modifyitems(config, items):
    """`gpu` tests need a CUDA device and the built library: on a CUDA-less machine a plain `pytest tests`
    skips them instead of failing (the driver selects them explicitly with -m gpu on the B200 box)."""
    try:
        import torch
        have_gpu = torch.cuda.is_available()
    except Exception:
        have_gpu = False
    lib = os.path.join(REPO, "gym_cellular_b200", "lib", "libgymcellular_b200.so")
    if have_gpu and os.path.exists(lib):
        return
    reason = "needs a CUDA device" if not have_gpu else "libgymcellular_b200.so has not been built"
    skip = pytest.mark.skip(reason=reason)
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


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
