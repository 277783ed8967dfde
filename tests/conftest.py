import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


def pytest_collection_modifyitems(config, items):
    """Without a CUDA device every test marked `gpu` is skipped, so that a plain `pytest tests` passes on a CPU-only
    machine; with one they run, and a missing liblsf_b200.so is an error (no fallback)."""
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="needs a CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def literals():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "reference_literals.npz"))


@pytest.fixture(scope="session")
def python_runs():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "reference_python_runs.npz"))


@pytest.fixture(scope="session")
def lsf():
    """The product package with its CUDA library loaded. Without a CUDA device the GPU tests are skipped (so a plain
    `pytest tests` passes on a CPU-only machine); with one, a missing or unloadable liblsf_b200.so is an error."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import lsf_b200
    lsf_b200._lib.load()
    return lsf_b200
