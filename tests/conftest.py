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


class TsdfCases:
    """tests/golden/reference_tsdf.npz (made by tests/golden/make_tsdf_golden.py): the reference's own TSDF-generation
    test cases (`cases`) and runs of the reference's Python generators (`python_runs`); every entry is
    (parameters dict, depth image uint16, expected field)."""

    def __init__(self):
        import json
        import numpy as np
        data = np.load(os.path.join(ROOT, "tests", "golden", "reference_tsdf.npz"))
        self.images = {key[len("image/"):]: data[key] for key in data.files if key.startswith("image/")}

        def entries(prefix):
            out = []
            while "%s/%02d/expected" % (prefix, len(out)) in data.files:
                k = len(out)
                parameters = json.loads(str(data["%s/%02d/parameters" % (prefix, k)]))
                image = self.images[parameters["image"]].copy()
                if parameters["zero_depth_to_maximum"]:  # tests/test_tsdf_ewa.py:30-37 image_load_helper
                    image[image == 0] = np.iinfo(np.uint16).max
                out.append((parameters, image, data["%s/%02d/expected" % (prefix, k)]))
            return out

        self.cases = entries("case")
        self.python_runs = entries("python")


@pytest.fixture(scope="session")
def tsdf_cases():
    return TsdfCases()
