"""CPU-side checks of the product library: it loads and exports every symbol include/lsf_b200.h declares
(no compute calls -- there is no GPU in the build container), and argument validation that needs no device."""
import os
import re

import lsf_b200
from lsf_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    header = open(os.path.join(ROOT, "include", "lsf_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    return sorted(set(re.findall(r"\b(lsf_[a-z0-9_]+)\s*\(", header)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = declared_symbols()
    assert len(names) >= 15
    for name in names:
        assert hasattr(lib, name), name
    assert sorted(_lib.EXPORTED_SYMBOLS) == names
    assert lib.lsf_version() >= 100


def test_alias_modules_expose_reference_names():
    import level_set_fusion_optimization as cpp
    for name in ("HierarchicalOptimizer2d", "HierarchicalOptimizer3d"):
        cls = getattr(cpp, name)
        assert cls.ResamplingStrategy.NEAREST_AND_AVERAGE == 0 and cls.ResamplingStrategy.LINEAR == 1
        cls.VerbosityParameters(print_max_warp_update=True)
        cls.LoggingParameters(collect_per_level_convergence_reports=True)
    optimizer = cpp.HierarchicalOptimizer3d()
    # reference defaults, cpp/src/nonrigid_optimization/hierarchical/optimizer.hpp:51-65
    assert optimizer.maximum_chunk_size == 8 and optimizer.rate == 0.1 and optimizer.maximum_iteration_count == 100
    assert optimizer.tikhonov_strength == 0.2 and optimizer.kernel is None


def test_struct_layouts_match_header():
    import ctypes
    assert ctypes.sizeof(_lib.HierParams) == 48
    assert ctypes.sizeof(_lib.IterationCapture) == 24
    assert ctypes.sizeof(_lib.LevelReport) == 92  # sizeof(lsf_level_report), checked with g++


def test_product_never_imports_oracle():
    """The oracle is test infrastructure; nothing under the product package may reference it."""
    package = os.path.join(ROOT, "levelsetfusion-python_b200")
    for base, _, files in os.walk(package):
        for name in files:
            if name.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(base, name)).read()
                assert "import oracle" not in text and "lsf_oracle" not in text, name
