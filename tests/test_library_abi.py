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
    # reference python_export/slavcheva_optimizer.cpp:64-137, telemetry.tpp:50-145, math.cpp:36-70
    for name in ("SobolevOptimizer2d", "SharedParameters", "SobolevParameters", "warp_field_advanced",
                 "warp_field_advanced_no_warp_change", "data_term_at_location", "WarpDeltaStatistics2d",
                 "WarpDeltaStatistics3d", "TsdfDifferenceStatistics2d", "TsdfDifferenceStatistics3d",
                 "ConvergenceReport2d", "ConvergenceReport3d", "build_warp_delta_statistics_2d",
                 "build_warp_delta_statistics_3d", "build_tsdf_difference_statistics_2d",
                 "build_tsdf_difference_statistics_3d", "Vector2i", "Vector3i", "Vector2f", "mean_vector_length"):
        assert hasattr(cpp, name), name
    shared = cpp.SharedParameters.get_instance()
    assert shared is cpp.SharedParameters.get_instance()
    # reference defaults optimizer2d.hpp:59-78, sobolev_optimizer2d.hpp:51-70
    assert shared.gradient_descent_rate == 0.1 and shared.maximum_warp_length_lower_threshold == 0.1
    assert shared.maximum_warp_length_upper_threshold == 10000 and shared.maximum_iteration_count == 100
    assert shared.minimum_iteration_count == 1
    sobolev = cpp.SobolevParameters.get_instance()
    assert sobolev.smoothing_term_weight == 0.2 and len(sobolev.get_sobolev_kernel()) == 7
    report = cpp.ConvergenceReport2d(2, True, cpp.WarpDeltaStatistics2d(0.25, 0.0, 1.0, 0.5, 0.1, cpp.Vector2i(1, 2),
                                                                         False, False),
                                     cpp.TsdfDifferenceStatistics2d(0, 0.2, 0.1, 0.05, cpp.Vector2i(3, 3)))
    same = cpp.ConvergenceReport2d(2, True, cpp.WarpDeltaStatistics2d(0.25 + 1e-6, 0.0, 1.0, 0.5, 0.1, cpp.Vector2i(1, 2),
                                                                       False, False),
                                   cpp.TsdfDifferenceStatistics2d(0, 0.2, 0.1, 0.05, cpp.Vector2i(3, 3)))
    assert report == same and "iter count: 2" in str(report)
    same.warp_delta_statistics.longest_warp_location = cpp.Vector2i(2, 1)
    assert report != same
    optimizer = cpp.HierarchicalOptimizer3d()
    # reference defaults, cpp/src/nonrigid_optimization/hierarchical/optimizer.hpp:51-65
    assert optimizer.maximum_chunk_size == 8 and optimizer.rate == 0.1 and optimizer.maximum_iteration_count == 100
    assert optimizer.tikhonov_strength == 0.2 and optimizer.kernel is None


def test_struct_layouts_match_header():
    import ctypes
    assert ctypes.sizeof(_lib.HierParams) == 48
    assert ctypes.sizeof(_lib.IterationCapture) == 24
    assert ctypes.sizeof(_lib.LevelReport) == 92  # sizeof(lsf_level_report), checked with g++
    assert ctypes.sizeof(_lib.SlavchevaParams) == 72
    assert ctypes.sizeof(_lib.WarpDeltaStatisticsRaw) == 40
    assert ctypes.sizeof(_lib.TsdfDifferenceStatisticsRaw) == 28
    assert ctypes.sizeof(_lib.SlavchevaReport) == 84


def test_product_never_imports_oracle():
    """The oracle is test infrastructure; nothing under the product package may reference it."""
    package = os.path.join(ROOT, "levelsetfusion-python_b200")
    for base, _, files in os.walk(package):
        for name in files:
            if name.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(base, name)).read()
                assert "import oracle" not in text and "lsf_oracle" not in text, name


def test_strict_inputs_refuse_what_the_reference_refuses():
    """reference python_export/eigen_numpy_tensor.cpp:120-156: float32, C-contiguous, aligned arrays only (TypeError through
    Boost.Python's overload resolution); lenient conversion is this package's default"""
    import numpy as np
    import pytest
    field = np.zeros((4, 6), dtype=np.float64)
    assert _lib.as_f32(field).dtype == np.float32
    previous = lsf_b200.set_strict_inputs(True)
    try:
        with pytest.raises(TypeError):
            _lib.as_f32(field)
        with pytest.raises(TypeError):
            _lib.as_f32(np.zeros((4, 6), dtype=np.float32).T)
        with pytest.raises(TypeError):
            _lib.as_f32([[0.0, 1.0]])
        good = np.zeros((4, 6), dtype=np.float32)
        assert _lib.as_f32(good) is good
    finally:
        lsf_b200.set_strict_inputs(previous)
    assert _lib.as_f32(field).dtype == np.float32
