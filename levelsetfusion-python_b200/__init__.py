"""lsf_b200 -- B200-native (sm_100a) implementation of LevelSetFusion's non-rigid warp-field optimisation.

The package directory is ``levelsetfusion-python_b200`` (not an importable name); import it through the
``lsf_b200`` alias package at the repository root, or as the reference's extension-module name
``level_set_fusion_optimization``.

Host-side mirror of the reference API for this path:
  * HierarchicalOptimizer2d / HierarchicalOptimizer3d      (hierarchical.py)
  * SobolevOptimizer2d + SharedParameters / SobolevParameters, SlavchevaOptimizer2d / 3d, warp_field_advanced (slavcheva.py)
  * telemetry types and builders (telemetry.py)
  * tsdf.FilteringMethod / Parameters2d / Parameters3d / Generator2d / Generator3d: TSDF generation from depth (tsdf.py)
  * Sdf2SdfOptimizer2d: rigid SDF-2-SDF tracker (rigid.py)
  * primitives warp / gradient / laplacian / convolution / resampling (ops.py)
backed by liblsf_b200.so (csrc/, C-ABI in include/lsf_b200.h). No CPU fallback exists.
"""
from . import _lib
from .hierarchical import (HierarchicalOptimizer2d, HierarchicalOptimizer3d, OptimizationIterationData2d,
                           OptimizationIterationData3d)
from . import ops
from . import telemetry
from .telemetry import (Vector2i, Vector3i, Vector2f, WarpDeltaStatistics2d, WarpDeltaStatistics3d,
                        TsdfDifferenceStatistics2d, TsdfDifferenceStatistics3d, ConvergenceReport2d, ConvergenceReport3d,
                        build_warp_delta_statistics_2d, build_warp_delta_statistics_3d,
                        build_tsdf_difference_statistics_2d, build_tsdf_difference_statistics_3d, mean_vector_length)
from ._lib import set_strict_inputs
from . import slavcheva
from . import tsdf
from .rigid import Sdf2SdfOptimizer2d
from .slavcheva import (SobolevOptimizer2d, SharedParameters, SobolevParameters, SlavchevaOptimizer2d,
                        SlavchevaOptimizer3d, ComputeMethod, AdaptiveLearningRateMethod, DataTermMethod,
                        SmoothingTermMethod, warp_field_advanced, warp_field_advanced_no_warp_change,
                        data_term_at_location)

# reference python_export/telemetry.tpp:113-145 exports std::vector wrappers of the report / iteration-data types; plain
# Python lists play that role here
ConvergenceReportVector2d = ConvergenceReportVector3d = list
OptimizationIterationDataVector2d = OptimizationIterationDataVector3d = list

__all__ = ["HierarchicalOptimizer2d", "HierarchicalOptimizer3d", "OptimizationIterationData2d",
           "OptimizationIterationData3d", "ConvergenceReportVector2d", "ConvergenceReportVector3d",
           "OptimizationIterationDataVector2d", "OptimizationIterationDataVector3d", "SobolevOptimizer2d", "SharedParameters",
           "SobolevParameters", "SlavchevaOptimizer2d", "SlavchevaOptimizer3d", "ComputeMethod",
           "AdaptiveLearningRateMethod", "DataTermMethod", "SmoothingTermMethod", "warp_field_advanced",
           "warp_field_advanced_no_warp_change", "data_term_at_location", "Vector2i", "Vector3i", "Vector2f",
           "WarpDeltaStatistics2d", "WarpDeltaStatistics3d", "TsdfDifferenceStatistics2d", "TsdfDifferenceStatistics3d",
           "ConvergenceReport2d", "ConvergenceReport3d", "build_warp_delta_statistics_2d",
           "build_warp_delta_statistics_3d", "build_tsdf_difference_statistics_2d",
           "build_tsdf_difference_statistics_3d", "mean_vector_length", "ops", "telemetry", "slavcheva", "tsdf", "Sdf2SdfOptimizer2d", "set_strict_inputs", "_lib"]
