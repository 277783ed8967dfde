"""lsf_b200 -- B200-native (sm_100a) implementation of LevelSetFusion's non-rigid warp-field optimisation.

The package directory is ``levelsetfusion-python_b200`` (not an importable name); import it through the
``lsf_b200`` alias package at the repository root, or as the reference's extension-module name
``level_set_fusion_optimization``.

Host-side mirror of the reference API for this path:
  * HierarchicalOptimizer2d / HierarchicalOptimizer3d      (hierarchical.py)
  * primitives warp / gradient / laplacian / convolution / resampling (ops.py)
backed by liblsf_b200.so (csrc/, C-ABI in include/lsf_b200.h). No CPU fallback exists.
"""
from . import _lib
from .hierarchical import HierarchicalOptimizer2d, HierarchicalOptimizer3d, ConvergenceReport
from . import ops

__all__ = ["HierarchicalOptimizer2d", "HierarchicalOptimizer3d", "ConvergenceReport", "ops", "_lib"]
