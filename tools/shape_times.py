#!/usr/bin/env python
"""Stage times of the finest-level iteration for volumes of equal voxel count but different row length Z
(diagnostic: how much does the length of the contiguous pieces a tile fetches matter?).
Usage: python tools/shape_times.py [iterations] [mode]"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lsf_b200
from lsf_b200 import synthetic

iterations = int(sys.argv[1]) if len(sys.argv) > 1 else 20
mode = sys.argv[2] if len(sys.argv) > 2 else "tikhonov_kernel"
lib = lsf_b200._lib.load()
canonical, live = synthetic.sphere_plane_pair_3d(256, xp=torch, device="cuda")
ptr = lambda t: ctypes.cast(ctypes.c_void_p(t.data_ptr()), lsf_b200._lib.c_float_p)
kwargs = dict(tikhonov_term_enabled="tikhonov" in mode, gradient_kernel_enabled="kernel" in mode,
              tikhonov_strength=0.1, kernel=synthetic.sobolev_kernel_1d(), maximum_iteration_count=100)
params = lsf_b200.HierarchicalOptimizer3d(**kwargs)._params()
for Z in (256, 128, 64):
    parts = 256 // Z
    c = torch.cat([canonical[:, :, i * Z:(i + 1) * Z] for i in range(parts)], dim=1).contiguous()
    l = torch.cat([live[:, :, i * Z:(i + 1) * Z] for i in range(parts)], dim=1).contiguous()
    X, Y = c.shape[0], c.shape[1]
    ms, launches = ctypes.c_float(0), ctypes.c_int(0)
    stage = (ctypes.c_float * 4)()
    for use_stage in (None, None, stage):
        lsf_b200._lib.check(lib.lsf_hier_iterate_3d(ctypes.byref(params), ptr(c), ptr(l), X, Y, Z, iterations,
                                                    ctypes.byref(ms), ctypes.byref(launches), use_stage,
                                                    lsf_b200._lib.current_stream_handle()))
        if use_stage is None:
            total = ms.value / iterations
    print("%s %4d x %4d x %4d: %.4f ms/iteration; stages %s" % (mode, X, Y, Z, total,
                                                          " ".join("%.4f" % (stage[i] / iterations) for i in range(4))))
