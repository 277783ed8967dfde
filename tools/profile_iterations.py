#!/usr/bin/env python
"""Runs a few finest-level iterations of the hierarchical optimizer at size^3 on cuda:0 (target of ncu captures;
never a source of bench numbers). Usage: python tools/profile_iterations.py [size] [iterations] [mode]"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lsf_b200
from lsf_b200 import synthetic

size = int(sys.argv[1]) if len(sys.argv) > 1 else 256
iterations = int(sys.argv[2]) if len(sys.argv) > 2 else 3
mode = sys.argv[3] if len(sys.argv) > 3 else "tikhonov_kernel"
kwargs = dict(tikhonov_term_enabled="tikhonov" in mode, gradient_kernel_enabled="kernel" in mode,
              tikhonov_strength=0.1, kernel=synthetic.sobolev_kernel_1d(), maximum_iteration_count=100)
optimizer = lsf_b200.HierarchicalOptimizer3d(**kwargs)
canonical, live = synthetic.sphere_plane_pair_3d(size, xp=torch, device="cuda")
lib = lsf_b200._lib.load()
params = optimizer._params()
ms, launches = ctypes.c_float(0), ctypes.c_int(0)
ptr = lambda t: ctypes.cast(ctypes.c_void_p(t.data_ptr()), lsf_b200._lib.c_float_p)
lsf_b200._lib.check(lib.lsf_hier_iterate_3d(ctypes.byref(params), ptr(canonical), ptr(live), size, size, size, iterations,
                                            ctypes.byref(ms), ctypes.byref(launches), None,
                                            lsf_b200._lib.current_stream_handle()))
print("size %d mode %s: %d iterations, %d launches, %.3f ms/iteration" % (size, mode, iterations, launches.value,
                                                                       ms.value / iterations))
