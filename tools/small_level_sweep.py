#!/usr/bin/env python
"""Which kernel generation is fastest on the small pyramid levels? CUDA-event time of the fixed-iteration driver
(lsf_hier_iterate_3d) at 32^3 .. 128^3 for the library's A/B switches. Usage: python tools/small_level_sweep.py"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lsf_b200
from lsf_b200 import synthetic

lib = lsf_b200._lib.load()
ptr = lambda t: ctypes.cast(ctypes.c_void_p(t.data_ptr()), lsf_b200._lib.c_float_p)
configs = sys.argv[1:] or ["", "LSF_DEFER=0", "LSF_TMA=0", "LSF_SPLIT_X=0", "LSF_LEGACY_KERNELS=1", "LSF_XCHUNK_T=4",
                           "LSF_XCHUNK_T=8", "LSF_XCHUNK_T=16", "LSF_XCHUNK_T=32", "LSF_YCHUNK_T=8", "LSF_YCHUNK_T=16",
                           "LSF_YCHUNK_T=32"]
kwargs = dict(tikhonov_term_enabled=True, gradient_kernel_enabled=True, tikhonov_strength=0.1,
              kernel=synthetic.sobolev_kernel_1d(), maximum_iteration_count=100)
params = lsf_b200.HierarchicalOptimizer3d(**kwargs)._params()
for size in (32, 64, 128):
    canonical, live = synthetic.sphere_plane_pair_3d(size, xp=torch, device="cuda")
    for config in configs:
        names = []
        for pair in filter(None, config.split(",")):
            name, value = pair.split("=")
            os.environ[name] = value
            names.append(name)
        ms, launches = ctypes.c_float(0), ctypes.c_int(0)
        best = None
        for _ in range(4):
            lsf_b200._lib.check(lib.lsf_hier_iterate_3d(ctypes.byref(params), ptr(canonical), ptr(live), size, size, size,
                                                        100, ctypes.byref(ms), ctypes.byref(launches), None,
                                                        lsf_b200._lib.current_stream_handle()))
            best = ms.value if best is None else min(best, ms.value)
        print("%4d^3 %-28s %.4f ms/iteration (%d launches)" % (size, config or "(defaults)", best / 100, launches.value // 100))
        for name in names:
            del os.environ[name]
