#!/usr/bin/env python
"""Per-stage CUDA-event times of the finest-level hierarchical iteration for a list of tuning settings
(environment variables read by the library at plan time). Usage:
    python tools/stage_times.py [size] [iterations] "LSF_LANE_XV=2" "LSF_STAGE1_VARIANT=1" ...
Each argument after the first two is one configuration: comma-separated NAME=VALUE pairs ("" = defaults)."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lsf_b200
from lsf_b200 import synthetic

size = int(sys.argv[1]) if len(sys.argv) > 1 else 256
iterations = int(sys.argv[2]) if len(sys.argv) > 2 else 20
configs = sys.argv[3:] or [""]
lib = lsf_b200._lib.load()
canonical, live = synthetic.sphere_plane_pair_3d(size, xp=torch, device="cuda")
ptr = lambda t: ctypes.cast(ctypes.c_void_p(t.data_ptr()), lsf_b200._lib.c_float_p)
for mode in ("tikhonov_kernel", "kernel", "tikhonov", "data_only"):
    kwargs = dict(tikhonov_term_enabled="tikhonov" in mode, gradient_kernel_enabled="kernel" in mode,
                  tikhonov_strength=0.1, kernel=synthetic.sobolev_kernel_1d(), maximum_iteration_count=100)
    params = lsf_b200.HierarchicalOptimizer3d(**kwargs)._params()
    for config in configs:
        names = []
        for pair in filter(None, config.split(",")):
            name, value = pair.split("=")
            os.environ[name] = value
            names.append(name)
        ms, launches = ctypes.c_float(0), ctypes.c_int(0)
        stage = (ctypes.c_float * 4)()
        for use_stage in (None, None, stage):
            lsf_b200._lib.check(lib.lsf_hier_iterate_3d(ctypes.byref(params), ptr(canonical), ptr(live), size, size, size,
                                                        iterations, ctypes.byref(ms), ctypes.byref(launches), use_stage,
                                                        lsf_b200._lib.current_stream_handle()))
            if use_stage is None:
                total = ms.value / iterations
        print("%-16s %-40s %.4f ms/iteration; stages %s" % (mode, config or "(defaults)", total,
                                                          " ".join("%.4f" % (stage[i] / iterations) for i in range(4))))
        for name in names:
            del os.environ[name]
