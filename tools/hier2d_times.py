#!/usr/bin/env python
"""2D hierarchical optimizer (reference HierarchicalOptimizer2d, the reference's own tests and 2D experiments) on one pair:
ms per optimize() from numpy arrays on the GPU path and on the CPU oracle, identical results. Usage: hier2d_times.py [size]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import lsf_b200
import oracle
from lsf_b200 import synthetic

size = int(sys.argv[1]) if len(sys.argv) > 1 else 128
canonical, live = synthetic.circle_line_pair_2d(size, shift=(5.0, -3.0), line_shift=-4.0)
for name, kwargs in (("data_only", dict(tikhonov_term_enabled=False, gradient_kernel_enabled=False)),
                     ("tikhonov_kernel", dict(tikhonov_term_enabled=True, tikhonov_strength=0.2, gradient_kernel_enabled=True,
                                              kernel=synthetic.sobolev_kernel_1d()))):
    kwargs = dict(kwargs, maximum_chunk_size=8, rate=0.2, maximum_iteration_count=100, maximum_warp_update_threshold=0.001)
    optimizer = lsf_b200.HierarchicalOptimizer2d(**kwargs)
    warp = optimizer.optimize(canonical, live)
    torch.cuda.synchronize()
    best = None
    for _ in range(5):
        t0 = time.perf_counter()
        warp = optimizer.optimize(canonical, live)
        torch.cuda.synchronize()
        best = time.perf_counter() - t0 if best is None else min(best, time.perf_counter() - t0)
    t0 = time.perf_counter()
    expected = oracle.hier_optimize(canonical, live, **kwargs)
    cpu = time.perf_counter() - t0
    print("2D hierarchical %s %dx%d: GPU %.3f ms per optimize (iterations %s, host arrays), CPU oracle %.3f ms; equal: %s"
          % (name, size, size, 1e3 * best, optimizer.get_per_level_iteration_counts(), 1e3 * cpu,
             bool(np.array_equal(warp, expected["warp"]))))
