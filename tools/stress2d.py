#!/usr/bin/env python
"""Repeats the single-launch 2D optimizers and compares every result with the first one bit for bit (a race between the blocks
of the cluster kernels would show as a run that differs). Usage: stress2d.py [repeats]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import lsf_b200
from lsf_b200 import synthetic

repeats = int(sys.argv[1]) if len(sys.argv) > 1 else 200
kernel = synthetic.sobolev_kernel_1d()
failures = 0
for size in (128, 96, 64):
    canonical, live = synthetic.circle_line_pair_2d(size, shift=(9.0, -7.0), line_shift=-6.0)
    chunk = 8 if size != 96 else 4
    cases = {
        "hier2d data term, early termination": dict(tikhonov_term_enabled=False, gradient_kernel_enabled=False, rate=0.3,
                                                    maximum_warp_update_threshold=0.02),
        "hier2d data term": dict(tikhonov_term_enabled=False, gradient_kernel_enabled=False, rate=0.3, maximum_warp_update_threshold=0.0),
        "hier2d tikhonov": dict(tikhonov_term_enabled=True, tikhonov_strength=0.05, gradient_kernel_enabled=False, rate=0.3,
                                maximum_warp_update_threshold=0.001),
        "hier2d tikhonov + kernel": dict(tikhonov_term_enabled=True, tikhonov_strength=0.1, gradient_kernel_enabled=True, kernel=kernel,
                                         rate=0.3, maximum_warp_update_threshold=0.001),
    }
    for name, kwargs in cases.items():
        optimizer = lsf_b200.HierarchicalOptimizer2d(maximum_chunk_size=chunk, maximum_iteration_count=60,
                                                     resampling_strategy=1 if size == 96 else 0, **kwargs)
        first, counts = optimizer.optimize(canonical, live), optimizer.get_per_level_iteration_counts()
        different = 0
        for _ in range(repeats):
            warp = optimizer.optimize(canonical, live)
            if not np.array_equal(warp, first) or optimizer.get_per_level_iteration_counts() != counts:
                different += 1
        failures += different
        print("%-38s %3d x %3d iterations %s: %d of %d runs differ" % (name, size, size, counts, different, repeats), flush=True)
    if size != 96:
        shared = lsf_b200.SharedParameters.get_instance()
        shared.maximum_iteration_count = 40
        shared.maximum_warp_length_lower_threshold = 0.05
        lsf_b200.SobolevParameters.get_instance().set_sobolev_kernel(kernel)
        optimizer = lsf_b200.SobolevOptimizer2d()
        first = optimizer.optimize(live.copy(), canonical)
        different = sum(0 if np.array_equal(optimizer.optimize(live.copy(), canonical), first) else 1 for _ in range(repeats))
        failures += different
        print("%-38s %3d x %3d: %d of %d runs differ" % ("SobolevFusion 2D", size, size, different, repeats), flush=True)
print("stress2d:", "ok" if failures == 0 else "%d runs differ" % failures)
sys.exit(1 if failures else 0)
