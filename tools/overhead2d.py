#!/usr/bin/env python
"""Where does a launch-bound 2D optimize() spend its time? Times HierarchicalOptimizer2d / SobolevOptimizer2d on one pair
(numpy arrays and torch CUDA tensors, full run and a one-iteration run = the fixed cost) and, with LSF_TRACE=1, prints the
library's host-side trace of the last call of each variant. Usage: [LSF_TRACE=1] overhead2d.py [size]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import lsf_b200
from lsf_b200 import synthetic

size = int(sys.argv[1]) if len(sys.argv) > 1 else 128
canonical, live = synthetic.circle_line_pair_2d(size, shift=(5.0, -3.0), line_shift=-4.0)
kernel = synthetic.sobolev_kernel_1d()


def best_of(call, repeats=20):
    call()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(repeats):
        t0 = time.perf_counter()
        call()
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return 1e3 * best


for iterations in (100, 1):
    for name, kwargs in (("data_only", dict(tikhonov_term_enabled=False, gradient_kernel_enabled=False)),
                         ("tikhonov_kernel", dict(tikhonov_term_enabled=True, tikhonov_strength=0.2, gradient_kernel_enabled=True,
                                                  kernel=kernel))):
        kwargs = dict(kwargs, maximum_chunk_size=8, rate=0.2, maximum_iteration_count=iterations,
                      maximum_warp_update_threshold=0.0)
        optimizer = lsf_b200.HierarchicalOptimizer2d(**kwargs)
        host = best_of(lambda: optimizer.optimize(canonical, live))
        counts = optimizer.get_per_level_iteration_counts()
        c_dev, l_dev = torch.from_numpy(canonical).cuda(), torch.from_numpy(live).cuda()
        device = best_of(lambda: optimizer.optimize(c_dev, l_dev))
        print("hier2d %s %dx%d iterations %s: numpy %.3f ms, torch cuda tensors %.3f ms" % (name, size, size, counts, host, device),
              flush=True)
        sys.stderr.flush()

shared = lsf_b200.SharedParameters.get_instance()
lsf_b200.SobolevParameters.get_instance().set_sobolev_kernel(kernel)
for iterations, lower in ((100, 0.05), (1, 0.05)):
    shared.maximum_iteration_count = iterations
    shared.maximum_warp_length_lower_threshold = lower
    optimizer = lsf_b200.SobolevOptimizer2d()
    host = best_of(lambda: optimizer.optimize(live.copy(), canonical))
    print("sobolev2d %dx%d %d iterations: numpy %.3f ms" % (size, size, optimizer.get_iteration_count(), host), flush=True)
