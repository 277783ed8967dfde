#!/usr/bin/env python
"""BASELINE.json configs[0]: 2D SobolevFusion (slavcheva optimizer, Sobolev smoothing) on one 128 x 128 pair with the
reference experiment's parameters (experiment/singleframe_experiment.py:91-116). Times the GPU path (numpy in, numpy
out: host staging included) and the CPU oracle (restatement of the reference's C++ SobolevOptimizer2d) on the same
pair and checks that they agree. Usage: python tools/sobolev2d_times.py [size]"""
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
kernel = synthetic.sobolev_kernel_1d()
lsf_b200.SharedParameters.get_instance().maximum_iteration_count = 100
lsf_b200.SharedParameters.get_instance().maximum_warp_length_lower_threshold = 0.05
lsf_b200.SobolevParameters.get_instance().set_sobolev_kernel(kernel)
optimizer = lsf_b200.SobolevOptimizer2d()
result = optimizer.optimize(live.copy(), canonical)
torch.cuda.synchronize()
best = None
for _ in range(5):
    t0 = time.perf_counter()
    result = optimizer.optimize(live.copy(), canonical)
    torch.cuda.synchronize()
    best = time.perf_counter() - t0 if best is None else min(best, time.perf_counter() - t0)
iterations = optimizer.get_iteration_count()
t0 = time.perf_counter()
expected = oracle.slavcheva_optimize(live, canonical, semantics=0, max_iterations=100,
                                     maximum_warp_length_lower_threshold=0.05, sobolev_kernel=kernel)
cpu = time.perf_counter() - t0
print("2D SobolevFusion %dx%d: GPU %.3f ms per optimize (%d iterations, host arrays), CPU oracle %.3f ms (%d iterations); "
      "live fields equal: %s" % (size, size, 1e3 * best, iterations, 1e3 * cpu, expected["iterations"],
                                bool(np.array_equal(result, expected["live"]))))
