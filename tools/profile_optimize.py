#!/usr/bin/env python
"""Runs whole hierarchical optimize() calls of the bench workload on cuda:0 and prints host-side phase times (target
of ncu launch lists; never a source of bench numbers). Usage: python tools/profile_optimize.py [size] [repeats]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lsf_b200
from lsf_b200 import synthetic
import bench

size = int(sys.argv[1]) if len(sys.argv) > 1 else 256
repeats = int(sys.argv[2]) if len(sys.argv) > 2 else 1
optimizer = lsf_b200.HierarchicalOptimizer3d(**bench.optimizer_kwargs())
canonical, live = synthetic.sphere_plane_pair_3d(size, xp=torch, device="cuda")
lib = lsf_b200._lib.load()
for i in range(repeats):
    torch.cuda.synchronize()
    before = lib.lsf_launch_count()
    t0 = time.perf_counter()
    optimizer.optimize(canonical, live)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    print("optimize %d: %.2f ms, %d launches, iterations %s" % (i, 1e3 * (t1 - t0), lib.lsf_launch_count() - before,
                                                                optimizer.get_per_level_iteration_counts()))
