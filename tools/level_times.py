#!/usr/bin/env python
"""Where the time of one hierarchical optimize() goes: CUDA-event time of single-level runs (maximum_chunk_size 1) of
100 iterations at every pyramid level's size, of the whole 4-level run, and of the run without host polling gaps
(LSF_PIPELINE_POLL A/B). Usage: python tools/level_times.py [size]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lsf_b200
from lsf_b200 import synthetic

size = int(sys.argv[1]) if len(sys.argv) > 1 else 256
common = dict(tikhonov_term_enabled=True, gradient_kernel_enabled=True, tikhonov_strength=0.1,
              kernel=synthetic.sobolev_kernel_1d(), maximum_iteration_count=100, maximum_warp_update_threshold=0.0)


def timed(optimizer, canonical, live, repeats=3):
    best = None
    for _ in range(repeats):
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        start.record()
        optimizer.optimize(canonical, live)
        stop.record()
        torch.cuda.synchronize()
        best = start.elapsed_time(stop) if best is None else min(best, start.elapsed_time(stop))
    return best


total_levels = 0.0
for level_size in (size // 8, size // 4, size // 2, size):
    canonical, live = synthetic.sphere_plane_pair_3d(level_size, xp=torch, device="cuda")
    for poll in ("1", "0"):
        os.environ["LSF_PIPELINE_POLL"] = poll
        ms = timed(lsf_b200.HierarchicalOptimizer3d(maximum_chunk_size=1, **common), canonical, live)
        print("single level %4d^3, 100 iterations, pipelined poll %s: %8.3f ms (%.4f ms/iteration)"
              % (level_size, poll, ms, ms / 100))
    total_levels += ms
canonical, live = synthetic.sphere_plane_pair_3d(size, xp=torch, device="cuda")
for poll in ("1", "0"):
    os.environ["LSF_PIPELINE_POLL"] = poll
    ms = timed(lsf_b200.HierarchicalOptimizer3d(maximum_chunk_size=8, **common), canonical, live)
    print("4 levels %d^3, pipelined poll %s: %8.3f ms" % (size, poll, ms))
del os.environ["LSF_PIPELINE_POLL"]
one = timed(lsf_b200.HierarchicalOptimizer3d(maximum_chunk_size=8, **dict(common, maximum_iteration_count=1)), canonical, live)
print("4 levels, 1 iteration each (pyramid + prolongation + output): %8.3f ms" % one)
