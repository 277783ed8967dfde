#!/usr/bin/env python
"""3D KillingFusion / SobolevFusion (slavcheva) iteration time on cuda:0 (BASELINE.json configs[2]: Killing + level-set
terms + 7-tap Sobolev filter at 256^3). Usage: python tools/killing_times.py [size] [iterations]
Prints ms per iteration (CUDA events around whole optimize() calls of two different iteration counts, so that the
set-up and read-back cancel), voxel-updates/s and the fraction of the HBM roofline at SURVEY.md 8(d)'s 36 B/update."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lsf_b200
from lsf_b200 import synthetic
from lsf_b200.slavcheva import SmoothingTermMethod

size = int(sys.argv[1]) if len(sys.argv) > 1 else 256
iterations = int(sys.argv[2]) if len(sys.argv) > 2 else 20
canonical, live = synthetic.sphere_plane_pair_3d(size, xp=torch, device="cuda")
peak = 6555.5
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def run(n, **kwargs):
    optimizer = lsf_b200.SlavchevaOptimizer3d(max_iterations=n, min_iterations=n, maximum_warp_length_lower_threshold=0.0,
                                              sobolev_kernel=synthetic.sobolev_kernel_1d(), **kwargs)
    best = None
    for _ in range(3):
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        start.record()
        optimizer.optimize(live.clone(), canonical)
        stop.record()
        torch.cuda.synchronize()
        ms = start.elapsed_time(stop)
        best = ms if best is None else min(best, ms)
    assert optimizer.get_iteration_count() == n, optimizer.get_iteration_count()
    return best


for name, kwargs in (("sobolev_tikhonov", dict(smoothing_term_method=SmoothingTermMethod.TIKHONOV)),
                     ("killing_levelset", dict(smoothing_term_method=SmoothingTermMethod.KILLING,
                                               level_set_term_enabled=True))):
    short, long_ = run(iterations, **kwargs), run(3 * iterations, **kwargs)
    per_iteration = (long_ - short) / (2 * iterations)
    updates = size ** 3 / (per_iteration * 1e-3)
    achieved = 36 * updates / 1e9
    print("%s %d^3: %.4f ms/iteration, %.3e voxel-updates/s, %.1f GB/s algorithmic (36 B) = %.1f%% of %.1f GB/s"
          % (name, size, per_iteration, updates, achieved, 100 * achieved / peak, peak))
