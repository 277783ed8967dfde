#!/usr/bin/env python
"""End-to-end throughput of the 256^3 workload from pageable host arrays with 1 / 2 / 3 pairs in flight per GPU
(multigpu.optimize_pairs(streams=k): one worker thread and CUDA stream per pair, so the copies of one pair and the
launch-bound coarse levels overlap the other pair's finest level). Usage: e2e_pipeline_probe.py [pairs]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
import lsf_b200
from lsf_b200 import multigpu, synthetic

torch.cuda.set_device(0)
pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 8
canonical, live = synthetic.sphere_plane_pair_3d(256)
factory = lambda: lsf_b200.HierarchicalOptimizer3d(**bench.optimizer_kwargs())
serial = factory()
reference = serial.optimize(canonical, live)
for streams in (1, 2, 3):
    # the result array is looked at and dropped (its page-locked block goes back to the library's cache)
    worker = multigpu.PerWorkerOptimizer(factory, lambda optimizer, c, l: float(optimizer.optimize(c, l)[100, 100, 100, 0]) == float(reference[100, 100, 100, 0]))
    load = lambda index: (canonical, live)
    multigpu.optimize_pairs(worker, streams, load, 0, 1, gather=False, streams=streams)  # warm-up of every worker's staging
    torch.cuda.synchronize()
    best = None
    for _ in range(2):
        t0 = time.perf_counter()
        results = multigpu.optimize_pairs(worker, pairs, load, 0, 1, gather=False, streams=streams)
        torch.cuda.synchronize()
        seconds = time.perf_counter() - t0
        best = seconds if best is None else min(best, seconds)
    same = all(results.values())
    print("streams %d: %.2f ms per pair (%d pairs), results identical: %s" % (streams, 1e3 * best / pairs, pairs, same))
