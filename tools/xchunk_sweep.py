#!/usr/bin/env python
"""Sweep of unequal x-chunk schedules of stage 1 (LSF_XCHUNKS) at size^3: prints the finest-level iteration time for
every candidate, best last. Usage: python tools/xchunk_sweep.py [size] [iterations]"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lsf_b200
from lsf_b200 import synthetic

size = int(sys.argv[1]) if len(sys.argv) > 1 else 256
iterations = int(sys.argv[2]) if len(sys.argv) > 2 else 20
lib = lsf_b200._lib.load()
canonical, live = synthetic.sphere_plane_pair_3d(size, xp=torch, device="cuda")
ptr = lambda t: ctypes.cast(ctypes.c_void_p(t.data_ptr()), lsf_b200._lib.c_float_p)
params = lsf_b200.HierarchicalOptimizer3d(tikhonov_term_enabled=True, gradient_kernel_enabled=True, tikhonov_strength=0.1,
                                          kernel=synthetic.sobolev_kernel_1d(), maximum_iteration_count=100)._params()


def measure(schedule):
    if schedule:
        os.environ["LSF_XCHUNKS"] = ":".join(str(c) for c in schedule)
    else:
        os.environ.pop("LSF_XCHUNKS", None)
    ms, launches = ctypes.c_float(0), ctypes.c_int(0)
    best = None
    for _ in range(2):
        lsf_b200._lib.check(lib.lsf_hier_iterate_3d(ctypes.byref(params), ptr(canonical), ptr(live), size, size, size,
                                                    iterations, ctypes.byref(ms), ctypes.byref(launches), None,
                                                    lsf_b200._lib.current_stream_handle()))
        best = ms.value if best is None else min(best, ms.value)
    return best / iterations


candidates = [None]
unit = max(size // 128, 1)
if len(sys.argv) > 3 and sys.argv[3] == "fine":
    # around the schedule marching_schedule() produces for 256 tiles on 444 slots
    for first in range(142 * size // 256, 160 * size // 256, 4 * size // 256):
        for second in range(60 * size // 256, 74 * size // 256, 3 * size // 256):
            rest = size - first - second
            for pieces in (2, 3):
                tail = [rest // pieces + (1 if i < rest % pieces else 0) for i in range(pieces)]
                candidates.append([first, second] + tail)
else:
    for first in range(50 * unit, 110 * unit, 5 * unit):
        for second in range(12 * unit, 50 * unit, 6 * unit):
            rest = size - first - second
            if rest < 0:
                continue
            for pieces in (1, 2, 3, 4):
                if rest == 0 and pieces > 1:
                    continue
                if rest and rest // pieces < 6 * unit:
                    continue
                tail = [rest // pieces + (1 if i < rest % pieces else 0) for i in range(pieces)] if rest else []
                candidates.append([first, second] + tail)
results = []
for schedule in candidates:
    results.append((measure(schedule), schedule))
    measure(None)
results.sort(key=lambda r: -r[0])
for ms, schedule in results[-25:]:
    print("%.4f ms  %s" % (ms, schedule if schedule else "uniform (default)"))
print("uniform: %.4f" % [ms for ms, s in results if s is None][0])
