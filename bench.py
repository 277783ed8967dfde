#!/usr/bin/env python
"""bench.py -- voxel-updates/s of the 3D warp-field optimisation at 256^3 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--size 256]

A "step" is one full hierarchical optimize() of one synthetic 256^3 TSDF pair (4-level pyramid, data term +
Tikhonov term + 7-tap Sobolev kernel, up to 100 iterations per level). voxel-updates = sum over levels of
voxels(level) * iterations(level). With N > 1 (one process per GPU, launched by torch.distributed.run) every
rank optimises its own pair (independent frame pairs, no data-path collective): weak scaling; the time is the
max over ranks and `value` the total over ranks.

Prints ONE JSON line (see the task contract): value (inputs resident in HBM), e2e (host buffers through the
public API, H2D/D2H inside the timed region), roofline (fused-iteration kernels at the finest level, measured
live with CUDA events), cpu_baseline (the CPU oracle on the host cores, bounded sample), clocks, gpu_launches.

--impl reference times the CPU restatement of the reference (oracle/; the reference's own C++ cannot be built
here: it needs Eigen + Boost.Python) on all host cores on a bounded sample of the same workload.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np

METRIC = "voxel-updates/s of 3D warp-field optimization at 256^3"
UNIT = "voxel-updates/s"
ALGORITHMIC_BYTES_PER_VOXEL_UPDATE = 68  # hierarchical 3D with Tikhonov (+- kernel), see DESIGN.md
# dram__bytes_read.sum + dram__bytes_write.sum of the kernels of one finest-level 256^3 iteration, from the committed
# `ncu --set full` capture (per launch group, like `achieved`)
NCU_DRAM_BYTES_PER_ITERATION = 1559500000
NCU_TRAFFIC_SOURCE = "profiles/r2_ncu_headline.md (stage 1 1186.0 MB + filter 373.5 MB)"
STAGE_NAMES = {1: ("fused_iteration",), 2: ("stage1_gather_terms_axis0", "ymarch_axis12_update_max"),
               4: ("gradient_stage", "filter_axis0", "filter_axis1", "filter_axis2_update_max")}


def optimizer_kwargs():
    from lsf_b200 import synthetic
    # reference run script values (run_hierarchical_optimizer3d.py:63-98: rate 0.1, threshold 0.01, 100
    # iterations, chunk 8) with the Tikhonov term and the Sobolev kernel switched on. tikhonov_strength 0.1:
    # the reference's Laplacian-of-previous-gradient feedback amplifies the 3D checkerboard mode by
    # 12 * strength * |K(pi)|^3 = 0.80 per iteration at 0.1 (stable) but 1.59 at the default 0.2 (diverges).
    return dict(tikhonov_term_enabled=True, gradient_kernel_enabled=True, maximum_chunk_size=8, rate=0.1,
                maximum_iteration_count=100, maximum_warp_update_threshold=0.01, data_term_amplifier=1.0,
                tikhonov_strength=0.1, kernel=synthetic.sobolev_kernel_1d(7, 0.1), resampling_strategy=0)


def workload_name(size):
    return "hierarchical3d_%d_tikhonov_sobolev7_4level" % size


def workload_config(size, kwargs, world):
    """The `config` object of the JSON line -- the same for both arms (the driver compares them)."""
    return {"workload": workload_name(size), "volume": [size] * 3, "pairs_per_gpu": 1, "levels": 4,
            "maximum_iterations_per_level": kwargs["maximum_iteration_count"],
            "maximum_warp_update_threshold": kwargs["maximum_warp_update_threshold"],
            "tikhonov_strength": kwargs["tikhonov_strength"],
            "tikhonov_strength_note": "0.1, not the reference default 0.2: with the Sobolev kernel the reference's "
                                      "Laplacian-of-previous-gradient feedback diverges in 3D at 0.2 "
                                      "(tests/test_oracle_golden.py::test_tikhonov_strength_0p2_diverges_in_3d)",
            "kernel_taps": 7,
            "l2_policy": "working set ~1.1 GB per iteration >> 126 MB L2 (inputs larger than L2)",
            "parallelism": "independent pairs, 1 per GPU" if world > 1 else "single GPU"}


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.lines = []
        self.process = None
        self.thread = None

    def start(self):
        try:
            self.process = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu_index), "--query-gpu=" + self.QUERY, "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.process = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.process.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.process is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.process.terminate()
        try:
            self.process.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.process.kill()
        sm, sm_max, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                sm_max.append(float(parts[2]))
            except ValueError:
                continue
            for name, value in zip(names, parts[5:9]):
                if value.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(sm_max) if sm_max else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def voxel_updates(reports):
    total = 0
    for r in reports:
        total += r.dims[0] * r.dims[1] * r.dims[2] * r.iteration_count
    return total


def run_ours(args):
    import torch
    import torch.distributed as dist
    import lsf_b200
    from lsf_b200 import synthetic

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device visible; the CUDA path has no fallback")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group(backend="nccl", device_id=device)
        # one process per GPU: keep this rank's host buffers and staging threads on the GPU's NUMA node
        from lsf_b200 import multigpu
        full_affinity = os.sched_getaffinity(0)
        numa_cpus = multigpu.bind_to_gpu_numa_node(local_rank)
    else:
        full_affinity = numa_cpus = None
    lib = lsf_b200._lib.load()
    size = args.size
    kwargs = optimizer_kwargs()
    optimizer = lsf_b200.HierarchicalOptimizer3d(**kwargs)

    # every rank gets its own pair (C4-style variation of the C2 geometry), generated on the device
    rng = np.random.default_rng(1234 + rank)
    shift = (2.5 + rng.uniform(-0.5, 0.5), -1.5 + rng.uniform(-0.5, 0.5), 1.0 + rng.uniform(-0.5, 0.5)) \
        if world > 1 else (2.5, -1.5, 1.0)
    canonical, live = synthetic.sphere_plane_pair_3d(size, shift=shift, xp=torch, device=device)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ------------------------------------------------------------------ device-resident throughput (`value`)
    for _ in range(args.warmup):
        optimizer.optimize(canonical, live)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches_before = lib.lsf_launch_count()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    updates = 0
    start.record()
    for _ in range(args.steps):
        optimizer.optimize(canonical, live)
        updates += voxel_updates(optimizer.get_per_level_convergence_reports())
    stop.record()
    barrier()
    elapsed_ms = start.elapsed_time(stop)
    launches = lib.lsf_launch_count() - launches_before
    iteration_counts = optimizer.get_per_level_iteration_counts()

    # ------------------------------------------------------------------ end to end through the public API (`e2e`)
    # (a) the reference-shaped call: pageable numpy arrays in, a fresh numpy array out
    host_c, host_l = canonical.cpu().numpy(), live.cpu().numpy()
    # warm-up of the staging path; two results alive at once, as in the timed loop (the result arrays come from a cache of
    # page-locked blocks, see _lib.result_array)
    warm = [optimizer.optimize(host_c, host_l) for _ in range(2)]
    del warm
    barrier()
    e2e_updates = 0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        warp_host = optimizer.optimize(host_c, host_l)  # H2D of both fields + D2H of the warp field inside
        e2e_updates += voxel_updates(optimizer.get_per_level_convergence_reports())
    torch.cuda.synchronize()
    e2e_seconds = time.perf_counter() - t0
    assert isinstance(warp_host, np.ndarray) and warp_host.shape == tuple(canonical.shape) + (3,)
    # (b) caller-pinned buffers and the out= extension
    pinned_c = torch.empty(canonical.shape, dtype=torch.float32, pin_memory=True)
    pinned_l = torch.empty(live.shape, dtype=torch.float32, pin_memory=True)
    pinned_out = torch.empty(tuple(canonical.shape) + (3,), dtype=torch.float32, pin_memory=True)
    pinned_c.copy_(canonical)
    pinned_l.copy_(live)
    optimizer.optimize(pinned_c.numpy(), pinned_l.numpy(), out=pinned_out.numpy())
    barrier()
    pinned_updates = 0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        optimizer.optimize(pinned_c.numpy(), pinned_l.numpy(), out=pinned_out.numpy())
        pinned_updates += voxel_updates(optimizer.get_per_level_convergence_reports())
    torch.cuda.synchronize()
    pinned_seconds = time.perf_counter() - t0
    del pinned_c, pinned_l, pinned_out
    clocks = sampler.stop() if rank == 0 else None

    if world > 1:
        stats = torch.tensor([elapsed_ms, e2e_seconds, pinned_seconds], dtype=torch.float64, device=device)
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
        elapsed_ms, e2e_seconds, pinned_seconds = float(stats[0]), float(stats[1]), float(stats[2])
        sums = torch.tensor([updates, e2e_updates, launches, pinned_updates], dtype=torch.float64, device=device)
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
        updates, e2e_updates, launches, pinned_updates = int(sums[0]), int(sums[1]), int(sums[2]), int(sums[3])

    result = None
    roofline = other_workloads = None
    lib.lsf_trim()  # the roofline kernels run in freshly allocated scratch, like a first optimize() call
    if rank == 0:
        # -------------------------------------------------------------- roofline of the finest-level iteration
        import ctypes
        N = size ** 3
        iterations = kwargs["maximum_iteration_count"]  # what a level of the workload executes (100)
        params = optimizer._params()
        ms = ctypes.c_float(0.0)
        n_launch = ctypes.c_int(0)
        stage_ms = (ctypes.c_float * 4)()
        ptr = lambda t: ctypes.cast(ctypes.c_void_p(t.data_ptr()), lsf_b200._lib.c_float_p)
        stream = lsf_b200._lib.current_stream_handle()
        for _ in range(2):  # first call warms up, second is kept
            lsf_b200._lib.check(lib.lsf_hier_iterate_3d(ctypes.byref(params), ptr(canonical), ptr(live), size, size, size,
                                                        iterations, ctypes.byref(ms), ctypes.byref(n_launch), None,
                                                        stream))
        iteration_ms = ms.value / iterations
        lsf_b200._lib.check(lib.lsf_hier_iterate_3d(ctypes.byref(params), ptr(canonical), ptr(live), size, size, size,
                                                    iterations, ctypes.byref(ms), ctypes.byref(n_launch), stage_ms,
                                                    stream))
        stages = [stage_ms[i] / iterations for i in range(4)]
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_source = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_source = 6650.0, "fallback (B200_PROFILING.md)"
        achieved = ALGORITHMIC_BYTES_PER_VOXEL_UPDATE * N / (iteration_ms * 1e-3) / 1e9
        # the same measurement for the optimizer's other three term configurations (44 B per
        # voxel-update without the Tikhonov term, 68 B with it); data-only is the reference run script's default
        other = {}
        for name, bytes_per_update, tikhonov, use_kernel in (("data_only", 44, False, False), ("tikhonov", 68, True, False),
                                                             ("sobolev_kernel", 44, False, True)):
            variant = dict(kwargs, tikhonov_term_enabled=tikhonov, gradient_kernel_enabled=use_kernel)
            variant_params = lsf_b200.HierarchicalOptimizer3d(**variant)._params()
            for _ in range(2):
                lsf_b200._lib.check(lib.lsf_hier_iterate_3d(ctypes.byref(variant_params), ptr(canonical), ptr(live), size,
                                                            size, size, iterations, ctypes.byref(ms),
                                                            ctypes.byref(n_launch), None, stream))
            variant_ms = ms.value / iterations
            variant_achieved = bytes_per_update * N / (variant_ms * 1e-3) / 1e9
            other[name] = {"ms_per_iteration": round(variant_ms, 4), "algorithmic_bytes_per_voxel_update": bytes_per_update,
                           "achieved": round(variant_achieved, 1), "frac": round(variant_achieved / peak, 4),
                           "launches_per_iteration": n_launch.value // iterations}
        launches_per_iteration = len([v for v in stages if v > 0])
        roofline = {
            "bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
            "frac": round(achieved / peak, 4), "traffic": NCU_DRAM_BYTES_PER_ITERATION if size == 256 else None,
            "traffic_source": NCU_TRAFFIC_SOURCE if size == 256 else None,
            "kernel": "finest-level iteration = k_hier_stage1_tma<APPLY> (TMA-fed: previous warp update + gather + data + "
                      "Tikhonov + axis-0 filter pass) + k_sobolev_ymarch3 (axis-1/2 filter passes + max norm), "
                      "%d launches/iteration, programmatic dependent launch"
                      % launches_per_iteration,
            "algorithmic_bytes_per_launch_group": ALGORITHMIC_BYTES_PER_VOXEL_UPDATE * N,
            "ms_per_iteration": round(iteration_ms, 4),
            "stage_ms": dict(zip(STAGE_NAMES[launches_per_iteration], (round(v, 4) for v in stages))),
            "peak_source": peak_source,
            "other_term_configurations": other,
        }
        # -------------------------------------------------------------- BASELINE.json configs[2]: 3D KillingFusion at 256^3
        other_workloads = {"killingfusion3d_%d" % size: killingfusion_iteration(lsf_b200, canonical, live, size, peak)}
        # -------------------------------------------------------------- TSDF generation from depth
        other_workloads["tsdf_generation_%d" % size] = tsdf_generation(lsf_b200, size, peak)
        # -------------------------------------------------------------- BASELINE.json configs[0]: the 2D optimizers at 128 x 128
        other_workloads.update(optimizers_2d(lsf_b200))
    # ------------------------------------------------------------------ BASELINE.json configs[3] and configs[4] (all ranks)
    del canonical, live
    torch.cuda.empty_cache()
    lib.lsf_trim()
    shared_workloads = {}
    if size == 256:
        shared_workloads["multipair_64x128"] = multipair_batch(rank, world, device)
        if world > 1:
            lib.lsf_trim()
            shared_workloads["slab_1024"] = slab_volume(1024, 20, rank, world, device)
    if rank == 0:
        other_workloads.update(shared_workloads)
        # -------------------------------------------------------------- CPU baseline (oracle port), bounded sample
        if numa_cpus:
            os.sched_setaffinity(0, full_affinity)  # the CPU baseline may use every host core
        cpu = cpu_baseline_sample(size, kwargs)
        value = updates / (elapsed_ms * 1e-3)
        N = size ** 3
        result = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(size, kwargs, world),
            "iterations_per_level": iteration_counts,
            # e2e = the reference-shaped call: pageable numpy arrays in, a fresh numpy array out (the library stages them
            # through its pinned ring); `pinned` = the same with caller-pinned buffers and the out= extension
            "e2e": {"value": e2e_updates / e2e_seconds, "unit": UNIT,
                    "h2d_bytes_per_step": 2 * 4 * N * world, "d2h_bytes_per_step": 3 * 4 * N * world,
                    "ms_per_step": 1e3 * e2e_seconds / args.steps,
                    "host_cpus_bound_to_gpu_numa_node": len(numa_cpus) if numa_cpus else 0,
                    "call": "optimizer.optimize(canonical: np.ndarray, live: np.ndarray) -> np.ndarray (pageable host memory)",
                    "pinned": {"value": pinned_updates / pinned_seconds, "ms_per_step": 1e3 * pinned_seconds / args.steps,
                               "call": "optimize(pinned numpy views, out=pinned numpy view)"}},
            "gpu_launches": int(launches),
            "roofline": roofline,
            "other_workloads": other_workloads,
            "cpu_baseline": cpu,
            "clocks": clocks,
        }
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if result is not None:
        print(json.dumps(result))


def multipair_batch(rank, world, device, pairs=64, size=128):
    """BASELINE.json configs[3]: 64 independent 128^3 pairs, sharded round-robin over the ranks (no data-path collective),
    each rank's share in ONE lsf_hier_optimize_3d_batch call (pairs advance in lockstep, per-pair convergence). Seconds =
    max over ranks."""
    import torch
    import lsf_b200
    from lsf_b200 import multigpu, synthetic
    rng = np.random.default_rng(1234)
    shifts = rng.uniform(-3, 3, size=(pairs, 3)) + np.array([2.5, -1.5, 1.0])
    indices = multigpu.pair_indices_of_rank(pairs, rank, world)
    fields = [synthetic.sphere_plane_pair_3d(size, shift=tuple(shifts[i]), xp=torch, device=device) for i in indices]
    canonical = torch.stack([f[0] for f in fields])
    live = torch.stack([f[1] for f in fields])
    del fields
    optimizer = lsf_b200.HierarchicalOptimizer3d(**optimizer_kwargs())
    best = None
    for _ in range(3):
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
        optimizer.optimize_batch(canonical, live)
        stop.record()
        torch.cuda.synchronize()
        seconds = multigpu.max_over_ranks(start.elapsed_time(stop) * 1e-3, device)
        best = seconds if best is None else min(best, seconds)
    counts = optimizer.get_per_pair_iteration_counts()
    level_voxels = [(size >> (len(counts[0]) - 1 - level)) ** 3 for level in range(len(counts[0]))]
    updates = multigpu.sum_over_ranks(sum(c * v for pair in counts for c, v in zip(pair, level_voxels)), device)
    return {"workload": "multipair batch: %d independent %d^3 pairs, %d per GPU, one batched call per rank" % (pairs, size, len(indices)),
            "n_gpus": world, "seconds": best, "pairs_per_s": pairs / best, "value": updates / best, "unit": UNIT}


def slab_volume(size, iterations, rank, world, device):
    """BASELINE.json configs[4]: ONE size^3 pair split into slabs along axis 0 over the ranks, halo exchange of the
    Sobolev radius per iteration through peer memory (levelsetfusion-python_b200/slab.py, csrc/slab_peer.cu). Every rank
    generates its own planes."""
    import torch
    import lsf_b200
    from lsf_b200 import multigpu, slab, synthetic
    kwargs = dict(optimizer_kwargs(), maximum_iteration_count=iterations)
    optimizer = lsf_b200.HierarchicalOptimizer3d(**kwargs)
    sharded = slab.SlabHierarchicalOptimizer3d(optimizer, pack_halo=32)
    plan = sharded.plan((size, size, size), rank, world)
    own_lo, own_hi = plan.own_range()
    live_lo, live_hi = plan.live_range()
    lo, hi = min(own_lo, live_lo), max(own_hi, live_hi)
    canonical_part, live_part = synthetic.sphere_plane_pair_3d(size, xp=torch, device=device, planes=(lo, hi))
    canonical_slab = canonical_part[own_lo - lo:own_hi - lo].contiguous()
    live_region = live_part[live_lo - lo:live_hi - lo].contiguous()
    del canonical_part, live_part
    best = None
    for _ in range(2):
        torch.distributed.barrier()
        torch.cuda.synchronize()
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
        sharded.optimize(canonical_slab, live_region, (size, size, size))
        stop.record()
        torch.cuda.synchronize()
        seconds = multigpu.max_over_ranks(start.elapsed_time(stop) * 1e-3, device)
        best = seconds if best is None else min(best, seconds)
    updates = sum((size >> (plan.level_count - 1 - level)) ** 3 * count
                  for level, count in enumerate(sharded.iteration_counts))
    counts = list(sharded.iteration_counts)
    halo_bytes = sharded.exchanged_bytes
    exchange = "peer memory (lsf_slab_exchange: one kernel per exchange stores into the neighbours' halo planes over NVLink)" \
        if sharded._peer_exchange is not None else "torch.distributed send/recv + all_reduce"
    sharded.close()
    del canonical_slab, live_region, sharded
    torch.cuda.empty_cache()
    return {"workload": "one %d^3 pair, slabs of %d planes per GPU, at most %d iterations per level" % (size, own_hi - own_lo, iterations),
            "n_gpus": world, "seconds": best, "value": updates / best, "unit": UNIT, "iterations_per_level": counts,
            "halo_bytes_sent_per_rank": halo_bytes, "exchange": exchange}


def killingfusion_iteration(lsf_b200, canonical, live, size, peak, iterations=10):
    """ms per iteration of the 3D slavcheva optimizer with the Killing and level-set terms and the 7-tap Sobolev filter
    (CUDA events around optimize() calls of n and 3n iterations; set-up and read-back cancel in the difference),
    voxel-updates/s and the fraction of the HBM roofline at SURVEY.md 8(d)'s 36 B per voxel-update."""
    import torch
    from lsf_b200 import synthetic
    from lsf_b200.slavcheva import SmoothingTermMethod

    def run(n):
        optimizer = lsf_b200.SlavchevaOptimizer3d(max_iterations=n, min_iterations=n, maximum_warp_length_lower_threshold=0.0,
                                                  sobolev_kernel=synthetic.sobolev_kernel_1d(),
                                                  smoothing_term_method=SmoothingTermMethod.KILLING,
                                                  level_set_term_enabled=True)
        best = None
        for _ in range(3):
            start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            start.record()
            optimizer.optimize(live.clone(), canonical)
            stop.record()
            torch.cuda.synchronize()
            best = start.elapsed_time(stop) if best is None else min(best, start.elapsed_time(stop))
        return best

    per_iteration = (run(3 * iterations) - run(iterations)) / (2 * iterations)
    updates = size ** 3 / (per_iteration * 1e-3)
    achieved = 36 * updates / 1e9
    # the default path iterates over the narrow band only: the compulsory bytes of a band-only iteration beside the dense
    # 36 B definition of SURVEY.md 8(d), and the DRAM traffic ncu measured for the four brick kernels of one iteration
    band_fraction = float((~((live.abs() == 1.0) & (canonical.abs() == 1.0))).float().mean())
    band_bytes = 36 * band_fraction * size ** 3
    out = {"ms_per_iteration": round(per_iteration, 4), "value": updates, "unit": UNIT,
           "terms": "data + Killing + level set, 7-tap Sobolev filter, masked re-warp",
           "algorithmic_bytes_per_voxel_update": 36, "achieved": round(achieved, 1), "frac": round(achieved / peak, 4),
           "band_fraction": round(band_fraction, 4), "band_only_compulsory_bytes": int(band_bytes),
           "achieved_band_only": round(band_bytes / (per_iteration * 1e-3) / 1e9, 1),
           "frac_band_only": round(band_bytes / (per_iteration * 1e-3) / 1e9 / peak, 4)}
    if size == 256:
        out["traffic"] = 504100000
        out["traffic_source"] = ("profiles/r2_ncu_headline.md, gpurun_out/r2f/killing.ncu-rep: terms 166.6 + filter passes 85.3 + "
                                 "81.9 + axis 2 / re-warp 170.3 MB of DRAM reads + writes per iteration")
    return out


def tsdf_generation(lsf_b200, size, peak, repeats=20):
    """tsdf.Generator3d.generate of a size^3 field from a synthetic 480 x 640 depth frame (a tilted wall with ripples at
    about 1.5 m) that is resident on the device: ms per call (CUDA events over `repeats` calls after a warm-up), voxels/s
    and the fraction of the HBM roofline at 4 B per voxel (the field is written once; the 0.6 MB image stays in L2),
    for filtering NONE; ms per call of EWA_IMAGE_SPACE on (size/2)^3; the CPU oracle on all cores beside both."""
    import time
    import numpy as np
    import torch
    import oracle
    rows, cols = 480, 640
    v, u = np.mgrid[0:rows, 0:cols].astype(np.float32)
    depth = (1500.0 + 0.35 * (u - 320.0) + 40.0 * np.sin(u / 23.0) * np.cos(v / 31.0)).astype(np.uint16)
    image = torch.from_numpy(depth.view(np.int16)).cuda()
    intrinsics = np.array([[700.0, 0.0, 320.0], [0.0, 700.0, 240.0], [0.0, 0.0, 1.0]], dtype=np.float32)
    out = {}
    for name, method, n in (("none", lsf_b200.tsdf.FilteringMethod.NONE, size),
                            ("ewa_image_space", lsf_b200.tsdf.FilteringMethod.EWA_IMAGE_SPACE, size // 2)):
        offset = (-n // 2, -n // 2, 375 - n // 2)
        generator = lsf_b200.tsdf.Generator3d(lsf_b200.tsdf.Parameters3d(
            projection_matrix=intrinsics, array_offset=lsf_b200.Vector3i(*offset), field_shape=lsf_b200.Vector3i(n, n, n),
            interpolation_method=method))
        field = generator.generate(image)
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        start.record()
        for _ in range(repeats):
            field = generator.generate(image)
        stop.record()
        torch.cuda.synchronize()
        ms = start.elapsed_time(stop) / repeats
        t0 = time.perf_counter()
        expected = oracle.tsdf_generate(depth, np.identity(4, dtype=np.float32), 3, intrinsics, offset, (n, n, n),
                                        filtering_method=int(method))
        cpu_ms = 1e3 * (time.perf_counter() - t0)
        difference = float(np.abs(field.cpu().numpy() - expected).max())
        achieved = 4 * n ** 3 / (ms * 1e-3) / 1e9
        out[name] = {"volume": [n, n, n], "ms_per_call": round(ms, 4), "voxels_per_s": n ** 3 / (ms * 1e-3),
                     "algorithmic_bytes_per_voxel": 4, "achieved": round(achieved, 1), "frac": round(achieved / peak, 4),
                     "cpu_oracle_ms": round(cpu_ms, 2), "cpu_cores": os.cpu_count(), "max_abs_difference_to_oracle": difference,
                     "band_fraction": float((np.abs(expected) < 1).mean())}
    return out


def optimizers_2d(lsf_b200, size=128, repeats=5):
    """BASELINE.json configs[0] and the reference's 2D class: one size x size pair from numpy arrays (host staging inside),
    ms per optimize() of SobolevOptimizer2d with the reference experiment's parameters (experiment/
    singleframe_experiment.py:91-116) and of HierarchicalOptimizer2d (4 levels x 100 iterations, data term), the CPU
    oracle beside them, results compared bit for bit. 2D fields are launch-bound: both run an iteration loop (a polling chunk /
    a pyramid level) as ONE launch of one thread-block cluster that keeps the fields in distributed shared memory
    (csrc/slavcheva_persistent.cu k_slav_strips, csrc/hier2d_persistent.cu k_hier_level2d_strips)."""
    import time
    import numpy as np
    import torch
    import oracle
    from lsf_b200 import synthetic
    canonical, live = synthetic.circle_line_pair_2d(size, shift=(5.0, -3.0), line_shift=-4.0)
    kernel = synthetic.sobolev_kernel_1d()

    def best_of(call):
        call()
        torch.cuda.synchronize()
        best = None
        for _ in range(repeats):
            t0 = time.perf_counter()
            result = call()
            torch.cuda.synchronize()
            seconds = time.perf_counter() - t0
            best = seconds if best is None else min(best, seconds)
        return 1e3 * best, result

    shared, sobolev = lsf_b200.SharedParameters.get_instance(), lsf_b200.SobolevParameters.get_instance()
    saved = dict(vars(shared)), dict(vars(sobolev))
    try:
        shared.maximum_iteration_count = 100
        shared.maximum_warp_length_lower_threshold = 0.05
        sobolev.set_sobolev_kernel(kernel)
        optimizer = lsf_b200.SobolevOptimizer2d()
        ms, result = best_of(lambda: optimizer.optimize(live.copy(), canonical))
        iterations = optimizer.get_iteration_count()
    finally:
        vars(shared).update(saved[0])
        vars(sobolev).update(saved[1])
    t0 = time.perf_counter()
    expected = oracle.slavcheva_optimize(live, canonical, semantics=0, max_iterations=100,
                                         maximum_warp_length_lower_threshold=0.05, sobolev_kernel=kernel)
    cpu_ms = 1e3 * (time.perf_counter() - t0)
    out = {"sobolevfusion2d_%d" % size: {
        "ms_per_optimize": round(ms, 3), "iterations": int(iterations), "cpu_oracle_ms": round(cpu_ms, 3),
        "identical_to_oracle": bool(np.array_equal(result, expected["live"])) and iterations == expected["iterations"]}}
    kwargs = dict(tikhonov_term_enabled=False, gradient_kernel_enabled=False, maximum_chunk_size=8, rate=0.2,
                  maximum_iteration_count=100, maximum_warp_update_threshold=0.001)
    hierarchical = lsf_b200.HierarchicalOptimizer2d(**kwargs)
    ms, warp = best_of(lambda: hierarchical.optimize(canonical, live))
    t0 = time.perf_counter()
    expected = oracle.hier_optimize(canonical, live, **kwargs)
    cpu_ms = 1e3 * (time.perf_counter() - t0)
    out["hierarchical2d_%d" % size] = {
        "ms_per_optimize": round(ms, 3), "iterations_per_level": hierarchical.get_per_level_iteration_counts(),
        "cpu_oracle_ms": round(cpu_ms, 3), "identical_to_oracle": bool(np.array_equal(warp, expected["warp"]))}
    return out


def cpu_baseline_sample(size, kwargs, iterations=3):
    """Times the CPU oracle (kind "port": restatement of the reference's loops with OpenMP at the same sites) on
    `iterations` finest-level iterations of the same 256^3 pair, all host threads."""
    import oracle
    from lsf_b200 import synthetic
    oracle.use_all_cores()
    canonical, live = synthetic.sphere_plane_pair_3d(size)
    seconds = oracle.hier_time_iterations3d(canonical, live, iterations, **kwargs)
    return {"value": iterations * size ** 3 / seconds, "unit": UNIT, "cores": oracle.num_threads(), "kind": "port",
            "sample": "%d finest-level iterations (%d^3) of the same workload after 1 warm-up iteration; "
                      "%.2f s" % (iterations, size, seconds)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle
    from lsf_b200 import synthetic
    oracle.use_all_cores()
    size = args.size
    kwargs = optimizer_kwargs()
    canonical, live = synthetic.sphere_plane_pair_3d(size)
    iterations_per_step = 2
    for _ in range(min(args.warmup, 1)):
        oracle.hier_time_iterations3d(canonical, live, 1, **kwargs)
    total = 0.0
    for _ in range(args.steps):
        total += oracle.hier_time_iterations3d(canonical, live, iterations_per_step, **kwargs)
    value = args.steps * iterations_per_step * size ** 3 / total
    sample = ("each step = %d finest-level iterations (%d^3) of the workload on the CPU oracle (restatement of the "
              "reference's C++/OpenMP loops; the reference itself needs Eigen+Boost.Python and cannot be built here)"
              % (iterations_per_step, size))
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": int(os.environ.get("WORLD_SIZE", "1")),
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(size, kwargs, int(os.environ.get("WORLD_SIZE", "1"))),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": oracle.num_threads(), "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument("--gpus", type=int, default=1)
    parser.add_argument("--steps", type=int, default=5)
    parser.add_argument("--warmup", type=int, default=3)
    parser.add_argument("--impl", default="ours", choices=["ours", "reference"])
    parser.add_argument("--size", type=int, default=256)
    args = parser.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
