#!/usr/bin/env python
"""BASELINE.json configs[3]: a batch of independent 128^3 frame pairs, data-parallel over the GPUs of a box
(torch.distributed.run, one process per GPU) and, inside a GPU, `streams` pairs in flight.

    python [-m torch.distributed.run --nproc-per-node N ...] tools/multipair_times.py [--pairs 64] [--size 128] [--streams 1,2,4]

Prints one JSON line per stream count: pairs/s and voxel-updates/s over all ranks (CUDA-synchronised wall time, max
over ranks)."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import bench
import lsf_b200
from lsf_b200 import multigpu, synthetic


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument("--pairs", type=int, default=64)
    parser.add_argument("--size", type=int, default=128)
    parser.add_argument("--streams", default="1,2,4")
    args = parser.parse_args()
    rank, world_size = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world_size > 1:
        dist.init_process_group("nccl", device_id=device)
    rng = np.random.default_rng(1234)
    shifts = rng.uniform(-3, 3, size=(args.pairs, 3)) + np.array([2.5, -1.5, 1.0])
    pairs = {}
    for index in multigpu.pair_indices_of_rank(args.pairs, rank, world_size):
        pairs[index] = synthetic.sphere_plane_pair_3d(args.size, shift=tuple(shifts[index]), xp=torch, device=device)
    kwargs = bench.optimizer_kwargs()

    def call(optimizer, canonical, live):
        optimizer.optimize(canonical, live)
        return bench.voxel_updates(optimizer.get_per_level_convergence_reports())

    for streams in [int(v) for v in args.streams.split(",")]:
        run = multigpu.PerWorkerOptimizer(lambda: lsf_b200.HierarchicalOptimizer3d(**kwargs), call)
        best, updates = None, 0
        for _ in range(2):
            if world_size > 1:
                dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            local = multigpu.optimize_pairs(run, args.pairs, lambda i: pairs[i], rank, world_size, gather=False,
                                            streams=streams)
            torch.cuda.synchronize()
            seconds = multigpu.max_over_ranks(time.perf_counter() - t0, device)
            best = seconds if best is None else min(best, seconds)
            updates = multigpu.sum_over_ranks(sum(local.values()), device)
        if rank == 0:
            print(json.dumps({"workload": "multipair_%dx%d^3" % (args.pairs, args.size), "n_gpus": world_size,
                              "streams_per_gpu": streams, "seconds": best, "pairs_per_s": args.pairs / best,
                              "voxel_updates_per_s": updates / best}))
    # the same batch through lsf_hier_optimize_3d_batch: one call per rank, no Python threads
    indices = multigpu.pair_indices_of_rank(args.pairs, rank, world_size)
    if indices:
        canonical = torch.stack([pairs[i][0] for i in indices])
        live = torch.stack([pairs[i][1] for i in indices])
        optimizer = lsf_b200.HierarchicalOptimizer3d(**kwargs)
        best = None
        for _ in range(3):
            if world_size > 1:
                dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            optimizer.optimize_batch(canonical, live)
            torch.cuda.synchronize()
            seconds = multigpu.max_over_ranks(time.perf_counter() - t0, device)
            best = seconds if best is None else min(best, seconds)
        counts = optimizer.get_per_pair_iteration_counts()
        level_voxels = [(args.size >> (len(counts[0]) - 1 - level)) ** 3 for level in range(len(counts[0]))]
        updates = multigpu.sum_over_ranks(sum(c * v for pair in counts for c, v in zip(pair, level_voxels)), device)
        if rank == 0:
            print(json.dumps({"workload": "multipair_%dx%d^3" % (args.pairs, args.size), "n_gpus": world_size,
                              "mode": "lsf_hier_optimize_3d_batch", "seconds": best, "pairs_per_s": args.pairs / best,
                              "voxel_updates_per_s": updates / best}))
    if world_size > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
