#!/usr/bin/env python
"""The reference's multi-pair experiment (run_hierarchical_optimizer3d_multipair.py) on the GPU path:

    python [-m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 ...] tools/run_multipair.py \
        --data DIR_WITH_data_F_R.npz --out OUT_DIR [--streams 4] [--tikhonov] [--kernel-size 7 --kernel-strength 0.1] ...

Reads the reference's pair cache, optimises the pairs over the ranks and `--streams` pairs in flight per GPU, and
writes convergence_reports.pk (+ .xlsx / .csv) and analysis.txt in the reference's formats. `--synthetic K` first fills
the cache with K synthetic 128^3 sphere/plane pairs (BASELINE.json configs[3] shape)."""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import lsf_b200
from lsf_b200 import multigpu, multipair, synthetic


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument("--data", required=True)
    parser.add_argument("--out", required=True)
    parser.add_argument("--streams", type=int, default=4)
    parser.add_argument("--synthetic", type=int, default=0)
    parser.add_argument("--size", type=int, default=128)
    parser.add_argument("--tikhonov", action="store_true")
    parser.add_argument("--tikhonov-strength", type=float, default=0.1)
    parser.add_argument("--no-kernel", action="store_true")
    parser.add_argument("--kernel-size", type=int, default=7)
    parser.add_argument("--kernel-strength", type=float, default=0.1)
    parser.add_argument("--rate", type=float, default=0.1)
    parser.add_argument("--maximum-iteration-count", type=int, default=100)
    parser.add_argument("--maximum-warp-update-threshold", type=float, default=0.01)
    parser.add_argument("--start-from-index", type=int, default=0)
    parser.add_argument("--stop-before-index", type=int, default=10000000)
    parser.add_argument("--save-warps", action="store_true")
    parser.add_argument("--save-telemetry", action="store_true",
                        help="reference --save_telemetry: per-level iteration data of every pair -> telemetry/pair_*/telemetry_log.npz")
    args = parser.parse_args()
    rank, world_size = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    if world_size > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    if args.synthetic > 0 and rank == 0:
        rng = np.random.default_rng(1234)
        for i in range(args.synthetic):
            shift = tuple(np.array([2.5, -1.5, 1.0]) + rng.uniform(-3, 3, 3))
            canonical, live = synthetic.sphere_plane_pair_3d(args.size, shift=shift)
            multipair.save_pair(args.data, i, 214, canonical, live)
    if world_size > 1:
        dist.barrier()
    logging = lsf_b200.HierarchicalOptimizer3d.LoggingParameters(collect_per_level_convergence_reports=True,
                                                                 collect_per_level_iteration_data=args.save_telemetry)
    factory = lambda: lsf_b200.HierarchicalOptimizer3d(
        tikhonov_term_enabled=args.tikhonov, gradient_kernel_enabled=not args.no_kernel, maximum_chunk_size=8,
        rate=args.rate, maximum_iteration_count=args.maximum_iteration_count,
        maximum_warp_update_threshold=args.maximum_warp_update_threshold, tikhonov_strength=args.tikhonov_strength,
        kernel=synthetic.sobolev_kernel_1d(args.kernel_size, args.kernel_strength), logging_parameters=logging)
    t0 = time.perf_counter()
    table = multipair.run_multipair(args.data, args.out, factory, args.streams, args.start_from_index,
                                    args.stop_before_index, args.save_warps, save_telemetry=args.save_telemetry)
    torch.cuda.synchronize()
    seconds = multigpu.max_over_ranks(time.perf_counter() - t0, torch.device("cuda", torch.cuda.current_device()))
    if rank == 0:
        print("%d pairs on %d GPU(s) x %d streams in %.2f s (%.1f pairs/s, files included); reports in %s"
              % (len(table), world_size, args.streams, seconds, len(table) / seconds, args.out))
    if world_size > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
