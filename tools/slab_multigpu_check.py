#!/usr/bin/env python
"""Slab-sharded hierarchical optimisation of ONE volume over the GPUs of a box (one process per GPU, NCCL):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tools/slab_multigpu_check.py [--size 256] [--check] [--iterations 20]

--check: rank 0 also runs the whole-volume optimizer and compares the gathered result bit for bit (sizes that fit one
GPU). Prints one JSON line with the timing (CUDA events, max over ranks), voxel-updates/s and halo traffic."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import lsf_b200
from lsf_b200 import slab, synthetic, multigpu


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument("--size", type=int, default=256)
    parser.add_argument("--iterations", type=int, default=20)
    parser.add_argument("--check", action="store_true")
    parser.add_argument("--repeat", type=int, default=2)
    parser.add_argument("--exchange", choices=["peer", "dist"], default="peer",
                        help="peer: kernels storing into the neighbours' memory (csrc/slab_peer.cu); dist: NCCL send/recv + all_reduce")
    args = parser.parse_args()
    rank, world_size = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world_size > 1:
        dist.init_process_group("nccl", device_id=device)
    size = args.size
    optimizer = lsf_b200.HierarchicalOptimizer3d(tikhonov_term_enabled=True, tikhonov_strength=0.1,
                                                 gradient_kernel_enabled=True, kernel=synthetic.sobolev_kernel_1d(),
                                                 maximum_chunk_size=8, maximum_iteration_count=args.iterations,
                                                 maximum_warp_update_threshold=0.01)
    sharded = slab.SlabHierarchicalOptimizer3d(optimizer, pack_halo=32, exchange=args.exchange)
    plan = sharded.plan((size, size, size), rank, world_size)
    own_lo, own_hi = plan.own_range()
    live_lo, live_hi = plan.live_range()
    # every rank generates only the planes it needs (the full 1024^3 pair would be 8 GiB)
    lo = min(own_lo, live_lo)
    hi = max(own_hi, live_hi)
    canonical_part, live_part = synthetic.sphere_plane_pair_3d(size, xp=torch, device=device, planes=(lo, hi))
    canonical_slab = canonical_part[own_lo - lo:own_hi - lo].contiguous()
    live_region = live_part[live_lo - lo:live_hi - lo].contiguous()
    del canonical_part, live_part
    times = []
    for _ in range(args.repeat):
        if world_size > 1:
            dist.barrier()
        torch.cuda.synchronize()
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
        warp = sharded.optimize(canonical_slab, live_region, (size, size, size))
        stop.record()
        torch.cuda.synchronize()
        times.append(multigpu.max_over_ranks(start.elapsed_time(stop), device=device))
    updates = sum((size >> (plan.level_count - 1 - level)) ** 3 * count
                  for level, count in enumerate(sharded.iteration_counts))
    result = {"workload": "hierarchical3d_%d_slab_sharded" % size, "n_gpus": world_size, "exchange": args.exchange,
              "ms": times[-1], "ms_all_repeats": times,
              "voxel_updates_per_s": updates / (times[-1] * 1e-3), "iterations_per_level": sharded.iteration_counts,
              "halo_bytes_sent_per_rank": sharded.exchanged_bytes, "planes_per_rank": own_hi - own_lo}
    if args.check:
        gathered = [torch.empty_like(warp) for _ in range(world_size)] if world_size > 1 else [warp]
        if world_size > 1:
            dist.all_gather(gathered, warp)
        if rank == 0:
            canonical, live = synthetic.sphere_plane_pair_3d(size, xp=torch, device=device)
            whole = optimizer.optimize(canonical, live)
            result["bit_identical_to_whole_volume"] = bool((torch.cat(gathered, dim=0) == whole).all())
            result["iteration_counts_equal"] = sharded.iteration_counts == optimizer.get_per_level_iteration_counts()
    if rank == 0:
        print(json.dumps(result))
    sharded.close()
    if world_size > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
