"""Multi-GPU drivers of the path: one process per GPU over torch.distributed (NCCL on the GPUs; the
host-side logic is backend-agnostic and is tested with gloo on CPU).

* Independent frame pairs -- the reference's multipair loop (run_hierarchical_optimizer3d_multipair.py:403-406,
  experiment/multiframe_experiment.py:185-233) calls optimize() once per pair in a serial Python loop. The pairs are
  independent, so they are sharded over the ranks with NO collective on the data path; the per-pair results are
  gathered on request at the end.
* A single large volume is decomposed into slabs along numpy axis 0: see slab.py.
"""
import os


def world():
    """(rank, world size, local rank) of this process; (0, 1, 0) outside torch.distributed.run"""
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size(), int(os.environ.get("LOCAL_RANK", dist.get_rank()))
    except ImportError:
        pass
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def pair_indices_of_rank(pair_count, rank, world_size):
    """Round-robin shard of the pair list: rank r owns pairs r, r + world, r + 2*world, ... (neighbouring frames of a
    sequence converge in similar iteration counts, so round-robin balances better than contiguous blocks)."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("invalid rank %d of %d" % (rank, world_size))
    return list(range(rank, int(pair_count), world_size))


def optimize_pairs(optimize, pair_count, load_pair, rank=None, world_size=None, gather=True, group=None):
    """Runs `optimize(canonical, live)` on this rank's share of `pair_count` independent frame pairs.

    optimize   -- callable(canonical, live) -> result (e.g. HierarchicalOptimizer3d(...).optimize, or a lambda that also
                  returns the optimizer's per-level reports)
    load_pair  -- callable(index) -> (canonical, live); only called for the pairs this rank owns, so a rank never
                  touches the inputs of another rank
    gather     -- True: every rank returns the list of all results in pair order (all_gather_object at the end, off
                  the data path); False: returns {pair index: result} of the local pairs only.
    """
    detected_rank, detected_world, _ = world()
    rank = detected_rank if rank is None else rank
    world_size = detected_world if world_size is None else world_size
    local = {}
    for index in pair_indices_of_rank(pair_count, rank, world_size):
        canonical, live = load_pair(index)
        local[index] = optimize(canonical, live)
    if not gather:
        return local
    if world_size == 1:
        return [local[i] for i in range(pair_count)]
    import torch.distributed as dist
    shards = [None] * world_size
    dist.all_gather_object(shards, local, group=group)
    merged = {}
    for shard in shards:
        merged.update(shard)
    missing = [i for i in range(pair_count) if i not in merged]
    if missing:
        raise RuntimeError("pairs %s were not processed by any rank" % missing)
    return [merged[i] for i in range(pair_count)]


def max_over_ranks(value, device=None, group=None):
    """max of a python float over all ranks (timing: a multi-GPU number is the slowest rank's)"""
    _, world_size, _ = world()
    if world_size == 1:
        return float(value)
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t[0])


def sum_over_ranks(value, device=None, group=None):
    _, world_size, _ = world()
    if world_size == 1:
        return float(value)
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return float(t[0])
