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


def bind_to_gpu_numa_node(device_index=None):
    """Pins the calling thread (and the threads it starts afterwards: the library's staging workers) to the CPUs NVML names
    as closest to the CUDA device, so that the page-locked staging buffers allocated from now on and the host copies live on
    the GPU's NUMA node. With one process per GPU and every rank moving its inputs and results at the same time (335 MB per
    256^3 optimize()), unbound ranks send half of that traffic across the socket interconnect. Returns the CPU set, or None
    when NVML / the affinity call is not available (nothing is changed then)."""
    try:
        import pynvml
        import torch
        index = torch.cuda.current_device() if device_index is None else int(device_index)
        pynvml.nvmlInit()
        uuid = str(torch.cuda.get_device_properties(index).uuid)
        handle = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(handle, words)
        cpus = {64 * w + b for w, word in enumerate(mask) for b in range(64) if (int(word) >> b) & 1}
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:  # no NVML, no such call on this platform, restricted cpuset: run unbound
        return None


def pair_indices_of_rank(pair_count, rank, world_size):
    """Round-robin shard of the pair list: rank r owns pairs r, r + world, r + 2*world, ... (neighbouring frames of a
    sequence converge in similar iteration counts, so round-robin balances better than contiguous blocks)."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("invalid rank %d of %d" % (rank, world_size))
    return list(range(rank, int(pair_count), world_size))


def _run_local_pairs(optimize, indices, load_pair, streams):
    """This rank's pairs, `streams` of them at a time: each worker thread owns one CUDA stream (the library takes the
    stream of the calling thread and keeps no shared mutable state), so the launch-bound coarse pyramid levels of one
    pair overlap the bandwidth-bound fine level of another. `optimize` may be a callable or a factory result per worker:
    if it has a `clone_for_worker()` method it is cloned per thread (optimizer objects keep per-call reports)."""
    if streams <= 1 or len(indices) <= 1:
        return {index: optimize(*load_pair(index)) for index in indices}
    import threading
    import torch
    local, errors, lock = {}, [], threading.Lock()
    queue = list(indices)
    device = torch.cuda.current_device()

    def worker():
        torch.cuda.set_device(device)
        stream = torch.cuda.Stream()
        run = optimize.clone_for_worker() if hasattr(optimize, "clone_for_worker") else optimize
        try:
            with torch.cuda.stream(stream):
                while True:
                    with lock:
                        if not queue:
                            break
                        index = queue.pop(0)
                    result = run(*load_pair(index))
                    with lock:
                        local[index] = result
            stream.synchronize()
        except BaseException as error:  # re-raised in the caller's thread
            with lock:
                errors.append(error)

    threads = [threading.Thread(target=worker) for _ in range(min(streams, len(indices)))]
    torch.cuda.current_stream().synchronize()  # inputs produced on the caller's stream are complete
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    return local


class PerWorkerOptimizer:
    """Wraps an optimizer factory for optimize_pairs(..., streams=k): every worker thread gets its own optimizer object.
    `call` maps (optimizer, canonical, live) to the per-pair result (default: the warp field)."""

    def __init__(self, factory, call=None):
        self.factory = factory
        self.call = call or (lambda optimizer, canonical, live: optimizer.optimize(canonical, live))
        self._optimizer = None

    def clone_for_worker(self):
        return PerWorkerOptimizer(self.factory, self.call)

    def __call__(self, canonical, live):
        if self._optimizer is None:
            self._optimizer = self.factory()
        return self.call(self._optimizer, canonical, live)


def optimize_pairs(optimize, pair_count, load_pair, rank=None, world_size=None, gather=True, group=None, streams=1):
    """Runs `optimize(canonical, live)` on this rank's share of `pair_count` independent frame pairs.

    streams    -- pairs in flight per GPU (worker threads with one CUDA stream each, see _run_local_pairs); with more
                  than one, pass a PerWorkerOptimizer (or any callable that is safe to call from several threads)

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
    local = _run_local_pairs(optimize, pair_indices_of_rank(pair_count, rank, world_size), load_pair, int(streams))
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
