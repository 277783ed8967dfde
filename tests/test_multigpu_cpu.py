"""Host-side logic of the multi-GPU drivers, exercised without a GPU: world_size-2 (and 3) process groups over gloo on
127.0.0.1. Covers the pair sharding + gather of multigpu.optimize_pairs, the slab geometry of slab.SlabPlan and the
neighbour halo exchange / max-reduction of slab.DistExchange (the same code runs over NCCL on the GPU box)."""
import os
import socket

import numpy as np
import pytest


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world_size, port, name, queue):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world_size),
                      LOCAL_RANK=str(rank))
    import torch
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    try:
        queue.put((rank, globals()[name](rank, world_size)))
    finally:
        dist.destroy_process_group()


def run_ranks(name, world_size):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    queue = ctx.SimpleQueue()
    port = free_port()
    procs = [ctx.Process(target=_worker, args=(r, world_size, port, name, queue)) for r in range(world_size)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    out = dict(queue.get() for _ in range(world_size))
    return [out[r] for r in range(world_size)]


# ----------------------------------------------------------------------------- independent frame pairs
def pairs_case(rank, world_size):
    from lsf_b200 import multigpu
    loaded = []

    def load_pair(index):
        loaded.append(index)
        return np.full((2, 2), index, np.float32), np.full((2, 2), -index, np.float32)

    def optimize(canonical, live):
        return {"warp": canonical - live, "rank": rank}

    results = multigpu.optimize_pairs(optimize, 7, load_pair)
    slowest = multigpu.max_over_ranks(float(rank + 1))
    total = multigpu.sum_over_ranks(float(len(loaded)))
    return [float(r["warp"][0, 0]) for r in results], [r["rank"] for r in results], loaded, slowest, total


@pytest.mark.parametrize("world_size", [2, 3])
def test_optimize_pairs_shards_and_gathers(world_size):
    from lsf_b200 import multigpu
    outputs = run_ranks("pairs_case", world_size)
    for rank, (values, owners, loaded, slowest, total) in enumerate(outputs):
        assert values == [2.0 * i for i in range(7)]                       # every rank sees every result, in order
        assert owners == [i % world_size for i in range(7)]                # round-robin ownership
        assert loaded == multigpu.pair_indices_of_rank(7, rank, world_size)  # a rank loads only its own pairs
        assert slowest == float(world_size) and total == 7.0


def test_pair_indices_cover_everything_once():
    from lsf_b200 import multigpu
    for pairs in (0, 1, 5, 64):
        for world_size in (1, 2, 4, 8):
            seen = sorted(i for r in range(world_size) for i in multigpu.pair_indices_of_rank(pairs, r, world_size))
            assert seen == list(range(pairs))
    with pytest.raises(ValueError):
        multigpu.pair_indices_of_rank(4, 2, 2)
    single = multigpu.optimize_pairs(lambda c, l: c + l, 3, lambda i: (i, 10 * i), rank=0, world_size=1)
    assert single == [0, 11, 22]


# ----------------------------------------------------------------------------- slab geometry
def test_slab_plan_geometry():
    from lsf_b200 import slab
    shape = (1024, 1024, 1024)
    for world_size in (2, 4, 8):
        plans = [slab.SlabPlan(shape, r, world_size, maximum_chunk_size=8, radius=3, tikhonov=True, pack_halo=32)
                 for r in range(world_size)]
        for level in range(4):
            geometries = [p.levels[level] for p in plans]
            X = geometries[0].X_global
            assert X == 1024 >> (3 - level)
            assert geometries[0].own_lo == 0 and geometries[-1].own_hi == X
            for a, b in zip(geometries, geometries[1:]):
                assert a.own_hi == b.own_lo                          # the slabs tile the level
                assert a.halo_hi == b.halo_lo == 3                   # both sides of a cut carry the halo
            assert geometries[0].halo_lo == 0 and geometries[-1].halo_hi == 0  # no halo at the volume border
            for g in geometries:
                assert g.planes == g.halo_lo + (g.own_hi - g.own_lo) + g.halo_hi
                assert g.x_origin + g.own_begin == g.own_lo
                assert g.pack_lo <= g.own_lo and g.pack_hi >= g.own_hi
                assert g.voxels * 3 < 2 ** 31                        # 32-bit voxel indices inside the kernels
            if level > 0:
                for coarse, fine in zip([p.levels[level - 1] for p in plans], geometries):
                    assert fine.own_lo == 2 * coarse.own_lo and fine.own_hi == 2 * coarse.own_hi
                    assert fine.pack_lo <= 2 * coarse.pack_lo and fine.pack_hi >= 2 * coarse.pack_hi
        for p in plans:
            lo, hi = p.live_range()
            assert lo <= max(p.levels[-1].pack_lo - 1, 0) and hi >= min(p.levels[-1].pack_hi + 1, 1024)
    with pytest.raises(ValueError):
        slab.SlabPlan((64, 64, 64), 0, 8, maximum_chunk_size=8, radius=3)   # coarsest level: 8 planes / 8 ranks < halo
    with pytest.raises(RuntimeError):
        slab.SlabPlan((60, 64, 64), 0, 2, maximum_chunk_size=8)


# ----------------------------------------------------------------------------- halo exchange over a process group
def exchange_case(rank, world_size):
    import torch
    from lsf_b200 import slab
    plan = slab.SlabPlan((16 * world_size, 8, 8), rank, world_size, maximum_chunk_size=2, radius=3, tikhonov=True)
    g = plan.levels[-1]
    field = torch.full((3, g.planes, g.Y, g.Z), -1.0)
    for plane in range(g.own_begin, g.own_end):  # owned planes carry their global plane index (+ component / 10)
        for c in range(3):
            field[c, plane] = float(g.x_origin + plane) + c / 10.0
    exchange = slab.DistExchange()
    exchange.halos([field], [g], 3)
    ok = True
    for plane in range(g.planes):
        for c in range(3):
            ok = ok and bool((field[c, plane] == float(g.x_origin + plane) + c / 10.0).all())
    narrow = field.clone()
    narrow[:, :g.own_begin] = -2.0
    narrow[:, g.own_end:] = -2.0
    exchange.halos([narrow], [g], 1)  # the one-plane exchange touches only the planes next to the cut
    touched = [p for p in range(g.planes) if float(narrow[0, p, 0, 0]) != -2.0]
    expected = list(range(max(g.own_begin - 1, 0), min(g.own_end + 1, g.planes)))
    slots = torch.zeros(4, dtype=torch.int32)
    slots[2] = int(np.float32(rank + 0.5).view(np.int32))
    exchange.reduce_max([slots], 2)
    return ok, touched == expected, float(np.int32(int(slots[2])).view(np.float32)), int(slots[1])


@pytest.mark.parametrize("world_size", [2, 3])
def test_dist_exchange_halos_and_max(world_size):
    for ok, narrow_ok, reduced, untouched in run_ranks("exchange_case", world_size):
        assert ok and narrow_ok
        assert reduced == world_size - 0.5 and untouched == 0


def test_local_exchange_matches_dist_semantics():
    import torch
    from lsf_b200 import slab
    world_size = 3
    plans = [slab.SlabPlan((48, 4, 4), r, world_size, maximum_chunk_size=2, radius=2, tikhonov=True) for r in range(3)]
    geometries = [p.levels[-1] for p in plans]
    fields = []
    for g in geometries:
        f = torch.full((3, g.planes, g.Y, g.Z), -1.0)
        for plane in range(g.own_begin, g.own_end):
            f[:, plane] = float(g.x_origin + plane)
        fields.append(f)
    slab.LocalExchange().halos(fields, geometries, 2)
    for f, g in zip(fields, geometries):
        for plane in range(g.planes):
            assert bool((f[:, plane] == float(g.x_origin + plane)).all())
    slots = [torch.tensor([0, int(np.float32(v).view(np.int32))], dtype=torch.int32) for v in (0.25, 3.0, 1.5)]
    slab.LocalExchange().reduce_max(slots, 1)
    assert all(float(np.int32(int(s[1])).view(np.float32)) == 3.0 for s in slots)
