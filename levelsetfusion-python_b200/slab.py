"""Slab decomposition of ONE large volume over several GPUs (BASELINE.json configs[4]).

The 3D hierarchical optimizer (reference Optimizer<Tensor3f,Tensor3v3f>, cpp/src/nonrigid_optimization/hierarchical/
optimizer.tpp:83-212) is partitioned along numpy axis 0: rank r owns planes [r*X/W, (r+1)*X/W) of every pyramid level.
Per iteration the ranks exchange
  * `radius` planes of the unfiltered gradient (before the axis-0 pass of the Sobolev filter),
  * 1 plane of the filtered gradient (the Tikhonov term of the next iteration takes its Laplacian),
  * one 4-byte max-reduction for the level-termination test (optimizer.tpp:166-171);
the live level {TSDF, gradient} is iteration-invariant, so every rank keeps its slab of it plus a static gather halo
(`pack_halo` planes at the finest level, halved per coarser level) and the kernels raise a flag if a warp vector ever
reaches beyond it. Per-voxel arithmetic is untouched, hence the sharded result is bit-identical to the whole-volume
optimizer -- that is what the tests assert.

The driver is written over a list of rank states with a pluggable exchange:
  * `PeerExchange` (default with one process per GPU on one NVLink / NVSwitch box): the exchanged fields live in memory
    every rank maps (CUDA IPC); ONE kernel per exchange (csrc/slab_peer.cu, lsf_slab_exchange) stores the boundary planes
    straight into the neighbours' halo planes, signals through a mailbox in peer memory, carries the 4-byte maximum of
    the termination test to all ranks and waits for the neighbours' planes -- no host round trip, no library collective;
    also runs with virtual ranks on one GPU (one stream per rank), which is how the single-GPU tests cover the kernel;
  * `DistExchange` (torch.distributed P2P + all_reduce: NCCL, or gloo in the CPU tests of the exchange logic);
  * `LocalExchange` (all virtual ranks in one process on one stream, plain tensor copies).
Restrictions of the slab mode: NEAREST_AND_AVERAGE resampling, Sobolev kernels of 3/5/7 taps, X / 2^(levels-1)
divisible by the rank count.
"""
import ctypes
import os

import numpy as np

from . import _lib

POLL_CHUNK = 16


class SlabGeometry:
    """Plane bookkeeping of one rank at one pyramid level (all indices along numpy axis 0)."""

    def __init__(self, X_global, Y, Z, rank, world_size, halo, pack_halo):
        if X_global % world_size != 0:
            raise ValueError("level with %d planes is not divisible by %d ranks" % (X_global, world_size))
        self.X_global, self.Y, self.Z = X_global, Y, Z
        self.rank, self.world_size = rank, world_size
        per_rank = X_global // world_size
        self.own_lo, self.own_hi = rank * per_rank, (rank + 1) * per_rank      # global planes owned
        self.halo_lo = halo if rank > 0 else 0                                   # halo planes towards the cuts
        self.halo_hi = halo if rank < world_size - 1 else 0
        if world_size > 1 and per_rank < halo:
            raise ValueError("%d planes per rank are fewer than the halo (%d)" % (per_rank, halo))
        self.planes = self.halo_lo + per_rank + self.halo_hi                     # allocation
        self.x_origin = self.own_lo - self.halo_lo
        self.own_begin, self.own_end = self.halo_lo, self.halo_lo + per_rank     # inside the allocation
        self.pack_lo = max(self.own_lo - pack_halo, 0)
        self.pack_hi = min(self.own_hi + pack_halo, X_global)
        self.pack_planes = self.pack_hi - self.pack_lo
        self.pack_interior_low = int(self.pack_lo > 0)
        self.pack_interior_high = int(self.pack_hi < X_global)

    @property
    def voxels(self):
        return self.planes * self.Y * self.Z

    @property
    def pack_padded_count(self):
        return (self.pack_planes + 4) * (self.Y + 4) * (self.Z + 4)


class SlabPlan:
    """Geometry of one rank at every level (index 0 = coarsest)."""

    def __init__(self, shape, rank, world_size, maximum_chunk_size=8, radius=3, tikhonov=True, pack_halo=32):
        X, Y, Z = (int(d) for d in shape)
        power = int(np.log2(maximum_chunk_size))
        if 2 ** power != maximum_chunk_size:
            raise RuntimeError("The argument 'maximum_chunk_size' must be an integer power of 2, i.e. 4, 8, 16, etc.")
        self.level_count = power + 1
        divisor = 2 ** (self.level_count - 1)
        if X % divisor or Y % divisor or Z % divisor:
            raise RuntimeError("each dimension %s must be divisible by %d for a %d-level pyramid"
                               % ((X, Y, Z), divisor, self.level_count))
        self.shape = (X, Y, Z)
        self.rank, self.world_size = rank, world_size
        self.halo = max(int(radius), 1 if tikhonov else 0)
        self.levels = []
        # gather halo: `pack_halo` planes at the finest level (at least 2 at the coarsest), exactly doubling from level
        # to level so that every finer pack region covers the coarser one it is restricted to
        coarsest_pack_halo = max(int(pack_halo) >> (self.level_count - 1), 2)
        for level in range(self.level_count):
            shrink = 2 ** (self.level_count - 1 - level)
            self.levels.append(SlabGeometry(X // shrink, Y // shrink, Z // shrink, rank, world_size, self.halo,
                                            coarsest_pack_halo << level))
        finest = self.levels[-1]
        self.live_lo = max(finest.pack_lo - 1, 0)
        self.live_hi = min(finest.pack_hi + 1, X)

    def own_range(self):
        """global planes of the canonical field / result this rank owns at full resolution"""
        return self.levels[-1].own_lo, self.levels[-1].own_hi

    def live_range(self):
        """global planes of the live field this rank needs (owned planes + gather halo + 1 for the gradient)"""
        return self.live_lo, self.live_hi


# ------------------------------------------------------------------------------------------------ halo exchange
class LocalExchange:
    """All virtual ranks live in this process: halos are copied between the ranks' tensors."""

    def halos(self, fields, geometries, width):
        for r in range(len(fields) - 1):
            low, high = fields[r], fields[r + 1]
            gl, gh = geometries[r], geometries[r + 1]
            # low rank's last owned planes -> high rank's low halo; high rank's first owned planes -> low rank's high halo
            high[:, gh.own_begin - width:gh.own_begin].copy_(low[:, gl.own_end - width:gl.own_end])
            low[:, gl.own_end:gl.own_end + width].copy_(high[:, gh.own_begin:gh.own_begin + width])

    def reduce_max(self, slots, iteration):
        if len(slots) > 1:
            import torch
            best = slots[0][iteration]
            for s in slots[1:]:
                best = torch.maximum(best, s[iteration])
            for s in slots:
                s[iteration] = best


class DistExchange:
    """One rank per process: neighbour P2P through torch.distributed (NCCL over NVLink / NVSwitch on the GPU box)."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.rank = dist.get_rank(group)
        self.world_size = dist.get_world_size(group)

    def halos(self, fields, geometries, width):
        import torch
        dist = self.dist
        field, g = fields[0], geometries[0]
        ops, received = [], []
        if self.rank > 0:  # low neighbour
            send = field[:, g.own_begin:g.own_begin + width].contiguous()
            recv = torch.empty_like(send)
            ops += [dist.P2POp(dist.isend, send, self.rank - 1, self.group),
                    dist.P2POp(dist.irecv, recv, self.rank - 1, self.group)]
            received.append((recv, slice(g.own_begin - width, g.own_begin)))
        if self.rank < self.world_size - 1:  # high neighbour
            send = field[:, g.own_end - width:g.own_end].contiguous()
            recv = torch.empty_like(send)
            ops += [dist.P2POp(dist.isend, send, self.rank + 1, self.group),
                    dist.P2POp(dist.irecv, recv, self.rank + 1, self.group)]
            received.append((recv, slice(g.own_end, g.own_end + width)))
        if ops:
            for work in dist.batch_isend_irecv(ops):
                work.wait()
            for recv, planes in received:
                field[:, planes].copy_(recv)

    def reduce_max(self, slots, iteration):
        # bit patterns of non-negative floats order like integers: an integer MAX is the float max
        self.dist.all_reduce(slots[0][iteration:iteration + 1], op=self.dist.ReduceOp.MAX, group=self.group)


class _DeviceFloats:
    """raw device memory as an object torch.as_tensor can wrap without a copy"""

    def __init__(self, pointer, count):
        self.__cuda_array_interface__ = {"shape": (int(count),), "typestr": "<f4", "data": (int(pointer), False),
                                         "version": 2}


class PeerExchange:
    """Halo exchange + termination maximum through peer memory (csrc/slab_peer.cu).

    Every rank owns one allocation [field A | field B | mailbox] sized for the finest level; the gradient fields of every
    level are views at offset 0 of A / B (g_pre = A, g_post = B; without a Sobolev kernel the two swap roles every
    iteration, on all ranks alike). `exchange()` enqueues one kernel per local rank; sequence numbers count the exchanges.

    Why a neighbour's stores never hit planes that are still being read (r = this rank, n = a neighbour):
      * r stores the planes of exchange e+1 only after its wait of exchange e has ended, i.e. after n's stores of exchange e,
        which n's stream issues after the phase kernel that last read the halo planes exchange e+1 overwrites
        (g_pre halo: read by phase 2 of the previous iteration, followed by the g_post / maximum exchange;
         g_post halo: read by phase 1 of this iteration, followed by the g_pre exchange; reduction-only exchanges wait for
         all ranks' maxima, which every rank posts after its phase kernel);
      * a level starts by zero-filling g_post ONLY (optimizer.tpp:142-143; its first delivery follows an exchange the
        owner took part in); g_pre is never cleared: every owned plane is written by phase 1 and every halo plane by the
        neighbour before phase 2 reads it -- a neighbour that is already in the next level may deliver g_pre planes before
        the owner gets there;
      * nothing reads the fields between the last exchange of a level and the first kernel of the next one.
    """
    fused = True

    def __init__(self, group=None, virtual_ranks=None):
        import torch
        self.lib = _lib.load()
        self.group = group
        if virtual_ranks is None:
            import torch.distributed as dist
            self.dist = dist
            self.rank, self.world_size = dist.get_rank(group), dist.get_world_size(group)
            self.local_ranks = [self.rank]
        else:
            self.dist = None
            self.rank, self.world_size = 0, int(virtual_ranks)
            self.local_ranks = list(range(self.world_size))
        if self.world_size > _lib.SLAB_MAX_PEERS:
            raise ValueError("PeerExchange supports at most %d ranks" % _lib.SLAB_MAX_PEERS)
        self.device = torch.device("cuda", torch.cuda.current_device())
        self.region_bytes = 0
        self.owned = {}      # rank -> pointer from lsf_peer_alloc
        self.mapped = {}     # rank -> pointer from lsf_peer_open
        self.sequence = 0

    # ---------------------------------------------------------------- memory
    def setup(self, field_floats):
        """(re)allocates for fields of up to `field_floats` floats (the same number on every rank)"""
        region_bytes = (int(field_floats) * 4 + 255) // 256 * 256
        if region_bytes <= self.region_bytes:
            return
        self.close()
        self.region_bytes = region_bytes
        total = 2 * region_bytes + _lib.SLAB_MAILBOX_BYTES
        handles = {}
        for rank in self.local_ranks:
            pointer = ctypes.c_void_p()
            handle = (ctypes.c_ubyte * _lib.PEER_HANDLE_BYTES)()
            _lib.check(self.lib.lsf_peer_alloc(ctypes.c_size_t(total), ctypes.byref(pointer),
                                               handle if self.dist is not None else None))
            self.owned[rank] = pointer.value
            handles[rank] = bytes(handle)
        if self.dist is not None:
            gathered = [None] * self.world_size
            self.dist.all_gather_object(gathered, handles[self.rank], group=self.group)
            for rank, handle in enumerate(gathered):
                if rank == self.rank:
                    continue
                pointer = ctypes.c_void_p()
                raw = (ctypes.c_ubyte * _lib.PEER_HANDLE_BYTES).from_buffer_copy(handle)
                _lib.check(self.lib.lsf_peer_open(raw, ctypes.byref(pointer)))
                self.mapped[rank] = pointer.value
        self.sequence = 0

    def close(self):
        """frees the allocations (collective with one process per GPU: nobody may still be storing into them)"""
        if not self.owned:
            return
        import torch
        torch.cuda.synchronize()
        if self.dist is not None:
            self.dist.barrier(group=self.group)
        for pointer in self.mapped.values():
            self.lib.lsf_peer_close(ctypes.c_void_p(pointer))
        if self.dist is not None:
            self.dist.barrier(group=self.group)
        for pointer in self.owned.values():
            self.lib.lsf_peer_free(ctypes.c_void_p(pointer))
        self.owned, self.mapped, self.region_bytes = {}, {}, 0

    def _peers(self, rank):
        peers = _lib.SlabPeers()
        peers.rank, peers.world_size = rank, self.world_size
        for r in range(self.world_size):
            peers.base[r] = self.owned[r] if r in self.owned else self.mapped[r]
        peers.mailbox_offset = 2 * self.region_bytes
        return peers

    def fields(self, rank, geometry):
        """(g_pre, g_post) of `rank` (a local rank) at a level: views [3][planes][Y][Z] of the two exchanged fields"""
        import torch
        count = 3 * geometry.voxels
        if count * 4 > self.region_bytes:
            raise ValueError("level does not fit the peer allocation")
        base = self.owned[rank]
        shape = (3, geometry.planes, geometry.Y, geometry.Z)
        views = [torch.as_tensor(_DeviceFloats(base + k * self.region_bytes, count), device=self.device).view(shape)
                 for k in range(2)]
        return views[0], views[1]

    # ---------------------------------------------------------------- one exchange on every local rank
    def exchange(self, states, level, which, width, reduce_iteration):
        """which: "pre" / "post" = the field of the states whose boundary planes travel (`width` planes; 0 = none);
        reduce_iteration >= 0: slot of the termination maximum to reduce over the ranks in the same kernel"""
        self.sequence += 1
        for state in states:
            rank = state.plan.rank
            g = state.plan.levels[level]
            field = state.g_pre if which == "pre" else state.g_post
            offset = field.data_ptr() - self.owned[rank]
            low = SlabGeometry(g.X_global, g.Y, g.Z, rank - 1, self.world_size, state.plan.halo, 0) if rank > 0 else None
            high = SlabGeometry(g.X_global, g.Y, g.Z, rank + 1, self.world_size, state.plan.halo, 0) \
                if rank < self.world_size - 1 else None
            peers = self._peers(rank)
            descriptor = state.descriptor(level)
            _lib.check(self.lib.lsf_slab_exchange(
                ctypes.byref(peers), ctypes.byref(descriptor), ctypes.c_size_t(offset), int(width),
                low.planes if low else 0, low.own_end if low else 0,
                high.planes if high else 0, high.own_begin - int(width) if high else 0,
                int(reduce_iteration), ctypes.c_uint(self.sequence), state.stream_handle()))

    def iterations(self, params, states, level, first_iteration, count):
        """enqueues `count` whole iterations (phases + exchanges) of every local rank with one library call per rank"""
        used = ctypes.c_uint(0)
        for state in states:
            rank = state.plan.rank
            g = state.plan.levels[level]
            low = SlabGeometry(g.X_global, g.Y, g.Z, rank - 1, self.world_size, state.plan.halo, 0) if rank > 0 else None
            high = SlabGeometry(g.X_global, g.Y, g.Z, rank + 1, self.world_size, state.plan.halo, 0) \
                if rank < self.world_size - 1 else None
            link = _lib.SlabLink()
            link.pre_offset = state.g_pre.data_ptr() - self.owned[rank]
            link.post_offset = state.g_post.data_ptr() - self.owned[rank]
            link.low_planes, link.low_own_end = (low.planes, low.own_end) if low else (0, 0)
            link.high_planes, link.high_own_begin = (high.planes, high.own_begin) if high else (0, 0)
            peers = self._peers(rank)
            descriptor = state.descriptor(level)
            _lib.check(self.lib.lsf_hier_slab_iterations(
                ctypes.byref(params), ctypes.byref(descriptor), ctypes.byref(peers), ctypes.byref(link),
                int(first_iteration), int(count), ctypes.c_uint(self.sequence + 1), ctypes.byref(used),
                state.stream_handle()))
        self.sequence += used.value

    def check(self, states):
        for state in states:
            error = ctypes.c_int(0)
            peers = self._peers(state.plan.rank)
            _lib.check(self.lib.lsf_slab_exchange_error(ctypes.byref(peers), ctypes.byref(error), state.stream_handle()))
            if error.value:
                raise RuntimeError("slab exchange: rank %d waited more than 4 s for its neighbours" % state.plan.rank)


# ------------------------------------------------------------------------------------------------ per-rank state
def _ptr(tensor):
    return ctypes.c_void_p(tensor.data_ptr())


class _RankState:
    """Device buffers of one (virtual) rank."""

    def __init__(self, plan, canonical_own, live_region, device, stream=None):
        import torch
        self.plan = plan
        self.device = device
        self.stream = stream  # torch.cuda.Stream of this (virtual) rank; None = the caller's current stream
        with self.on_stream():
            self._build(canonical_own, live_region)

    def on_stream(self):
        import contextlib
        import torch
        return torch.cuda.stream(self.stream) if self.stream is not None else contextlib.nullcontext()

    def stream_handle(self):
        if self.stream is not None:
            return ctypes.c_void_p(self.stream.cuda_stream)
        return _lib.current_stream_handle()

    def _build(self, canonical_own, live_region):
        import torch
        plan, device = self.plan, self.device
        lib = _lib.load()
        stream = self.stream_handle()
        finest = plan.levels[-1]
        live_region = live_region.to(device=device, dtype=torch.float32).contiguous()
        canonical_own = canonical_own.to(device=device, dtype=torch.float32).contiguous()
        if tuple(live_region.shape) != (plan.live_hi - plan.live_lo, finest.Y, finest.Z):
            raise ValueError("live region has shape %s, expected planes [%d, %d) x %d x %d"
                             % (tuple(live_region.shape), plan.live_lo, plan.live_hi, finest.Y, finest.Z))
        if tuple(canonical_own.shape) != (finest.own_hi - finest.own_lo, finest.Y, finest.Z):
            raise ValueError("canonical slab has shape %s, expected planes [%d, %d) x %d x %d"
                             % (tuple(canonical_own.shape), finest.own_lo, finest.own_hi, finest.Y, finest.Z))
        fp = lambda t: ctypes.cast(_ptr(t), _lib.c_float_p)
        self.packs, self.canonicals = [None] * plan.level_count, [None] * plan.level_count
        # finest level: pack from the live region, canonical into the allocation
        pack = torch.empty((finest.pack_padded_count, 4), dtype=torch.float32, device=device)
        _lib.check(lib.lsf_slab_pack_finest(fp(live_region), plan.live_hi - plan.live_lo, plan.live_lo,
                                            finest.X_global, finest.Y, finest.Z, _ptr(pack), finest.pack_planes,
                                            finest.pack_lo, stream))
        self.packs[-1] = pack
        canonical = torch.zeros((finest.planes, finest.Y, finest.Z), dtype=torch.float32, device=device)
        canonical[finest.own_begin:finest.own_end].copy_(canonical_own)
        self.canonicals[-1] = canonical
        # coarser levels: restrict (reference pyramid.tpp:51-74, downsampleX2_average)
        for level in range(plan.level_count - 2, -1, -1):
            g, f = plan.levels[level], plan.levels[level + 1]
            pack = torch.empty((g.pack_padded_count, 4), dtype=torch.float32, device=device)
            _lib.check(lib.lsf_slab_restrict(1, _ptr(self.packs[level + 1]), f.pack_planes, f.pack_lo, f.Y, f.Z,
                                             _ptr(pack), g.pack_planes, g.pack_lo, 0, g.pack_planes, stream))
            self.packs[level] = pack
            canonical = torch.zeros((g.planes, g.Y, g.Z), dtype=torch.float32, device=device)
            _lib.check(lib.lsf_slab_restrict(0, _ptr(self.canonicals[level + 1]), f.planes, f.x_origin, f.Y, f.Z,
                                             _ptr(canonical), g.planes, g.x_origin, g.own_begin, g.own_end, stream))
            self.canonicals[level] = canonical
        self.violation = torch.zeros(1, dtype=torch.int32, device=device)
        self.warp = None
        self.g_post = self.g_pre = self.slots = None

    def start_level(self, level, max_iterations, exchange=None):
        import torch
        g = self.plan.levels[level]
        shape = (3, g.planes, g.Y, g.Z)
        if level == 0:
            self.warp = torch.zeros(shape, dtype=torch.float32, device=self.device)
        if exchange is not None and getattr(exchange, "fused", False):
            # views of the peer-visible allocation; g_pre is NOT cleared (see PeerExchange)
            self.g_pre, self.g_post = exchange.fields(self.plan.rank, g)
            self.g_post.zero_()
        else:
            self.g_post = torch.zeros(shape, dtype=torch.float32, device=self.device)  # optimizer.tpp:142-143
            self.g_pre = torch.zeros(shape, dtype=torch.float32, device=self.device)
        self.slots = torch.zeros(max(max_iterations, 1), dtype=torch.int32, device=self.device)

    def descriptor(self, level):
        g = self.plan.levels[level]
        d = _lib.SlabLevel()
        d.planes, d.Y, d.Z = g.planes, g.Y, g.Z
        d.own_begin, d.own_end = g.own_begin, g.own_end
        d.x_origin, d.X_global = g.x_origin, g.X_global
        d.pack = self.packs[level].data_ptr()
        d.pack_planes, d.pack_origin = g.pack_planes, g.pack_lo
        d.pack_interior_low, d.pack_interior_high = g.pack_interior_low, g.pack_interior_high
        d.canonical = self.canonicals[level].data_ptr()
        d.warp = self.warp.data_ptr()
        d.g_post = self.g_post.data_ptr()
        d.g_pre = self.g_pre.data_ptr()
        d.max_sq_bits = self.slots.data_ptr()
        d.violation = self.violation.data_ptr()
        return d

    def prolong(self, level):
        """reference optimizer.tpp:124-126 (values are NOT doubled)"""
        import torch
        lib = _lib.load()
        g, f = self.plan.levels[level], self.plan.levels[level + 1]
        fine = torch.zeros((3, f.planes, f.Y, f.Z), dtype=torch.float32, device=self.device)
        fp = lambda t: ctypes.cast(_ptr(t), _lib.c_float_p)
        _lib.check(lib.lsf_slab_prolong_nearest(fp(self.warp), g.planes, g.x_origin, g.Y, g.Z, fp(fine), f.planes,
                                                f.x_origin, f.own_begin, f.own_end, self.stream_handle()))
        self.warp = fine


class SlabHierarchicalOptimizer3d:
    """HierarchicalOptimizer3d over slabs. `optimizer` is a lsf_b200.HierarchicalOptimizer3d carrying the parameters."""

    def __init__(self, optimizer, pack_halo=32, exchange=None):
        """exchange: "peer" (kernels storing into the neighbours' memory; default with NCCL ranks of one box, at most 8),
        "dist" (torch.distributed send / recv + all_reduce); the environment variable LSF_SLAB_EXCHANGE sets the default."""
        if int(optimizer.resampling_strategy) != 0:
            raise RuntimeError("the slab decomposition supports the NEAREST_AND_AVERAGE resampling strategy only")
        self.optimizer = optimizer
        self.pack_halo = int(pack_halo)
        self.exchange_kind = exchange or os.environ.get("LSF_SLAB_EXCHANGE", "peer")
        if self.exchange_kind not in ("peer", "dist"):
            raise ValueError("exchange must be 'peer' or 'dist', got %r" % (self.exchange_kind,))
        self._peer_exchange = None
        self.iteration_counts = []
        self.max_update_lengths = []
        self.exchanged_bytes = 0

    def _flags(self):
        o = self.optimizer
        tikhonov = o.tikhonov_term_enabled and o.tikhonov_strength > 0
        use_kernel = o.gradient_kernel_enabled and o.kernel is not None and o.kernel.size > 0
        radius = int(o.kernel.size) // 2 if use_kernel else 0
        return tikhonov, use_kernel, radius

    def plan(self, shape, rank, world_size):
        tikhonov, _, radius = self._flags()
        return SlabPlan(shape, rank, world_size, self.optimizer.maximum_chunk_size, radius, tikhonov, self.pack_halo)

    def _run(self, states, exchange):
        """Level / iteration loop over the rank states of this process (reference optimizer.tpp:112-171)."""
        import torch
        lib = _lib.load()
        o = self.optimizer
        params = o._params()
        tikhonov, use_kernel, radius = self._flags()
        fused = getattr(exchange, "fused", False)
        plan0 = states[0].plan
        if fused:
            finest = plan0.levels[-1]
            per_rank = finest.own_hi - finest.own_lo
            exchange.setup(3 * (per_rank + 2 * plan0.halo) * finest.Y * finest.Z)
        self.iteration_counts, self.max_update_lengths = [], []
        self.exchanged_bytes = 0

        def phase(number, it, level):
            for s in states:
                d = s.descriptor(level)
                _lib.check(lib.lsf_hier_slab_iteration(ctypes.byref(params), ctypes.byref(d), it, number, s.stream_handle()))

        for level in range(plan0.level_count):
            geometries = [s.plan.levels[level] for s in states]
            for s in states:
                with s.on_stream():
                    s.start_level(level, o.maximum_iteration_count, exchange)
            executed, enqueued, converged, last_max = 0, 0, False, float("inf")
            plane_bytes = geometries[0].Y * geometries[0].Z * 3 * 4
            per_iteration_bytes = (2 * radius * plane_bytes if use_kernel else 0) + (2 * plane_bytes if tikhonov else 0)
            if fused:
                # one library call per chunk and rank; one chunk is always in flight behind the one the host waits for
                # (iterations past the convergence point are no-ops on the device: the kernels test the previous slot)
                in_flight = []
                while True:
                    if not converged and enqueued < o.maximum_iteration_count:
                        chunk_end = min(o.maximum_iteration_count, enqueued + POLL_CHUNK)
                        exchange.iterations(params, states, level, enqueued, chunk_end - enqueued)
                        if not use_kernel and (chunk_end - enqueued) % 2:
                            for s in states:  # the library swapped the roles once per iteration
                                s.g_pre, s.g_post = s.g_post, s.g_pre
                        with states[0].on_stream():
                            host_bits = torch.empty(chunk_end - enqueued, dtype=torch.int32, pin_memory=True)
                            host_bits.copy_(states[0].slots[enqueued:chunk_end], non_blocking=True)
                            done = torch.cuda.Event()
                            done.record()
                        in_flight.append((enqueued, chunk_end, host_bits, done))
                        self.exchanged_bytes += (chunk_end - enqueued) * per_iteration_bytes
                        enqueued = chunk_end
                        if len(in_flight) < 2 and enqueued < o.maximum_iteration_count:
                            continue
                    if not in_flight:
                        break
                    begin, end, host_bits, done = in_flight.pop(0)
                    done.synchronize()
                    if converged:
                        continue
                    for it, value in zip(range(begin, end), host_bits.numpy().view(np.float32)):
                        last_max = float(np.sqrt(value))
                        executed = it + 1
                        if last_max < o.maximum_warp_update_threshold:  # optimizer.tpp:166-171
                            converged = True
                            break
                exchange.check(states)
            while not fused and not converged and enqueued < o.maximum_iteration_count:
                chunk_end = min(o.maximum_iteration_count, enqueued + POLL_CHUNK)
                for it in range(enqueued, chunk_end):
                    phase(1, it, level)
                    if use_kernel:
                        exchange.halos([s.g_pre for s in states], geometries, radius)
                        phase(2, it, level)
                    else:
                        for s in states:  # without a filter phase 1 wrote the final gradient into g_pre
                            s.g_pre, s.g_post = s.g_post, s.g_pre
                    if tikhonov:
                        exchange.halos([s.g_post for s in states], geometries, 1)
                    exchange.reduce_max([s.slots for s in states], it)
                    self.exchanged_bytes += per_iteration_bytes
                with states[0].on_stream():
                    bits = states[0].slots[enqueued:chunk_end].cpu().numpy()  # synchronises once per chunk
                for it, value in zip(range(enqueued, chunk_end), bits.view(np.float32)):
                    last_max = float(np.sqrt(value))
                    executed = it + 1
                    if last_max < o.maximum_warp_update_threshold:  # optimizer.tpp:166-171
                        converged = True
                        break
                enqueued = chunk_end
            # iterations enqueued behind the convergence point are no-ops on the device (at most two polling chunks); their
            # exchange rounds still run but move nothing the result depends on: counted as executed iterations only
            self.exchanged_bytes -= (enqueued - executed) * per_iteration_bytes
            self.iteration_counts.append(executed)
            self.max_update_lengths.append(last_max)
            if level != plan0.level_count - 1:
                for s in states:
                    with s.on_stream():
                        s.prolong(level)
        results = []
        for s in states:
            with s.on_stream():
                if int(s.violation.item()):
                    raise RuntimeError("a warp vector reached beyond the rank's gather halo (%d planes at the finest "
                                       "level): increase pack_halo" % self.pack_halo)
                g = s.plan.levels[-1]
                results.append(s.warp[:, g.own_begin:g.own_end].permute(1, 2, 3, 0).contiguous())
                if s.stream is not None:
                    s.stream.synchronize()
        return results

    # ---------------------------------------------------------------- one process per GPU
    def optimize(self, canonical_slab, live_region, shape, group=None):
        """This rank's part of optimize(canonical_field, live_field): `canonical_slab` = planes plan.own_range() of the
        canonical field, `live_region` = planes plan.live_range() of the live field, `shape` = the whole volume.
        Returns this rank's slab of the warp field, [planes, Y, Z, 3] on the GPU."""
        import torch
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            world_size = dist.get_world_size(group)
            if (self.exchange_kind == "peer" and 1 < world_size <= _lib.SLAB_MAX_PEERS
                    and dist.get_backend(group) == "nccl"):
                if self._peer_exchange is None:
                    self._peer_exchange = PeerExchange(group)
                exchange = self._peer_exchange
            else:
                exchange = DistExchange(group)
            rank, world_size = exchange.rank, exchange.world_size
        else:
            exchange, rank, world_size = LocalExchange(), 0, 1
        plan = self.plan(shape, rank, world_size)
        device = torch.device("cuda", torch.cuda.current_device())
        state = _RankState(plan, torch.as_tensor(canonical_slab), torch.as_tensor(live_region), device)
        if getattr(exchange, "fused", False):
            dist.barrier(group=group)  # the ranks wait for each other on the device: start them together
        return self._run([state], exchange)[0]

    # ---------------------------------------------------------------- all ranks emulated in this process
    def optimize_emulated(self, canonical_field, live_field, world_size, exchange="local"):
        """Runs `world_size` virtual ranks on the current GPU and returns the assembled warp field [X, Y, Z, 3] (numpy).
        Verifies the decomposition without a multi-GPU box. exchange="local": one stream, halos copied between the ranks'
        tensors; "peer": one stream per rank and the peer-memory exchange kernel (PeerExchange) -- the ranks wait for each
        other on the device exactly like the GPUs of a box do. (The virtual ranks need their kernels to run concurrently on
        the one GPU: under tools that serialise kernels, e.g. compute-sanitizer, the waits time out and the call raises.)"""
        import torch
        canonical = torch.as_tensor(np.ascontiguousarray(canonical_field, dtype=np.float32))
        live = torch.as_tensor(np.ascontiguousarray(live_field, dtype=np.float32))
        device = torch.device("cuda", torch.cuda.current_device())
        if exchange not in ("local", "peer"):
            raise ValueError("exchange must be 'local' or 'peer'")
        peer = exchange == "peer" and world_size > 1
        states = []
        torch.cuda.synchronize()
        for rank in range(world_size):
            plan = self.plan(tuple(canonical.shape), rank, world_size)
            own_lo, own_hi = plan.own_range()
            live_lo, live_hi = plan.live_range()
            states.append(_RankState(plan, canonical[own_lo:own_hi], live[live_lo:live_hi], device,
                                     torch.cuda.Stream(device) if peer else None))
        if peer:
            link = PeerExchange(virtual_ranks=world_size)
            try:
                slabs = self._run(states, link)
            finally:
                link.close()
        else:
            slabs = self._run(states, LocalExchange())
        return torch.cat(slabs, dim=0).cpu().numpy()

    def close(self):
        """frees the peer-visible allocation (collective over the ranks)"""
        if self._peer_exchange is not None:
            self._peer_exchange.close()
            self._peer_exchange = None
