"""Slab decomposition of one volume on a single GPU: `world_size` virtual ranks run in lockstep and
exchange halos by device copies (slab.LocalExchange); the assembled warp field must be BIT-IDENTICAL to the
whole-volume optimizer's and the per-level iteration counts equal. The multi-process NCCL path uses the same kernels
and the same exchange points (tools/slab_multigpu_check.py runs it on 2+ GPUs)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

MODES = {
    "tikhonov_kernel": dict(tikhonov_term_enabled=True, gradient_kernel_enabled=True, tikhonov_strength=0.1),
    "kernel": dict(tikhonov_term_enabled=False, gradient_kernel_enabled=True),
    "tikhonov": dict(tikhonov_term_enabled=True, gradient_kernel_enabled=False, tikhonov_strength=0.05),
    "data_only": dict(tikhonov_term_enabled=False, gradient_kernel_enabled=False),
}


def make_optimizer(lsf, mode, iterations=12, threshold=0.01):
    from lsf_b200 import synthetic
    return lsf.HierarchicalOptimizer3d(maximum_chunk_size=4, maximum_iteration_count=iterations,
                                       maximum_warp_update_threshold=threshold, kernel=synthetic.sobolev_kernel_1d(),
                                       **MODES[mode])


@pytest.mark.parametrize("mode", sorted(MODES))
@pytest.mark.parametrize("world_size", [2, 4])
@pytest.mark.parametrize("slab_kernels", ["fast", "first_generation"])
def test_slabs_match_whole_volume(lsf, mode, world_size, slab_kernels, monkeypatch):
    """slab_kernels: the TMA-fed stage 1 + marching filter kernels of slab mode (default) and the first-generation
    slab kernels (LSF_SLAB_FAST=0) -- both bit-identical to the whole-volume optimizer"""
    from lsf_b200 import slab, synthetic
    if slab_kernels == "first_generation":
        monkeypatch.setenv("LSF_SLAB_FAST", "0")
    canonical, live = synthetic.sphere_plane_pair_3d(64)
    canonical, live = canonical[:, :48, :56].copy(), live[:, :48, :56].copy()
    optimizer = make_optimizer(lsf, mode)
    whole = optimizer.optimize(canonical, live)
    sharded = slab.SlabHierarchicalOptimizer3d(optimizer, pack_halo=16)
    warp = sharded.optimize_emulated(canonical, live, world_size)
    assert sharded.iteration_counts == optimizer.get_per_level_iteration_counts()
    assert np.array_equal(warp, whole)
    assert np.allclose(sharded.max_update_lengths,
                       [r.max_update_length for r in optimizer.get_per_level_convergence_reports()], rtol=0, atol=0)


@pytest.mark.parametrize("mode", sorted(MODES))
@pytest.mark.parametrize("world_size", [2, 4])
def test_peer_memory_exchange_matches_whole_volume(lsf, mode, world_size):
    """the peer-memory exchange kernel (csrc/slab_peer.cu: boundary planes stored into the neighbours' halo planes, mailbox
    signals, termination maximum to all ranks, device-side wait) with one stream per virtual rank -- the ranks wait for
    each other on the device like the GPUs of a box; bit-identical to the whole-volume optimizer in all term modes"""
    from lsf_b200 import slab, synthetic
    canonical, live = synthetic.sphere_plane_pair_3d(64)
    canonical, live = canonical[:, :48, :56].copy(), live[:, :48, :56].copy()
    optimizer = make_optimizer(lsf, mode)
    whole = optimizer.optimize(canonical, live)
    sharded = slab.SlabHierarchicalOptimizer3d(optimizer, pack_halo=16)
    warp = sharded.optimize_emulated(canonical, live, world_size, exchange="peer")
    assert sharded.iteration_counts == optimizer.get_per_level_iteration_counts()
    assert np.array_equal(warp, whole)
    assert np.allclose(sharded.max_update_lengths,
                       [r.max_update_length for r in optimizer.get_per_level_convergence_reports()], rtol=0, atol=0)


def test_peer_memory_exchange_early_termination_and_odd_plane_size(lsf):
    """levels that converge inside a polling chunk (the remaining iterations of the chunk are no-ops on every rank, the
    exchanges still pair up) and planes whose float count is not a multiple of 4 (scalar copy path of the exchange kernel
    at the coarsest level: 13 x 11 voxels per plane)"""
    from lsf_b200 import slab, synthetic
    canonical, live = synthetic.sphere_plane_pair_3d(64)
    optimizer = make_optimizer(lsf, "kernel", iterations=40, threshold=0.05)
    whole = optimizer.optimize(canonical, live)
    counts = optimizer.get_per_level_iteration_counts()
    assert any(c < 40 for c in counts)
    sharded = slab.SlabHierarchicalOptimizer3d(optimizer)
    assert np.array_equal(sharded.optimize_emulated(canonical, live, 2, exchange="peer"), whole)
    assert sharded.iteration_counts == counts
    canonical, live = canonical[:, :52, :44].copy(), live[:, :52, :44].copy()
    optimizer = make_optimizer(lsf, "tikhonov_kernel", iterations=6)
    whole = optimizer.optimize(canonical, live)
    sharded = slab.SlabHierarchicalOptimizer3d(optimizer, pack_halo=16)
    assert np.array_equal(sharded.optimize_emulated(canonical, live, 4, exchange="peer"), whole)


def test_slabs_early_termination_and_single_rank(lsf):
    from lsf_b200 import slab, synthetic
    canonical, live = synthetic.sphere_plane_pair_3d(64)
    optimizer = make_optimizer(lsf, "kernel", iterations=40, threshold=0.05)
    whole = optimizer.optimize(canonical, live)
    counts = optimizer.get_per_level_iteration_counts()
    assert any(c < 40 for c in counts)
    for world_size in (1, 2):
        sharded = slab.SlabHierarchicalOptimizer3d(optimizer)
        assert np.array_equal(sharded.optimize_emulated(canonical, live, world_size), whole)
        assert sharded.iteration_counts == counts


def test_gather_halo_violation_is_reported(lsf):
    """a warp that reaches beyond the rank's static gather halo must raise the flag, never silently read padding"""
    import ctypes
    import torch
    from lsf_b200 import slab, synthetic, _lib
    canonical, live = synthetic.sphere_plane_pair_3d(64)
    optimizer = make_optimizer(lsf, "kernel")
    sharded = slab.SlabHierarchicalOptimizer3d(optimizer, pack_halo=2)
    device = torch.device("cuda", torch.cuda.current_device())
    flags = []
    for displacement in (0.5, 9.0):
        plan = sharded.plan(canonical.shape, 1, 4)  # an interior rank: both pack edges are cuts
        own_lo, own_hi = plan.own_range()
        live_lo, live_hi = plan.live_range()
        state = slab._RankState(plan, torch.from_numpy(canonical[own_lo:own_hi]), torch.from_numpy(live[live_lo:live_hi]),
                                device)
        level = plan.level_count - 1
        state.start_level(0, 4)
        state.warp = torch.zeros((3, plan.levels[level].planes, 64, 64), dtype=torch.float32, device=device)
        state.warp[0] = displacement
        state.start_level(level, 4)
        descriptor = state.descriptor(level)
        params = optimizer._params()
        _lib.check(_lib.load().lsf_hier_slab_iteration(ctypes.byref(params), ctypes.byref(descriptor), 0, 1,
                                                       _lib.current_stream_handle()))
        flags.append(int(state.violation.item()))
    assert flags == [0, 1]  # finest-level gather halo is 8 planes here: 0.5 stays inside, 9.0 does not


def test_slab_restrictions(lsf):
    from lsf_b200 import slab
    linear = lsf.HierarchicalOptimizer3d(resampling_strategy=lsf.HierarchicalOptimizer3d.ResamplingStrategy.LINEAR)
    with pytest.raises(RuntimeError):
        slab.SlabHierarchicalOptimizer3d(linear)
