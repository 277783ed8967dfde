"""GPU test of the multipair driver (BASELINE.json configs[3]): several pairs in flight per GPU (one CUDA
stream per worker thread) give exactly the results of the serial loop."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lsf():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import lsf_b200
    return lsf_b200


def test_streams_match_serial_loop(lsf):
    import torch
    from lsf_b200 import multigpu, synthetic
    rng = np.random.default_rng(7)
    pairs = []
    for _ in range(6):
        shift = tuple(np.array([2.5, -1.5, 1.0]) + rng.uniform(-2, 2, 3))
        pairs.append(synthetic.sphere_plane_pair_3d(32, shift=shift, xp=torch, device="cuda"))
    kwargs = dict(tikhonov_term_enabled=True, tikhonov_strength=0.1, gradient_kernel_enabled=True,
                  kernel=synthetic.sobolev_kernel_1d(), maximum_chunk_size=4, maximum_iteration_count=15,
                  maximum_warp_update_threshold=0.02)

    def call(optimizer, canonical, live):
        warp = optimizer.optimize(canonical, live)
        return warp.cpu().numpy(), optimizer.get_per_level_iteration_counts()

    serial = multigpu.optimize_pairs(multigpu.PerWorkerOptimizer(lambda: lsf.HierarchicalOptimizer3d(**kwargs), call),
                                     len(pairs), lambda i: pairs[i], rank=0, world_size=1, streams=1)
    threaded = multigpu.optimize_pairs(multigpu.PerWorkerOptimizer(lambda: lsf.HierarchicalOptimizer3d(**kwargs), call),
                                       len(pairs), lambda i: pairs[i], rank=0, world_size=1, streams=3)
    assert len(serial) == len(threaded) == len(pairs)
    for (warp_a, counts_a), (warp_b, counts_b) in zip(serial, threaded):
        assert counts_a == counts_b
        assert np.array_equal(warp_a, warp_b)
    assert any(np.abs(w).max() > 0 for w, _ in serial)
    # an exception in a worker reaches the caller
    def failing(optimizer, canonical, live):
        raise RuntimeError("boom")
    with pytest.raises(RuntimeError, match="boom"):
        multigpu.optimize_pairs(multigpu.PerWorkerOptimizer(lambda: lsf.HierarchicalOptimizer3d(**kwargs), failing),
                                len(pairs), lambda i: pairs[i], rank=0, world_size=1, streams=2)
