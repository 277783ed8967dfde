"""GPU test of the multipair driver (BASELINE.json configs[3]): several pairs in flight per GPU (one CUDA
stream per worker thread) give exactly the results of the serial loop."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


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


def test_run_multipair_matches_serial_reports(lsf, tmp_path):
    """the experiment driver (pair cache -> optimize over streams -> report table) against a plain serial loop"""
    pytest.importorskip("pandas")
    from lsf_b200 import multipair, synthetic
    rng = np.random.default_rng(11)
    data_path, out_path = str(tmp_path / "data"), str(tmp_path / "out")
    for frame in (3, 12, 7):
        shift = tuple(np.array([2.5, -1.5, 1.0]) + rng.uniform(-2, 2, 3))
        canonical, live = synthetic.sphere_plane_pair_3d(32, shift=shift)
        multipair.save_pair(data_path, frame, 214, canonical, live)
    kwargs = dict(tikhonov_term_enabled=False, gradient_kernel_enabled=True, kernel=synthetic.sobolev_kernel_1d(),
                  maximum_chunk_size=4, maximum_iteration_count=12, maximum_warp_update_threshold=0.02,
                  logging_parameters=lsf.HierarchicalOptimizer3d.LoggingParameters(
                      collect_per_level_convergence_reports=True))
    factory = lambda: lsf.HierarchicalOptimizer3d(**kwargs)
    table = multipair.run_multipair(data_path, out_path, factory, streams=2, save_warps=True)
    entries = multipair.list_pair_cache(data_path)
    assert list(table["canonical_frame"]) == [frame for frame, _, _ in entries] == [12, 3, 7]
    optimizer = factory()
    report_sets = []
    for frame, row, path in entries:
        canonical, live = multipair.load_pair(path)
        warp = optimizer.optimize(canonical, live)
        report_sets.append(optimizer.get_per_level_convergence_reports())
        assert np.array_equal(np.load(os.path.join(out_path, "warp_%d_%d.npy" % (frame, row))), warp)
    expected = multipair.post_process_convergence_report_sets(report_sets, [(f, r) for f, r, _ in entries])
    assert table.equals(expected)
    assert os.path.exists(os.path.join(out_path, "convergence_reports.pk"))
    assert os.path.exists(os.path.join(out_path, "analysis.txt"))
    # reference --save_telemetry: per-level iteration data of every pair in telemetry/pair_<f>-<f+1>_<row>/telemetry_log.npz
    kwargs["logging_parameters"] = lsf.HierarchicalOptimizer3d.LoggingParameters(
        collect_per_level_convergence_reports=True, collect_per_level_iteration_data=True)
    telemetry_table = multipair.run_multipair(data_path, out_path, factory, streams=1, save_telemetry=True)
    assert telemetry_table.equals(expected)  # collecting the iteration data does not change the result
    optimizer = factory()
    for frame, row, path in entries:
        canonical, live = multipair.load_pair(path)
        optimizer.optimize(canonical, live)
        log = optimizer.get_per_level_iteration_data()
        stored = np.load(os.path.join(multipair.get_telemetry_subfolder_path(os.path.join(out_path, "telemetry"), frame, row),
                                      "telemetry_log.npz"))
        assert len(stored.files) == 3 * len(log)
        for level, level_data in enumerate(log):
            assert np.array_equal(stored["l%d_warp_fields" % level], np.dstack(level_data.get_warp_fields()))
            assert np.array_equal(stored["l%d_data_term_gradients" % level], np.dstack(level_data.get_data_term_gradients()))
            assert stored["l%d_tikhonov_term_gradients" % level].size == 0  # the Tikhonov term is off in this run


@pytest.mark.parametrize("mode", ["tikhonov_kernel", "kernel", "tikhonov", "data_only"])
def test_batched_optimize_matches_pair_by_pair(lsf, mode, monkeypatch):
    """lsf_hier_optimize_3d_batch: the pairs of a batch advance through the pyramid together (every iteration kernel
    covers all pairs; per-pair convergence slots and termination). Every pair's warp field and iteration counts must be
    those of its own optimize() call bit for bit -- with pairs that terminate after different numbers of iterations
    (odd and even: the deferred update's ping-pong buffers), from host arrays and from device tensors, and for the
    serial fallback (LSF_BATCH=0)."""
    import torch
    from lsf_b200 import synthetic
    modes = {
        "data_only": dict(tikhonov_term_enabled=False, gradient_kernel_enabled=False),
        "tikhonov": dict(tikhonov_term_enabled=True, gradient_kernel_enabled=False, tikhonov_strength=0.05),
        "kernel": dict(tikhonov_term_enabled=False, gradient_kernel_enabled=True),
        "tikhonov_kernel": dict(tikhonov_term_enabled=True, gradient_kernel_enabled=True, tikhonov_strength=0.1),
    }
    shifts = [(2.5, -1.5, 1.0), (0.4, 0.2, -0.3), (1.5, 0.5, 0.7), (-2.0, 1.0, 0.5), (0.1, 0.05, 0.02)]
    pairs = [synthetic.sphere_plane_pair_3d(48, shift=shift) for shift in shifts]
    canonical = np.stack([p[0][:, :40, :].copy() for p in pairs])   # 48 x 40 x 48: ragged tiles
    live = np.stack([p[1][:, :40, :].copy() for p in pairs])
    threshold = 0.02 if "kernel" in mode else 0.016  # levels of different pairs end after 1 .. 37 iterations
    kwargs = dict(modes[mode], maximum_chunk_size=4, maximum_iteration_count=37, maximum_warp_update_threshold=threshold,
                  kernel=synthetic.sobolev_kernel_1d())
    optimizer = lsf.HierarchicalOptimizer3d(**kwargs)
    expected, expected_counts = [], []
    for index in range(len(pairs)):
        expected.append(optimizer.optimize(canonical[index], live[index]))
        expected_counts.append(optimizer.get_per_level_iteration_counts())
    assert len({tuple(c) for c in expected_counts}) > 1, expected_counts  # the pairs really stop at different iterations
    batch = optimizer.optimize_batch(canonical, live)
    assert optimizer.get_per_pair_iteration_counts() == expected_counts
    for index in range(len(pairs)):
        assert np.array_equal(batch[index], expected[index]), index
    device_batch = optimizer.optimize_batch(torch.from_numpy(canonical).cuda(), torch.from_numpy(live).cuda())
    assert device_batch.is_cuda and np.array_equal(device_batch.cpu().numpy(), batch)
    monkeypatch.setenv("LSF_BATCH", "0")
    serial = optimizer.optimize_batch(canonical, live)
    assert np.array_equal(serial, batch) and optimizer.get_per_pair_iteration_counts() == expected_counts


def test_numpy_calls_run_on_the_callers_current_stream():
    """multigpu.optimize_pairs / multipair.run_multipair give every worker thread its own torch stream and feed numpy
    pairs: the LSF_HOST path must enqueue on that stream (it used to hard-code the legacy default stream, which serialised
    the workers). Stream identity is checked directly, and a numpy call inside a side stream must leave results equal to
    the default-stream call."""
    import torch
    from lsf_b200 import _lib, synthetic
    import lsf_b200
    assert _lib.host_stream_handle().value in (None, 0) or _lib.host_stream_handle().value == torch.cuda.current_stream().cuda_stream
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        assert _lib.host_stream_handle().value == side.cuda_stream
        assert _lib.current_stream_handle().value == side.cuda_stream
        canonical, live = synthetic.sphere_plane_pair_3d(32)
        optimizer = lsf_b200.HierarchicalOptimizer3d(maximum_chunk_size=4, maximum_iteration_count=10)
        on_side = optimizer.optimize(canonical, live)
    on_default = lsf_b200.HierarchicalOptimizer3d(maximum_chunk_size=4, maximum_iteration_count=10).optimize(canonical, live)
    assert np.array_equal(on_side, on_default)
