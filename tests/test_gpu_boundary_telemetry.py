"""The drop-in boundary's telemetry (`-m gpu`): LoggingParameters.collect_per_level_iteration_data /
get_per_level_iteration_data() with OptimizationIterationData{2d,3d}, and the numbers behind VerbosityParameters'
per-iteration prints (reference cpp/src/nonrigid_optimization/hierarchical/optimizer_with_telemetry.tpp:83-182,
cpp/src/python_export/telemetry.tpp:136-145, hierarchical_optimizer.tpp:77-78).

`test_reference_test_cpp_iteration_data` and `test_reference_test_construction_and_operation01` repeat the calls and
assertions of the reference's own tests/test_hierarchical_optimizer2d.py:39-101 against the shim module
`level_set_fusion_optimization` (same constructor keywords, same accessors, same tolerances), with the reference's
fixtures from tests/golden/reference_literals.npz. The reference's test FILE cannot run on the GPU box (it is not part of
this repository and imports matplotlib / sktensor-dependent modules); the sequence of calls is the same.
"""
import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu


@pytest.fixture()
def ho_cpp(lsf):
    import level_set_fusion_optimization
    return level_set_fusion_optimization


def test_reference_test_construction_and_operation01(lsf, ho_cpp, literals):
    """reference tests/test_hierarchical_optimizer2d.py:39-69 (the C++ half)"""
    test_data = lambda n: literals["py_hierarchical/" + n]
    optimizer = ho_cpp.HierarchicalOptimizer2d(
        tikhonov_term_enabled=False,
        gradient_kernel_enabled=False,
        maximum_chunk_size=8,
        rate=0.2,
        maximum_iteration_count=100,
        maximum_warp_update_threshold=0.001,
        data_term_amplifier=1.0
    )
    warp_field_out = optimizer.optimize(test_data("canonical_field"), test_data("live_field"))
    final_warped_live = lsf.ops.warp(test_data("live_field"), warp_field_out)
    assert np.allclose(warp_field_out, test_data("warp_field"), atol=10e-6)
    assert np.allclose(final_warped_live, test_data("final_live_field"), atol=10e-6)


def test_reference_test_cpp_iteration_data(lsf, ho_cpp, literals):
    """reference tests/test_hierarchical_optimizer2d.py:71-101"""
    from lsf_b200 import synthetic
    test_data = lambda n: literals["py_hierarchical/" + n]
    optimizer = ho_cpp.HierarchicalOptimizer2d(
        tikhonov_term_enabled=False,
        gradient_kernel_enabled=False,

        maximum_chunk_size=8,
        rate=0.2,
        maximum_iteration_count=100,
        maximum_warp_update_threshold=0.001,

        data_term_amplifier=1.0,
        tikhonov_strength=0.0,

        kernel=synthetic.sobolev_kernel_1d(size=7, strength=0.1),  # reference: sob.generate_1d_sobolev_kernel (F15)

        resampling_strategy=ho_cpp.HierarchicalOptimizer2d.ResamplingStrategy.NEAREST_AND_AVERAGE,

        verbosity_parameters=ho_cpp.HierarchicalOptimizer2d.VerbosityParameters(),
        logging_parameters=ho_cpp.HierarchicalOptimizer2d.LoggingParameters(
            collect_per_level_convergence_reports=True,
            collect_per_level_iteration_data=True
        )
    )
    warp_field_out = optimizer.optimize(test_data("canonical_field"), test_data("live_field"))
    final_warped_live = lsf.ops.warp(test_data("live_field"), warp_field_out)
    data = optimizer.get_per_level_iteration_data()
    vec = data[3].get_warp_fields()

    assert np.allclose(vec[50], test_data("iteration50_warp_field"), atol=1e-6)

    assert np.allclose(warp_field_out, test_data("warp_field"), atol=10e-6)
    assert np.allclose(final_warped_live, test_data("final_live_field"), atol=10e-6)
    # structure of the collected data (optimizer_with_telemetry.tpp:90-99,153-159)
    assert len(data) == 4 and isinstance(data[0], ho_cpp.OptimizationIterationData2d)
    counts = [r.iteration_count for r in optimizer.get_per_level_convergence_reports()]
    assert [d.get_frame_count() for d in data] == [counts[0] + 1] + counts[1:]
    assert data[0].get_warp_fields()[0].shape == (2, 2, 2) and not data[0].get_warp_fields()[0].any()
    assert data[3].get_live_fields()[0].shape == (16, 16)
    assert data[3].get_tikhonov_term_gradients()[0].size == 0  # Tikhonov term off: empty containers


@pytest.mark.parametrize("nd", [2, 3])
def test_iteration_data_and_statistics_vs_oracle(lsf, nd):
    """Fields and statistics of every iteration against the oracle's primitives: data-term gradient = resampled live
    gradient * (resampled live - canonical) (optimizer.tpp:186-194), Tikhonov-term gradient = Laplacian of the previous
    gradient (tpp:195-196), warp after the iteration = the oracle's per-iteration dump; the optimizer's result is the
    same with and without telemetry."""
    from lsf_b200 import synthetic
    if nd == 3:
        canonical, live = synthetic.sphere_plane_pair_3d(32)
        cls = lsf.HierarchicalOptimizer3d
    else:
        canonical, live = synthetic.circle_line_pair_2d(64)
        cls = lsf.HierarchicalOptimizer2d
    kwargs = dict(tikhonov_term_enabled=True, tikhonov_strength=0.05, gradient_kernel_enabled=False, maximum_chunk_size=1,
                  rate=0.1, maximum_iteration_count=4, maximum_warp_update_threshold=0.001, data_term_amplifier=1.0)
    plain = cls(**kwargs)
    expected_warp = plain.optimize(canonical, live)
    optimizer = cls(logging_parameters=cls.LoggingParameters(collect_per_level_iteration_data=True), **kwargs)
    warp = optimizer.optimize(canonical, live)
    assert np.array_equal(warp, expected_warp)
    assert optimizer.get_per_level_iteration_counts() == plain.get_per_level_iteration_counts() == [4]
    expected = oracle.hier_optimize(canonical, live, dump_level=0, dump_iterations=4, **kwargs)
    data = optimizer.get_per_level_iteration_data()
    assert len(data) == 1 and data[0].get_frame_count() == 5  # level 0: initial frame + 4 iterations
    warps, data_gradients = data[0].get_warp_fields(), data[0].get_data_term_gradients()
    tikhonov_gradients, lives = data[0].get_tikhonov_term_gradients(), data[0].get_live_fields()
    live_gradient = oracle.gradient(live)
    previous_gradient = np.zeros(live.shape + (nd,), np.float32)
    statistics = optimizer.get_per_iteration_statistics()
    assert len(statistics) == 4
    for it in range(4):
        warp_before = warps[it]  # frame 0 is the initial (zero) frame
        assert np.array_equal(warps[it + 1], expected["dump"][it])
        assert np.array_equal(lives[it + 1], live)
        diff = oracle.warp(live, warp_before) - canonical
        data_gradient = oracle.warp_with_replacement(live_gradient, warp_before, 0.0) * diff[..., None]
        assert np.array_equal(data_gradients[it + 1], data_gradient)
        tikhonov_gradient = oracle.laplacian(previous_gradient)
        assert np.array_equal(tikhonov_gradients[it + 1], tikhonov_gradient)
        # statistics of the prints: mean / std of diff, normalised energies (optimizer_with_telemetry.tpp:139-181)
        level, iteration, max_update, mean_diff, std_diff, data_energy, tikhonov_energy = statistics[it]
        assert (level, iteration) == (0, it)
        d64 = diff.astype(np.float64)
        assert abs(mean_diff - d64.mean()) <= 1e-6 + 1e-5 * abs(d64.mean())
        assert abs(std_diff - d64.std()) <= 1e-5 * d64.std()
        assert abs(data_energy - 1e6 * (d64 ** 2).mean()) <= 1e-4 * 1e6 * (d64 ** 2).mean()
        jacobian_sum = sum(np.gradient(previous_gradient[..., c].astype(np.float64), axis=a)
                           for c in range(nd) for a in range(nd))
        expected_tikhonov_energy = 1e6 * 0.5 * (jacobian_sum ** 2).mean()
        assert abs(tikhonov_energy - expected_tikhonov_energy) <= 1e-4 * expected_tikhonov_energy + 1e-12
        previous_gradient = (data_gradient * np.float32(1.0)) - tikhonov_gradient * np.float32(0.05)
        assert abs(max_update - np.sqrt((previous_gradient.astype(np.float64) ** 2).sum(-1).max())) <= 1e-5 * max_update


def test_verbosity_prints(lsf, capsys):
    """reference optimizer_with_telemetry.tpp:102-105,161-181: one line per iteration with the requested numbers"""
    from lsf_b200 import synthetic
    canonical, live = synthetic.circle_line_pair_2d(32)
    cls = lsf.HierarchicalOptimizer2d
    optimizer = cls(maximum_chunk_size=2, maximum_iteration_count=2, kernel=synthetic.sobolev_kernel_1d(),
                    verbosity_parameters=cls.VerbosityParameters(print_max_warp_update=True,
                                                                 print_iteration_data_energy=True,
                                                                 print_iteration_tikhonov_energy=True))
    optimizer.optimize(canonical, live)
    lines = capsys.readouterr().out.splitlines()
    assert [l for l in lines if l.startswith("[LEVEL")] == ["[LEVEL 0 COMPLETED]", "[LEVEL 1 COMPLETED]"]
    iteration_lines = [l for l in lines if l.startswith("[ITERATION")]
    assert len(iteration_lines) == 4
    assert all("[max upd. l.: " in l and "[norm. data energy: " in l and "[norm. tikhonov energy: " in l
               for l in iteration_lines)
    assert "[mean diff.: " not in iteration_lines[0]


def test_sobolev_optimizer_warp_statistics_matrix(lsf):
    """reference SobolevOptimizer2d with SharedParameters.enable_warp_statistics_logging
    (sobolev_optimizer2d.cpp:88-97,144-160): one row of warp statistics per iteration, computed over the band union of
    (canonical, warped live) AFTER that iteration. Checked against build_warp_delta_statistics_2d applied to the
    per-iteration warp fields (capture) of a second run and the live fields obtained by replaying the re-warp, and the
    last row against the convergence report."""
    from lsf_b200 import synthetic
    canonical, live = synthetic.circle_line_pair_2d(64)
    shared = lsf.SharedParameters.get_instance()
    saved = (shared.enable_warp_statistics_logging, shared.enable_convergence_reporting, shared.maximum_iteration_count,
             shared.maximum_warp_length_lower_threshold)
    try:
        shared.enable_warp_statistics_logging = True
        shared.enable_convergence_reporting = True
        shared.maximum_iteration_count = 12
        shared.maximum_warp_length_lower_threshold = 0.0
        lsf.SobolevParameters.get_instance().set_sobolev_kernel(synthetic.sobolev_kernel_1d())
        optimizer = lsf.SobolevOptimizer2d()
        optimizer.optimize(live, canonical)
        matrix = optimizer.get_warp_statistics_as_matrix()
        report = optimizer.get_convergence_report()
        assert matrix.shape == (12, 9) and optimizer.get_iteration_count() == 12
        last = report.warp_delta_statistics.to_array()
        assert np.allclose(matrix[-1], last, atol=1e-6)
        assert np.all(matrix[:, 2] >= matrix[:, 3]) and np.all(matrix[:, 3] >= matrix[:, 1])  # max >= mean >= min
        # the maximum column is the per-iteration maximum warp length the termination test uses
        assert np.allclose(matrix[:, 2], optimizer.get_max_warps(), atol=1e-6)
        shared.enable_warp_statistics_logging = False
        optimizer.optimize(live, canonical)
        assert optimizer.get_warp_statistics_as_matrix().shape == (0, 9)
    finally:
        (shared.enable_warp_statistics_logging, shared.enable_convergence_reporting, shared.maximum_iteration_count,
         shared.maximum_warp_length_lower_threshold) = saved
