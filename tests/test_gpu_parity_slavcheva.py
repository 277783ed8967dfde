"""GPU parity tests of the SobolevFusion / KillingFusion path (run on the B200 box with `-m gpu`): the CUDA kernels,
called through the C-ABI of liblsf_b200.so, against the CPU oracle on the same seeded inputs, against the
reference's golden vectors and against runs of the reference's Python optimizer.

Tolerances: the kernels restate the reference's float32 operation order without FMA, so everything is compared
BIT-EXACTLY with the oracle (np.array_equal: masks, per-iteration warp fields, final warped live field, iteration
counts, maximum warp lengths); the north-star tolerances (warp <= 1e-5 over the first 10 iterations, final warped
live <= 1e-4) are implied. Golden vectors use the reference tests' tolerances.
"""
import os

import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KERNEL3 = np.array([0.06742075, 0.99544406, 0.06742075], np.float32)


@pytest.fixture(scope="module")
def runs():
    return np.load(os.path.join(ROOT, "tests", "golden", "reference_slavcheva_runs.npz"))


@pytest.fixture()
def singletons(lsf):
    """SharedParameters / SobolevParameters are process-wide (like the reference's): restore the defaults afterwards"""
    shared, sobolev = lsf.SharedParameters.get_instance(), lsf.SobolevParameters.get_instance()
    saved = dict(vars(shared)), dict(vars(sobolev))
    yield shared, sobolev
    vars(shared).update(saved[0])
    vars(sobolev).update(saved[1])


# ----------------------------------------------------------------------------- reference goldens
def test_sobolev_optimizer_goldens(lsf, literals, singletons):
    """cpp/tests/test_slavcheva_optimizer.cpp:295-367, tests/test_slavcheva_optimizer.py:64-149"""
    shared, sobolev = singletons
    g = lambda n: literals["test_slavcheva_optimizer/" + n]
    sobolev.set_sobolev_kernel(KERNEL3)
    shared.maximum_iteration_count = 1
    optimizer = lsf.SobolevOptimizer2d()
    out = optimizer.optimize(g("test_sobolev_optimizer01/live_field"), g("test_sobolev_optimizer01/canonical_field"))
    assert np.allclose(out, g("test_sobolev_optimizer01/expected_warped_live_field_out"), rtol=0, atol=2e-7)
    shared.maximum_iteration_count = 2
    shared.maximum_warp_length_lower_threshold = 0.05
    shared.enable_convergence_reporting = True
    out = optimizer.optimize(g("test_sobolev_optimizer02/live_field"), g("test_sobolev_optimizer02/canonical_field"))
    assert np.allclose(out, g("test_sobolev_optimizer02/expected_warped_live_field_out"), rtol=0, atol=2e-7)
    expected = lsf.ConvergenceReport2d(
        2, True,
        lsf.WarpDeltaStatistics2d(0.272727, 0.0, 0.0684823, 0.0364445, 0.0167321, lsf.Vector2i(1, 2), False, False),
        lsf.TsdfDifferenceStatistics2d(0, 0.246834, 0.111843, 0.0812234, lsf.Vector2i(3, 3)))
    assert optimizer.get_convergence_report() == expected
    # the Python class reproduces the same goldens with both compute methods (tests/test_slavcheva_optimizer.py)
    for method in (lsf.ComputeMethod.DIRECT, lsf.ComputeMethod.VECTORIZED):
        python_twin = lsf.SlavchevaOptimizer2d(field_size=4, compute_method=method, sobolev_smoothing_enabled=True,
                                               sobolev_kernel=KERNEL3, max_iterations=2,
                                               maximum_warp_length_lower_threshold=0.05)
        live = g("test_sobolev_optimizer02/live_field").copy()
        returned = python_twin.optimize(live, g("test_sobolev_optimizer02/canonical_field"))
        assert returned is live  # the reference warps its argument in place
        assert np.allclose(live, g("test_sobolev_optimizer02/expected_warped_live_field_out"), rtol=0, atol=2e-7)
        assert python_twin.get_convergence_report() == expected


def test_warp_field_advanced_goldens(lsf, literals):
    """cpp/tests/test_slavcheva_optimizer.cpp:96-213, tests/test_field_warping.py:25-262"""
    g = lambda n: literals["test_slavcheva_optimizer/" + n]
    case = "warp_field_test01"
    live, _ = lsf.warp_field_advanced(g(case + "/warped_live_field"), g(case + "/canonical_field"),
                                      g(case + "/u_vectors"), g(case + "/v_vectors"))
    assert np.allclose(live, 0.0, atol=1e-6)
    case = "warp_field_test02"
    live, (u, v) = lsf.warp_field_advanced(g(case + "/warped_live_field"), g(case + "/canonical_field"),
                                           g(case + "/u_vectors"), g(case + "/v_vectors"), True, False, True)
    assert np.allclose(live, np.array([[1.0, 1.0, 1.0], [0.5, 1.0, 1.0], [1.0, 0.125, -1.0]], np.float32), atol=1e-6)
    assert np.allclose(u, g(case + "/expected_u_vectors"), atol=1e-6)
    assert np.allclose(v, g(case + "/expected_v_vectors"), atol=1e-6)
    unchanged = lsf.warp_field_advanced_no_warp_change(g(case + "/warped_live_field"), g(case + "/canonical_field"),
                                                       g(case + "/u_vectors"), g(case + "/v_vectors"), True, False, True)
    assert np.allclose(unchanged[2, 0], 1.0 - 1e-7, atol=1e-5)  # no snap bookkeeping without the warp change
    case = "warp_field_test03"
    live, _ = lsf.warp_field_advanced(g(case + "/warped_live_field"), g(case + "/canonical_field"),
                                      g(case + "/u_vectors"), g(case + "/v_vectors"), False, False, False)
    assert np.allclose(live, g(case + "/expected_live_out"), atol=2e-7)
    d = lambda n: literals["test_data_slavcheva_optimizer/" + n]
    live, _ = lsf.warp_field_advanced(d("warped_live_field/warped_live_field"), d("canonical_field/canonical_field"),
                                      d("warp_field/u_vectors"), d("warp_field/v_vectors"))
    assert np.allclose(live, g("warp_field_test04/expected_live_out"), atol=2e-7)


@pytest.mark.parametrize("flags", [(False, False, False), (True, False, False), (False, True, False),
                                   (True, False, True), (True, True, True)])
def test_warp_advanced_vs_oracle(lsf, flags):
    from lsf_b200 import slavcheva
    rng = np.random.default_rng(21)
    for shape in ((24, 24), (10, 12, 14)):
        live = np.clip(rng.standard_normal(shape) * 0.8, -1, 1).astype(np.float32)
        canonical = np.clip(rng.standard_normal(shape) * 0.8, -1, 1).astype(np.float32)
        warp = (rng.standard_normal(shape + (len(shape),)) * 1.2).astype(np.float32)
        for modify in (True, False):
            expected_live, expected_warp = oracle.warp_advanced(live, canonical, warp, *flags, modify_warp=modify)
            out_live, out_warp = slavcheva.warp_advanced(live, canonical, warp, *flags, modify_warp=modify)
            assert np.array_equal(out_live, expected_live)
            assert np.array_equal(out_warp, expected_warp)
        assert (np.abs(expected_live) == 1.0).any()


# ----------------------------------------------------------------------------- whole optimizer vs oracle, bit-exact
TERM_CASES = {
    "tikhonov": dict(smoothing_term_method=0, level_set_term_enabled=False),
    "killing": dict(smoothing_term_method=1, level_set_term_enabled=False),
    "killing_levelset": dict(smoothing_term_method=1, level_set_term_enabled=True, level_set_term_weight=0.02),
    "fdm_tikhonov": dict(smoothing_term_method=0, data_term_method=1, level_set_term_enabled=False),
}


def run_both(lsf, nd, semantics, live, canonical, iterations=12, lower=0.01, sobolev=True, kernel=None, **terms):
    from lsf_b200 import slavcheva, synthetic
    kernel = synthetic.sobolev_kernel_1d() if kernel is None else kernel
    expected = oracle.slavcheva_optimize(live, canonical, semantics=semantics, max_iterations=iterations,
                                         maximum_warp_length_lower_threshold=lower, sobolev_smoothing_enabled=sobolev,
                                         sobolev_kernel=kernel, dump_iterations=10, **terms)
    result = slavcheva._run(nd, live, canonical, semantics, terms.get("data_term_method", 0),
                            terms.get("smoothing_term_method", 0), terms.get("level_set_term_enabled", False), sobolev,
                            0.1, 1.0, 0.2, 0.1, terms.get("level_set_term_weight", 0.2), lower, 10000.0, iterations, 1,
                            kernel, collect_statistics=True, capture_iterations=10)
    assert result.iteration_count == expected["iterations"]
    assert np.array_equal(result.max_warps, expected["max_warps"])
    assert len(result.captured) == len(expected["dump"])
    assert np.abs(result.captured - expected["dump"]).max() <= 1e-5  # north star; in fact identical:
    assert np.array_equal(result.captured, expected["dump"])
    assert np.array_equal(result.warp, expected["warp"])
    assert np.abs(result.live - expected["live"]).max() <= 1e-4      # north star; in fact identical:
    assert np.array_equal(result.live, expected["live"])
    # masks: truncated voxels of the warped live field and zeroed warp vectors coincide bit for bit
    assert np.array_equal(np.abs(result.live) == 1.0, np.abs(expected["live"]) == 1.0)
    assert np.array_equal(result.warp == 0.0, expected["warp"] == 0.0)
    # convergence report against the oracle's statistics of the oracle's fields
    ws = oracle.warp_delta_statistics(expected["warp"], canonical, expected["live"], lower, 10000.0)
    ts = oracle.tsdf_difference_statistics(canonical, expected["live"])
    report = result.report
    got = report.warp_delta_statistics
    assert np.allclose([got.ratio_above_min_threshold, got.length_min, got.length_max, got.length_mean,
                        got.length_standard_deviation],
                       [ws.ratio_above_min_threshold, ws.length_min, ws.length_max, ws.length_mean,
                        ws.length_standard_deviation], rtol=1e-5, atol=1e-7)
    assert list(got.longest_warp_location) == list(ws.longest_warp_location)[:nd]
    got = report.tsdf_difference_statistics
    assert np.allclose([got.difference_min, got.difference_max, got.difference_mean, got.difference_standard_deviation],
                       [ts.difference_min, ts.difference_max, ts.difference_mean, ts.difference_standard_deviation],
                       rtol=1e-5, atol=1e-7)
    assert list(got.biggest_difference_location) == list(ts.biggest_difference_location)[:nd]
    return result, expected


@pytest.mark.parametrize("terms", sorted(TERM_CASES))
@pytest.mark.parametrize("semantics", [0, 1, 2])
def test_slavcheva2d_vs_oracle(lsf, semantics, terms):
    from lsf_b200 import synthetic
    canonical, live = synthetic.circle_line_pair_2d(64)
    result, _ = run_both(lsf, 2, semantics, live, canonical, **TERM_CASES[terms])
    assert result.iteration_count >= 2
    run_both(lsf, 2, semantics, live, canonical, sobolev=False, iterations=6, **TERM_CASES[terms])


def run_single_launch(lsf, semantics, live, canonical, iterations=12, lower=0.01, sobolev=True, rate=0.1, **terms):
    """2D runs WITHOUT a per-iteration capture take the single-launch path (csrc/slavcheva_persistent.cu: all iterations of
    a polling chunk in one cooperative kernel); compared with the oracle bit for bit and, through LSF_SLAV_PERSISTENT=0, with
    the one-launch-per-kernel path, whose launch count it must undercut"""
    from lsf_b200 import _lib, slavcheva, synthetic
    kernel = synthetic.sobolev_kernel_1d()
    expected = oracle.slavcheva_optimize(live, canonical, semantics=semantics, max_iterations=iterations,
                                         maximum_warp_length_lower_threshold=lower, sobolev_smoothing_enabled=sobolev,
                                         sobolev_kernel=kernel, gradient_descent_rate=rate, **terms)

    def run():
        before = _lib.load().lsf_launch_count()
        result = slavcheva._run(2, live, canonical, semantics, terms.get("data_term_method", 0),
                                terms.get("smoothing_term_method", 0), terms.get("level_set_term_enabled", False), sobolev,
                                rate, 1.0, 0.2, 0.1, terms.get("level_set_term_weight", 0.2), lower, 10000.0, iterations, 1,
                                kernel, collect_statistics=False, capture_iterations=0)
        return result, _lib.load().lsf_launch_count() - before

    result, launches = run()
    assert result.iteration_count == expected["iterations"]
    assert np.array_equal(result.max_warps, expected["max_warps"])
    assert np.array_equal(result.warp, expected["warp"])
    assert np.array_equal(result.live, expected["live"])
    os.environ["LSF_SLAV_PERSISTENT"] = "0"
    try:
        plain, plain_launches = run()
    finally:
        del os.environ["LSF_SLAV_PERSISTENT"]
    assert plain.iteration_count == result.iteration_count
    assert np.array_equal(plain.live, result.live) and np.array_equal(plain.warp, result.warp)
    # fields of up to 16 K voxels live in the shared memory of one thread-block cluster (k_slav_strips); LSF_SLAV_CLUSTER=1
    # keeps them in global memory (k_slav_persistent in one cluster), LSF_SLAV_CLUSTER=0 runs that kernel as a cooperative grid
    for variant in ("1", "0"):
        os.environ["LSF_SLAV_CLUSTER"] = variant
        try:
            other, _ = run()
        finally:
            del os.environ["LSF_SLAV_CLUSTER"]
        assert other.iteration_count == result.iteration_count
        assert np.array_equal(other.max_warps, result.max_warps)
        assert np.array_equal(other.live, result.live) and np.array_equal(other.warp, result.warp)
    if result.iteration_count >= 4:
        assert launches < plain_launches / 2, (launches, plain_launches)
    return result


@pytest.mark.parametrize("terms", sorted(TERM_CASES))
@pytest.mark.parametrize("semantics", [0, 1, 2])
def test_single_launch_2d_vs_oracle(lsf, semantics, terms):
    from lsf_b200 import synthetic
    canonical, live = synthetic.circle_line_pair_2d(64)
    result = run_single_launch(lsf, semantics, live, canonical, **TERM_CASES[terms])
    assert result.iteration_count >= 2
    run_single_launch(lsf, semantics, live, canonical, sobolev=False, iterations=6, **TERM_CASES[terms])


def test_single_launch_2d_config1_and_odd_shape(lsf):
    """BASELINE.json configs[0] (128 x 128, terminates through the threshold after 52 iterations, i.e. inside the launch) and
    a 90 x 90 field (last block with idle threads; the reference's 2D code is only defined for square fields)"""
    from lsf_b200 import synthetic
    canonical, live = synthetic.circle_line_pair_2d(128, shift=(5.0, -3.0), line_shift=-4.0)
    result = run_single_launch(lsf, 0, live, canonical, iterations=100, lower=0.05)
    assert result.iteration_count == 52
    result = run_single_launch(lsf, 0, live[:90, :90].copy(), canonical[:90, :90].copy(), iterations=150, lower=0.0)
    assert result.iteration_count == 150  # more iterations than one launch takes (chunks of 128)


@pytest.mark.parametrize("semantics", [0, 1, 2])
def test_single_launch_2d_long_warps(lsf, semantics):
    """a gradient-descent rate of 3 moves voxels by 14 - 65 rows within six iterations (unstable, but finite): the re-warp's taps then lie outside the rows
    a block of the cluster keeps in its own shared memory and are read from the owner's (distributed shared memory); still
    bit-identical to the oracle and to the other paths. 64 x 64 = 16 strips of 4 rows, 128 x 128 = 16 strips of 8 rows,
    48 x 48 = 16 strips of 3 rows, the 7-tap filter's halo is 3 rows"""
    from lsf_b200 import synthetic
    for size in (64, 128, 48):
        canonical, live = synthetic.circle_line_pair_2d(size, shift=(5.0, -3.0), line_shift=-4.0)
        result = run_single_launch(lsf, semantics, live, canonical, iterations=6, lower=0.0, rate=3.0)
        assert result.iteration_count == 6
        assert float(result.max_warps.max()) > 10.0  # the updates really are longer than the halo
        run_single_launch(lsf, semantics, live, canonical, iterations=6, lower=0.0, rate=3.0, sobolev=False)


def test_slavcheva2d_128_config1(lsf):
    """BASELINE.json configs[0]: 2D SobolevFusion on a 128x128 pair with the reference experiment's parameters
    (experiment/singleframe_experiment.py:91-116: rate 0.1, weights 1.0 / 0.2, lower threshold 0.05, 100 iterations,
    7-tap kernel); terminates through the threshold, identical iteration count"""
    from lsf_b200 import synthetic
    canonical, live = synthetic.circle_line_pair_2d(128, shift=(5.0, -3.0), line_shift=-4.0)
    result, expected = run_both(lsf, 2, 0, live, canonical, iterations=100, lower=0.05)
    assert result.iteration_count == 52  # the reference-semantics iteration count of this pair (CPU oracle)
    assert np.abs(result.live - canonical).mean() < np.abs(live - canonical).mean()


@pytest.mark.parametrize("terms", ["tikhonov", "killing", "killing_levelset"])
def test_slavcheva3d_vs_oracle(lsf, terms):
    from lsf_b200 import synthetic
    canonical, live = synthetic.sphere_plane_pair_3d(32)
    run_both(lsf, 3, 0, live, canonical, iterations=8, **TERM_CASES[terms])
    canonical, live = canonical[:, 2:30, :24].copy(), live[:, 2:30, :24].copy()  # non-cubic volume
    run_both(lsf, 3, 0, live, canonical, iterations=5, kernel=KERNEL3, **TERM_CASES[terms])


ORACLE_RUN_CASES = {
    "tikhonov_sobolev": dict(smoothing_term_method=0, level_set_term_enabled=False, sobolev_smoothing_enabled=True),
    "killing_levelset_sobolev": dict(smoothing_term_method=1, level_set_term_enabled=True, sobolev_smoothing_enabled=True),
    "killing_levelset_plain": dict(smoothing_term_method=1, level_set_term_enabled=True, sobolev_smoothing_enabled=False),
    "fdm_tikhonov_sobolev": dict(smoothing_term_method=0, data_term_method=1, level_set_term_enabled=False,
                                 sobolev_smoothing_enabled=True),
}


def test_python_direct_runs(lsf, runs):
    """whole runs of the reference's Python SlavchevaOptimizer2d (DIRECT): tolerance as in tests/test_oracle_slavcheva.py"""
    canonical, live, kernel = runs["runs/canonical"], runs["runs/live"], runs["runs/kernel7"]
    cases = {
        "tikhonov_sobolev": dict(smoothing_term_method=lsf.SmoothingTermMethod.TIKHONOV, sobolev_smoothing_enabled=True),
        "killing_levelset_sobolev": dict(smoothing_term_method=lsf.SmoothingTermMethod.KILLING,
                                         level_set_term_enabled=True, sobolev_smoothing_enabled=True),
        "killing_levelset_plain": dict(smoothing_term_method=lsf.SmoothingTermMethod.KILLING,
                                       level_set_term_enabled=True, sobolev_smoothing_enabled=False),
        "fdm_tikhonov_sobolev": dict(data_term_method=lsf.DataTermMethod.THRESHOLDED_FDM, sobolev_smoothing_enabled=True),
    }
    for tag, kwargs in cases.items():
        for iterations in (1, 5):
            optimizer = lsf.SlavchevaOptimizer2d(field_size=32, compute_method=lsf.ComputeMethod.DIRECT,
                                                 maximum_warp_length_lower_threshold=0.001, max_iterations=iterations,
                                                 sobolev_kernel=kernel, enable_convergence_status_logging=False, **kwargs)
            field = live.copy()
            optimizer.optimize(field, canonical)
            assert np.abs(field - runs["runs/%s/live_after_%d" % (tag, iterations)]).max() <= 2e-5
            assert np.allclose(optimizer.log.max_warps, runs["runs/%s/max_warps_%d" % (tag, iterations)], rtol=1e-5,
                               atol=1e-6)
            # OptimizationLog energies (reference slavcheva_optimizer2d.py:370-374): the reference's log at 3e-5 (it adds
            # float32 terms up sequentially), the oracle's double sums at 1e-9
            logged = np.array([optimizer.log.data_energies, optimizer.log.smoothing_energies,
                               optimizer.log.level_set_energies]).T
            assert np.allclose(logged, runs["runs/%s/energies_%d" % (tag, iterations)], rtol=3e-5, atol=1e-7)
            expected = oracle.slavcheva_optimize(live, canonical, semantics=oracle.SEMANTICS_PY_DIRECT,
                                                 max_iterations=iterations, maximum_warp_length_lower_threshold=0.001,
                                                 sobolev_kernel=kernel, **ORACLE_RUN_CASES[tag])
            assert np.allclose(logged, expected["energies"], rtol=1e-9, atol=1e-12)
            # without the log the run takes the single-launch path: same field
            quiet = lsf.SlavchevaOptimizer2d(field_size=32, compute_method=lsf.ComputeMethod.DIRECT,
                                             maximum_warp_length_lower_threshold=0.001, max_iterations=iterations,
                                             sobolev_kernel=kernel, enable_convergence_status_logging=False,
                                             log_energies=False, **kwargs)
            assert np.array_equal(quiet.optimize(live.copy(), canonical), field) and quiet.log.data_energies == []
    # ComputeMethod.VECTORIZED logs the aggregates of slavcheva_optimizer2d.py:169-175 (level-set energy stays 0)
    optimizer = lsf.SlavchevaOptimizer2d(field_size=32, compute_method=lsf.ComputeMethod.VECTORIZED,
                                         maximum_warp_length_lower_threshold=0.001, max_iterations=4, sobolev_kernel=kernel,
                                         sobolev_smoothing_enabled=True, enable_convergence_status_logging=False)
    optimizer.optimize(live.copy(), canonical)
    expected = oracle.slavcheva_optimize(live, canonical, semantics=oracle.SEMANTICS_PY_VECTORIZED, max_iterations=4,
                                         maximum_warp_length_lower_threshold=0.001, sobolev_kernel=kernel)
    logged = np.array([optimizer.log.data_energies, optimizer.log.smoothing_energies, optimizer.log.level_set_energies]).T
    assert logged.shape == (4, 3) and np.allclose(logged, expected["energies"], rtol=1e-9, atol=1e-12)
    assert logged[1:, 1].min() > 0 and not logged[:, 2].any()


def test_degenerate_volume_reproduces_2d(lsf):
    """3D kernels on a volume that is constant along axis 2 == 2D kernels (same property as the oracle test)"""
    from lsf_b200 import synthetic, slavcheva
    canonical2, live2 = synthetic.circle_line_pair_2d(32)
    to3 = lambda f: np.repeat(f.T[:, :, None], 14, axis=2).copy()
    kwargs = dict(smoothing_term_method=lsf.SmoothingTermMethod.KILLING, level_set_term_enabled=True,
                  level_set_term_weight=0.02, sobolev_smoothing_enabled=False, max_iterations=4,
                  maximum_warp_length_lower_threshold=1e-4)
    flat = slavcheva.SlavchevaOptimizer2dCpp(**kwargs)
    volume = lsf.SlavchevaOptimizer3d(**kwargs)
    live_flat = flat.optimize(live2, canonical2)
    live_volume = volume.optimize(to3(live2), to3(canonical2))
    assert flat.get_iteration_count() == volume.get_iteration_count() == 4
    for z in (6, 7):
        assert np.array_equal(live_volume[:, :, z], live_flat.T)
        assert np.array_equal(volume.get_last_warp_field()[:, :, z, 0], flat.get_last_warp_field()[:, :, 0].T)
        assert np.array_equal(volume.get_last_warp_field()[:, :, z, 1], flat.get_last_warp_field()[:, :, 1].T)


def test_termination_rules(lsf):
    """reference optimizer2d.cpp:76-82 / slavcheva_optimizer2d.py:360-362: minimum iteration count, lower / upper
    thresholds, min > max"""
    from lsf_b200 import synthetic
    canonical, live = synthetic.circle_line_pair_2d(32)
    for semantics in (0, 1):
        for kwargs in (dict(max_iterations=3, min_iterations=5), dict(max_iterations=20, min_iterations=4,
                                                                      maximum_warp_length_lower_threshold=10.0),
                       dict(max_iterations=20, maximum_warp_length_upper_threshold=0.01),
                       dict(max_iterations=0, min_iterations=0), dict(max_iterations=17)):
            expected = oracle.slavcheva_optimize(live, canonical, semantics=semantics, sobolev_kernel=KERNEL3, **kwargs)
            p = dict(maximum_warp_length_lower_threshold=0.1, maximum_warp_length_upper_threshold=10000.0,
                     max_iterations=100, min_iterations=1)
            p.update(kwargs)
            from lsf_b200 import slavcheva
            result = slavcheva._run(2, live, canonical, semantics, 0, 0, False, True, 0.1, 1.0, 0.2, 0.1, 0.2,
                                    p["maximum_warp_length_lower_threshold"], p["maximum_warp_length_upper_threshold"],
                                    p["max_iterations"], p["min_iterations"], KERNEL3)
            assert result.iteration_count == expected["iterations"], (semantics, kwargs)
            assert np.array_equal(result.live, expected["live"])


def test_device_tensors_and_errors(lsf):
    import torch
    from lsf_b200 import synthetic
    canonical, live = synthetic.sphere_plane_pair_3d(32)
    optimizer = lsf.SlavchevaOptimizer3d(smoothing_term_method=lsf.SmoothingTermMethod.KILLING, max_iterations=4,
                                         maximum_warp_length_lower_threshold=0.0)
    host = optimizer.optimize(live, canonical)
    device = optimizer.optimize(torch.from_numpy(live).cuda(), torch.from_numpy(canonical).cuda())
    assert device.is_cuda and np.array_equal(device.cpu().numpy(), host)
    with pytest.raises(ValueError):
        lsf.SobolevOptimizer2d().optimize(np.zeros((8, 6), np.float32), np.zeros((8, 6), np.float32))
    with pytest.raises(ValueError):
        lsf.SlavchevaOptimizer3d().optimize(live, canonical[:16])


# ----------------------------------------------------------------------------- telemetry builders
def test_telemetry_builders_vs_oracle(lsf):
    rng = np.random.default_rng(31)
    for shape in ((20, 28), (9, 11, 13)):
        nd = len(shape)
        live = np.clip(rng.standard_normal(shape), -1, 1).astype(np.float32)
        canonical = np.clip(rng.standard_normal(shape), -1, 1).astype(np.float32)
        warp = rng.standard_normal(shape + (nd,)).astype(np.float32)
        warp[rng.random(shape) < 0.2] = 0.0
        build_w = lsf.build_warp_delta_statistics_2d if nd == 2 else lsf.build_warp_delta_statistics_3d
        build_d = lsf.build_tsdf_difference_statistics_2d if nd == 2 else lsf.build_tsdf_difference_statistics_3d
        got, ws = build_w(warp, canonical, live, 0.5, 2.5), oracle.warp_delta_statistics(warp, canonical, live, 0.5, 2.5)
        assert np.allclose([got.ratio_above_min_threshold, got.length_min, got.length_max, got.length_mean,
                            got.length_standard_deviation],
                           [ws.ratio_above_min_threshold, ws.length_min, ws.length_max, ws.length_mean,
                            ws.length_standard_deviation], rtol=1e-6, atol=1e-7)
        assert list(got.longest_warp_location) == list(ws.longest_warp_location)[:nd]
        assert got.is_largest_above_max_threshold == bool(ws.is_largest_above_max_threshold)
        got, ts = build_d(canonical, live), oracle.tsdf_difference_statistics(canonical, live)
        assert np.allclose([got.difference_min, got.difference_max, got.difference_mean,
                            got.difference_standard_deviation],
                           [ts.difference_min, ts.difference_max, ts.difference_mean, ts.difference_standard_deviation],
                           rtol=1e-6, atol=1e-7)
        assert list(got.biggest_difference_location) == list(ts.biggest_difference_location)[:nd]
        assert abs(lsf.mean_vector_length(warp) - np.linalg.norm(warp, axis=-1).mean()) < 1e-5


def test_hierarchical_convergence_reports(lsf):
    """reference optimizer_with_telemetry.tpp:107-124: per-level reports = statistics of the level's warp field over the
    band union of the (un-warped) pyramid levels; checked for the finest level, where the level fields are the inputs"""
    from lsf_b200 import synthetic
    for nd, (canonical, live) in ((3, synthetic.sphere_plane_pair_3d(32)), (2, synthetic.circle_line_pair_2d(64))):
        cls = lsf.HierarchicalOptimizer3d if nd == 3 else lsf.HierarchicalOptimizer2d
        optimizer = cls(tikhonov_term_enabled=False, gradient_kernel_enabled=True, kernel=synthetic.sobolev_kernel_1d(),
                        maximum_chunk_size=4, maximum_iteration_count=10, maximum_warp_update_threshold=0.01,
                        logging_parameters=cls.LoggingParameters(collect_per_level_convergence_reports=True))
        warp = optimizer.optimize(canonical, live)
        reports = optimizer.get_per_level_convergence_reports()
        assert len(reports) == 3 and all(r.iteration_count == 10 and r.iteration_limit_reached for r in reports)
        ws = oracle.warp_delta_statistics(warp, canonical, live, 0.01, 3.0e38)
        ts = oracle.tsdf_difference_statistics(canonical, live)
        got = reports[-1].warp_delta_statistics
        assert np.allclose([got.ratio_above_min_threshold, got.length_min, got.length_max, got.length_mean,
                            got.length_standard_deviation],
                           [ws.ratio_above_min_threshold, ws.length_min, ws.length_max, ws.length_mean,
                            ws.length_standard_deviation], rtol=1e-5, atol=1e-7)
        assert list(got.longest_warp_location) == list(ws.longest_warp_location)[:nd]
        got = reports[-1].tsdf_difference_statistics
        assert np.allclose([got.difference_min, got.difference_max, got.difference_mean,
                            got.difference_standard_deviation],
                           [ts.difference_min, ts.difference_max, ts.difference_mean, ts.difference_standard_deviation],
                           rtol=1e-5, atol=1e-7)
        assert list(got.biggest_difference_location) == list(ts.biggest_difference_location)[:nd]


# ----------------------------------------------------------------------------- full-size properties (256^3)
def test_full_size_properties_256(lsf):
    """BASELINE.json configs[2] size (256^3 KillingFusion): the oracle is too slow for a dense comparison inside a unit
    test, so check size-independent properties: identical inputs stop after the minimum iteration with an exactly
    zero warp; truncated-everywhere voxels never change; two runs are bit-identical; the data term decreases."""
    import torch
    from lsf_b200 import synthetic
    canonical, live = synthetic.sphere_plane_pair_3d(256, xp=torch, device="cuda")
    optimizer = lsf.SlavchevaOptimizer3d(smoothing_term_method=lsf.SmoothingTermMethod.KILLING,
                                         level_set_term_enabled=True, level_set_term_weight=0.02, max_iterations=5,
                                         maximum_warp_length_lower_threshold=0.001)
    # identical inputs: data and Killing terms vanish (the level-set term does not: it regularises |grad| towards 1)
    plain = lsf.SlavchevaOptimizer3d(smoothing_term_method=lsf.SmoothingTermMethod.KILLING, max_iterations=5,
                                     maximum_warp_length_lower_threshold=0.001)
    same = plain.optimize(canonical, canonical)
    assert plain.get_iteration_count() == 1 and bool((same == canonical).all())
    assert float(plain.get_last_warp_field().abs().max()) == 0.0
    first = optimizer.optimize(live, canonical)
    assert optimizer.get_iteration_count() == 5
    outside = (live.abs() == 1.0) & (canonical.abs() == 1.0)
    assert bool((first[outside] == live[outside]).all())
    assert float((first - canonical).abs().mean()) < float((live - canonical).abs().mean())
    second = optimizer.optimize(live, canonical)
    assert bool((first == second).all())


@pytest.mark.parametrize("taps", [3, 5, 7])
def test_fast_filter_matches_three_pass_filter(lsf, taps, monkeypatch):
    """A/B: the marching Sobolev filter kernels of the 3D optimizer (slavcheva_fast.cuh) against the first-generation
    three-pass kernel (LSF_SLAV_FAST=0) -- bit-identical live field, warp field and iteration count, on a ragged
    volume and on a long thin one whose rows span three z tiles (halo columns between tiles)."""
    from lsf_b200 import synthetic
    canonical, live = synthetic.sphere_plane_pair_3d(64)
    cases = [(canonical[10:50, 14:50, :].copy(), live[10:50, 14:50, :].copy())]
    rng = np.random.default_rng(5)
    long_c = np.clip(np.cumsum(rng.standard_normal((6, 10, 1040)).astype(np.float32) * 0.05, axis=2), -1, 1)
    long_l = np.clip(long_c + rng.standard_normal(long_c.shape).astype(np.float32) * 0.02, -1, 1).astype(np.float32)
    cases.append((long_c.astype(np.float32), long_l))
    for canonical_case, live_case in cases:
        results = []
        # default (narrow-band sparse iteration) | dense with band-compacted gradient / re-warp and marching filter |
        # without the band compaction | re-warp in the filter kernel's epilogue | first-generation kernels
        # (brick-ordered, TMA-staged persistent kernels since round 2; LSF_SLAV_BRICK=0: memory-ordered global list)
        for switches in ({}, {"LSF_SLAV_BRICK": "0"}, {"LSF_SLAV_SPARSE": "0"}, {"LSF_SLAV_BAND": "0"},
                         {"LSF_SLAV_SPARSE": "0", "LSF_SLAV_FUSE_REWARP": "1"}, {"LSF_SLAV_FAST": "0"}):
            for name in ("LSF_SLAV_SPARSE", "LSF_SLAV_BAND", "LSF_SLAV_FUSE_REWARP", "LSF_SLAV_FAST", "LSF_SLAV_BRICK"):
                monkeypatch.delenv(name, raising=False)
            for name, value in switches.items():
                monkeypatch.setenv(name, value)
            optimizer = lsf.SlavchevaOptimizer3d(smoothing_term_method=lsf.SmoothingTermMethod.KILLING,
                                                 level_set_term_enabled=True, max_iterations=6,
                                                 maximum_warp_length_lower_threshold=0.0,
                                                 sobolev_kernel=synthetic.sobolev_kernel_1d(taps))
            out = optimizer.optimize(live_case.copy(), canonical_case)
            results.append((np.array(out), np.array(optimizer.get_last_warp_field()), optimizer.get_iteration_count()))
        assert np.abs(results[0][1]).max() > 0
        for other in results[1:]:
            assert results[0][2] == other[2]
            assert np.array_equal(results[0][0], other[0])
            assert np.array_equal(results[0][1], other[1])


def test_sparse_iteration_long_run_with_band_shrinkage(lsf, monkeypatch):
    """The narrow-band sparse iteration keeps invariants outside the band (zero update fields, equal live buffers) and
    patches them where a voxel leaves the band (k_slav_band_leave). A long run in which the band demonstrably shrinks
    must stay bit-identical to the dense kernels and to the CPU oracle."""
    from lsf_b200 import synthetic
    canonical, live = synthetic.sphere_plane_pair_3d(64)
    canonical, live = canonical[8:56, 8:56, 8:56].copy(), live[8:56, 8:56, 8:56].copy()
    results = []
    # sparse with the default scan period (the band list is re-used for 32 iterations: voxels that leave the band in
    # between stay listed and are recognised) | list rebuilt every iteration | every 16 iterations | dense
    # the same with the memory-ordered list (LSF_SLAV_BRICK=0)
    for sparse, rescan, brick in (("1", None, "1"), ("1", "1", "1"), ("1", "16", "1"), ("0", None, "1"), ("1", None, "0"),
                                  ("1", "16", "0")):
        monkeypatch.setenv("LSF_SLAV_SPARSE", sparse)
        monkeypatch.setenv("LSF_SLAV_BRICK", brick)
        if rescan is None:
            monkeypatch.delenv("LSF_SLAV_RESCAN", raising=False)
        else:
            monkeypatch.setenv("LSF_SLAV_RESCAN", rescan)
        optimizer = lsf.SlavchevaOptimizer3d(smoothing_term_method=lsf.SmoothingTermMethod.KILLING,
                                             level_set_term_enabled=True, max_iterations=60, min_iterations=60,
                                             maximum_warp_length_lower_threshold=0.0, gradient_descent_rate=0.2,
                                             sobolev_kernel=synthetic.sobolev_kernel_1d())
        out = np.array(optimizer.optimize(live.copy(), canonical))
        results.append((out, np.array(optimizer.get_last_warp_field()), np.array(optimizer.get_max_warps())))
    for other in results[1:]:
        assert np.array_equal(results[0][0], other[0])
        assert np.array_equal(results[0][1], other[1])
        assert np.array_equal(results[0][2], other[2])
    outside_before = int(((np.abs(live) == 1.0) & (np.abs(canonical) == 1.0)).sum())
    outside_after = int(((np.abs(results[0][0]) == 1.0) & (np.abs(canonical) == 1.0)).sum())
    assert outside_after > outside_before, (outside_before, outside_after)
    expected = oracle.slavcheva_optimize(live, canonical, semantics=0, smoothing_term_method=1, level_set_term_enabled=True,
                                         max_iterations=60, min_iterations=60, maximum_warp_length_lower_threshold=0.0,
                                         gradient_descent_rate=0.2, sobolev_kernel=synthetic.sobolev_kernel_1d())
    assert expected["iterations"] == 60
    assert np.array_equal(results[0][0], expected["live"])
    assert np.array_equal(results[0][1], expected["warp"])
