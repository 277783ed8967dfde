"""Pins the slavcheva (SobolevFusion / KillingFusion) part of the CPU oracle (oracle/lsf_oracle_slavcheva.cpp):

* C++ semantics against the reference's own golden vectors (cpp/tests/test_slavcheva_optimizer.cpp,
  cpp/tests/data/test_data_slavcheva_optimizer.hpp, tests/test_slavcheva_optimizer.py) -- reference_literals.npz;
* Python semantics and the Killing / level-set / thresholded-FDM terms against RUNS of the reference's Python code
  (tests/golden/make_golden.py collect_slavcheva_runs) -- reference_slavcheva_runs.npz;
* the 3D generalisation (no reference counterpart) through the degenerate-volume == 2D property.
"""
import os

import numpy as np
import pytest

import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KERNEL3 = np.array([0.06742075, 0.99544406, 0.06742075], np.float32)  # cpp/tests/test_slavcheva_optimizer.cpp:297-298


@pytest.fixture(scope="module")
def runs():
    return np.load(os.path.join(ROOT, "tests", "golden", "reference_slavcheva_runs.npz"))


def stack_uv(u, v):
    return np.stack([u, v], axis=-1).astype(np.float32)


# ----------------------------------------------------------------------------- whole optimizer, reference goldens
@pytest.mark.parametrize("semantics", [oracle.SEMANTICS_CPP, oracle.SEMANTICS_PY_DIRECT, oracle.SEMANTICS_PY_VECTORIZED])
def test_sobolev_optimizer_goldens(literals, semantics):
    """cpp/tests/test_slavcheva_optimizer.cpp:295-367, tests/test_slavcheva_optimizer.py:64-149 (the same expected
    arrays serve the C++ optimizer and both Python compute methods)"""
    g = lambda n: literals["test_slavcheva_optimizer/" + n]
    for case, iterations, extra in (("test_sobolev_optimizer01", 1, {}),
                                    ("test_sobolev_optimizer02", 2, dict(maximum_warp_length_lower_threshold=0.05))):
        result = oracle.slavcheva_optimize(g(case + "/live_field"), g(case + "/canonical_field"), semantics=semantics,
                                           max_iterations=iterations, sobolev_kernel=KERNEL3, **extra)
        assert result["iterations"] == iterations
        # Eigen isApprox (relative 1e-5 of the norm); the literals carry 8 significant digits
        assert np.allclose(result["live"], g(case + "/expected_warped_live_field_out"), rtol=0, atol=2e-7)


def test_convergence_report_golden(literals):
    """cpp/tests/test_slavcheva_optimizer.cpp:354-366: WarpDeltaStatistics2d(0.272727, 0.0, 0.0684823, 0.0364445,
    0.0167321, (1,2), false, false), TsdfDifferenceStatistics2d(0, 0.246834, 0.111843, 0.0812234, (3,3))"""
    g = lambda n: literals["test_slavcheva_optimizer/test_sobolev_optimizer02/" + n]
    result = oracle.slavcheva_optimize(g("live_field"), g("canonical_field"), max_iterations=2, sobolev_kernel=KERNEL3,
                                       maximum_warp_length_lower_threshold=0.05)
    ws = oracle.warp_delta_statistics(result["warp"], g("canonical_field"), result["live"], 0.05, 10000.0)
    assert np.allclose([ws.ratio_above_min_threshold, ws.length_min, ws.length_max, ws.length_mean,
                        ws.length_standard_deviation], [0.272727, 0.0, 0.0684823, 0.0364445, 0.0167321], atol=1e-6)
    assert list(ws.longest_warp_location)[:2] == [1, 2]
    assert not ws.is_largest_below_min_threshold and not ws.is_largest_above_max_threshold
    ts = oracle.tsdf_difference_statistics(g("canonical_field"), result["live"])
    assert np.allclose([ts.difference_min, ts.difference_max, ts.difference_mean, ts.difference_standard_deviation],
                       [0.0, 0.246834, 0.111843, 0.0812234], atol=1e-6)
    assert list(ts.biggest_difference_location)[:2] == [3, 3]


# ----------------------------------------------------------------------------- warp_2d_advanced goldens
def test_warp_advanced_goldens(literals):
    """cpp/tests/test_slavcheva_optimizer.cpp:96-213, tests/test_field_warping.py:25-262"""
    g = lambda n: literals["test_slavcheva_optimizer/" + n]
    # 01: default flags
    case = "warp_field_test01"
    live, _ = oracle.warp_advanced(g(case + "/warped_live_field"), g(case + "/canonical_field"),
                                   stack_uv(g(case + "/u_vectors"), g(case + "/v_vectors")))
    assert np.allclose(live, 0.0, atol=1e-6)
    # 02: band_union_only=True, substitute_original=True; the warp is zeroed where the result snaps to +-1
    case = "warp_field_test02"
    live, warp = oracle.warp_advanced(g(case + "/warped_live_field"), g(case + "/canonical_field"),
                                      stack_uv(g(case + "/u_vectors"), g(case + "/v_vectors")), band_union_only=True,
                                      substitute_original=True)
    expected_live = np.array([[1.0, 1.0, 1.0], [0.5, 1.0, 1.0], [1.0, 0.125, -1.0]], np.float32)
    assert np.allclose(live, expected_live, atol=1e-6)
    assert np.allclose(warp, stack_uv(g(case + "/expected_u_vectors"), g(case + "/expected_v_vectors")), atol=1e-6)
    # 03: explicit all-false flags
    case = "warp_field_test03"
    live, _ = oracle.warp_advanced(g(case + "/warped_live_field"), g(case + "/canonical_field"),
                                   stack_uv(g(case + "/u_vectors"), g(case + "/v_vectors")))
    assert np.allclose(live, g(case + "/expected_live_out"), atol=2e-7)
    # 04: the static test data
    d = lambda n: literals["test_data_slavcheva_optimizer/" + n]
    live, _ = oracle.warp_advanced(d("warped_live_field/warped_live_field"), d("canonical_field/canonical_field"),
                                   stack_uv(d("warp_field/u_vectors"), d("warp_field/v_vectors")))
    assert np.allclose(live, g("warp_field_test04/expected_live_out"), atol=2e-7)


# ----------------------------------------------------------------------------- data / Tikhonov term goldens
def test_data_term_goldens(literals):
    """cpp/tests/test_slavcheva_optimizer.cpp:215-236 + data/test_data_slavcheva_optimizer.hpp"""
    d = lambda n: literals["test_data_slavcheva_optimizer/" + n]
    live, canonical = d("warped_live_field/warped_live_field"), d("canonical_field/canonical_field")
    assert np.allclose(oracle.slavcheva_data_term(live, canonical), d("data_term_gradient/grad"), atol=1e-6)
    assert np.allclose(oracle.slavcheva_data_term(live, canonical, band_union_only=True),
                       d("data_term_gradient_band_union_only/grad"), atol=1e-6)


def test_tikhonov_term_goldens(literals):
    """cpp/tests/test_slavcheva_optimizer.cpp:238-293"""
    g = lambda n: literals["test_slavcheva_optimizer/test_tikhonov_regularization_gradient01/" + n]
    out = oracle.slavcheva_smoothing_term(g("warp_field"), g("live_field"), g("canonical_field"), band_union_only=True)
    assert np.allclose(out, g("expected_gradient_out"), atol=1e-6)
    d = lambda n: literals["test_data_slavcheva_optimizer/" + n]
    warp = (d("data_term_gradient_band_union_only/grad") * np.float32(0.1)).astype(np.float32)
    assert np.allclose(oracle.slavcheva_smoothing_term(warp), d("tikhonov_gradient/grad"), atol=1e-6)
    out = oracle.slavcheva_smoothing_term(warp, d("warped_live_field2/warped_live_field"),
                                          d("canonical_field/canonical_field"), band_union_only=True)
    assert np.allclose(out, d("tikhonov_gradient_band_union_only/grad"), atol=1e-6)


# ----------------------------------------------------------------------------- reference Python runs
def test_python_terms(runs):
    """per-voxel functions of the reference's Python code: smoothing_term.py:50-139, level_set_term.py:28-64,
    data_term.py:169-227 -- the restatement is bit-identical on these seeded fields"""
    warp, live, canonical = runs["terms/warp"], runs["terms/live"], runs["terms/canonical"]
    py = dict(semantics=oracle.SEMANTICS_PY_DIRECT)
    assert np.array_equal(oracle.slavcheva_smoothing_term(warp, smoothing_term_method=oracle.SMOOTHING_KILLING, **py),
                          runs["terms/killing_lambda0.1"])
    assert np.array_equal(oracle.slavcheva_smoothing_term(warp, **py), runs["terms/tikhonov_direct"])
    assert np.allclose(oracle.slavcheva_smoothing_term(warp), runs["terms/tikhonov_direct"], atol=1e-6)
    assert np.array_equal(oracle.slavcheva_level_set_term(live), runs["terms/level_set"])
    assert np.array_equal(oracle.slavcheva_data_term(live, canonical, **py), runs["terms/data_basic"])
    assert np.allclose(oracle.slavcheva_data_term(live, canonical, data_term_method=oracle.DATA_TERM_THRESHOLDED_FDM,
                                                  **py), runs["terms/data_thresholded_fdm"], atol=1e-7)


PYTHON_RUN_CASES = {
    "tikhonov_sobolev": dict(smoothing_term_method=0, level_set_term_enabled=False, sobolev_smoothing_enabled=True),
    "killing_levelset_sobolev": dict(smoothing_term_method=1, level_set_term_enabled=True,
                                     sobolev_smoothing_enabled=True),
    "killing_levelset_plain": dict(smoothing_term_method=1, level_set_term_enabled=True,
                                   sobolev_smoothing_enabled=False),
    "fdm_tikhonov_sobolev": dict(smoothing_term_method=0, data_term_method=1, level_set_term_enabled=False,
                                 sobolev_smoothing_enabled=True),
}


@pytest.mark.parametrize("tag", sorted(PYTHON_RUN_CASES))
def test_python_direct_runs(runs, tag):
    """whole runs of the reference's Python SlavchevaOptimizer2d (DIRECT), 32x32, 1 and 5 iterations. The Python code
    mixes float64 scalars into float32 arrays and filters with np.convolve; tolerance 2e-5 on the warped live field
    (values in [-1,1]) and 1e-5 on the per-iteration maximum warp lengths."""
    canonical, live, kernel = runs["runs/canonical"], runs["runs/live"], runs["runs/kernel7"]
    for iterations in (1, 5):
        result = oracle.slavcheva_optimize(live, canonical, semantics=oracle.SEMANTICS_PY_DIRECT,
                                           max_iterations=iterations, maximum_warp_length_lower_threshold=0.001,
                                           sobolev_kernel=kernel, **PYTHON_RUN_CASES[tag])
        assert result["iterations"] == iterations
        assert np.abs(result["live"] - runs["runs/%s/live_after_%d" % (tag, iterations)]).max() <= 2e-5
        assert np.allclose(result["max_warps"], runs["runs/%s/max_warps_%d" % (tag, iterations)], rtol=1e-5, atol=1e-6)
        # OptimizationLog.data_energies / smoothing_energies / level_set_energies (slavcheva_optimizer2d.py:370-374): the
        # reference adds float32 terms up sequentially (1.2e-5 off the exact sum at iteration 0), the oracle in double
        expected = runs["runs/%s/energies_%d" % (tag, iterations)]
        assert result["energies"].shape == expected.shape == (iterations, 3)
        assert np.allclose(result["energies"], expected, rtol=3e-5, atol=1e-7)
        assert (expected[:, 2] > 0).all() == bool(PYTHON_RUN_CASES[tag]["level_set_term_enabled"])


def test_vectorized_energy_aggregates(runs):
    """ComputeMethod.VECTORIZED (slavcheva_optimizer2d.py:169-175): data_term.py:352-358 and smoothing_term.py:162-177
    (np.gradient of the warp components over the band union) evaluated by the reference on random 9 x 9 fields"""
    warp, live, canonical = runs["terms/warp"], runs["terms/live"], runs["terms/canonical"]
    energies = oracle.slavcheva_energies(live, canonical, warp, semantics=oracle.SEMANTICS_PY_VECTORIZED,
                                         data_term_weight=1.0, smoothing_term_weight=1.0)
    assert np.allclose(energies[:2], runs["terms/vectorized_energies"], rtol=1e-6)
    assert energies[2] == 0.0
    weighted = oracle.slavcheva_energies(live, canonical, warp, semantics=oracle.SEMANTICS_PY_VECTORIZED,
                                         data_term_weight=2.0, smoothing_term_weight=0.25)
    assert np.allclose(weighted[:2], runs["terms/vectorized_energies"] * [2.0, 0.25], rtol=1e-6)
    # the C++ optimizer keeps no energy log (sobolev_optimizer2d.cpp:121-138 drops them): zeros
    assert not oracle.slavcheva_energies(live, canonical, warp, semantics=oracle.SEMANTICS_CPP).any()


# ----------------------------------------------------------------------------- 3D generalisation
@pytest.mark.parametrize("smoothing,level_set", [(0, False), (1, False), (1, True)])
def test_degenerate_volume_reproduces_2d(smoothing, level_set):
    """a 3D pair that is constant along axis 2 must reproduce the 2D result plane by plane away from
    the axis-2 borders (there the level-set stencil sees the out-of-bounds value 1 and the Killing cross terms see the
    centre-value replacement; the disturbance travels one plane per iteration). 2D x (component 0) <-> 3D
    axis 0, 2D y (component 1) <-> 3D axis 1, i.e. volume[x, y, z] = image[y, x]. The Sobolev filter is off: its
    axis-2 pass scales a constant line by the tap sum."""
    from lsf_b200 import synthetic
    canonical2, live2 = synthetic.circle_line_pair_2d(32)
    depth = 14
    to3 = lambda f: np.repeat(f.T[:, :, None], depth, axis=2).copy()
    kwargs = dict(smoothing_term_method=smoothing, level_set_term_enabled=level_set, sobolev_smoothing_enabled=False,
                  max_iterations=4, maximum_warp_length_lower_threshold=1e-4, level_set_term_weight=0.02)
    r2 = oracle.slavcheva_optimize(live2, canonical2, dump_iterations=4, **kwargs)
    r3 = oracle.slavcheva_optimize(to3(live2), to3(canonical2), dump_iterations=4, **kwargs)
    assert r2["iterations"] == r3["iterations"] == 4
    for z in (6, 7):
        assert np.array_equal(r3["live"][:, :, z], r2["live"].T)
        assert np.array_equal(r3["warp"][:, :, z, 0], r2["warp"][:, :, 0].T)
        assert np.array_equal(r3["warp"][:, :, z, 1], r2["warp"][:, :, 1].T)
        # the level-set cross terms (A - B - A + B in float32) leave rounding noise in the axis-2 component
        assert np.abs(r3["warp"][:, :, z, 2]).max() < 1e-7
    assert np.abs(r2["warp"]).max() > 1e-3


def test_3d_filter_matches_the_pinned_3d_convolution():
    """the slavcheva oracle's 3D preserve-zeros filter on a field without zero vectors equals the hierarchical
    oracle's convolve3d (pinned by the reference's 3D convolution golden)"""
    rng = np.random.default_rng(11)
    live = np.clip(rng.standard_normal((10, 9, 8)) * 0.4, -0.9, 0.9).astype(np.float32)
    canonical = (live * 0.7 + 0.1).astype(np.float32)
    kernel = rng.random(5).astype(np.float32)
    plain = oracle.slavcheva_optimize(live, canonical, sobolev_smoothing_enabled=False, max_iterations=1,
                                      maximum_warp_length_lower_threshold=0.0)
    filtered = oracle.slavcheva_optimize(live, canonical, sobolev_smoothing_enabled=True, sobolev_kernel=kernel,
                                         max_iterations=1, maximum_warp_length_lower_threshold=0.0)
    # the warp before the resample step: no value snaps to +-1 here, so the returned warp is the filtered update
    assert np.abs(plain["warp"]).max(axis=-1).all()  # no exactly-zero vector: the preserve-zeros rule stays idle
    assert np.array_equal(filtered["warp"], oracle.convolve_with_kernel(plain["warp"], kernel))


def test_square_2d_only():
    field = np.zeros((8, 6), np.float32)
    with pytest.raises(RuntimeError):
        oracle.slavcheva_optimize(field, field)
