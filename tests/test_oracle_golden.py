"""Pins the CPU oracle (oracle/lsf_oracle.cpp) against the reference's own golden vectors.

Fixtures: tests/golden/reference_literals.npz (literals of the reference's Python and C++ tests) and
tests/golden/reference_python_runs.npz (outputs of the reference's Python implementation), both made by
tests/golden/make_golden.py. Each test names the reference test it mirrors.
"""
import numpy as np
import pytest

import oracle


def arange_tensor3v3f(shape):
    """cpp/tests/test_convolution.cpp:241-252 `gen_arange_tensor3v3f`: x fastest, then y, then z."""
    X, Y, Z = shape
    x, y, z = np.meshgrid(np.arange(X), np.arange(Y), np.arange(Z), indexing="ij")
    base = 1.0 + 3.0 * (x + X * y + X * Y * z)
    return np.stack([base, base + 1, base + 2], axis=-1).astype(np.float32)


# ----------------------------------------------------------------------------- hierarchical optimizer, 2D
def test_hierarchical_optimizer01(literals):
    """cpp/tests/test_hierarchical_optimizer.cpp:161-178, tests/test_hierarchical_optimizer2d.py:39-69"""
    for prefix in ("py_hierarchical/%s", "test_data_hierarchical_optimizer/%s/%s"):
        def get(name):
            return literals[prefix % ((name,) * prefix.count("%s"))]
        canonical, live = get("canonical_field"), get("live_field")
        result = oracle.hier_optimize(canonical, live, tikhonov_term_enabled=False, gradient_kernel_enabled=False,
                                      maximum_chunk_size=8, rate=0.2, maximum_iteration_count=100,
                                      maximum_warp_update_threshold=0.001, data_term_amplifier=1.0)
        assert np.allclose(result["warp"], get("warp_field"), atol=10e-6)
        final_live = oracle.warp(live, result["warp"])
        assert np.allclose(final_live, get("final_live_field"), atol=10e-6)
        assert result["iterations"] == [1, 100, 100, 100]


def test_hierarchical_optimizer_iteration_data(literals):
    """cpp/tests/test_hierarchical_optimizer.cpp:180-205: warp field of level 3 after iteration 50"""
    canonical = literals["py_hierarchical/canonical_field"]
    live = literals["py_hierarchical/live_field"]
    result = oracle.hier_optimize(canonical, live, tikhonov_term_enabled=False, gradient_kernel_enabled=False,
                                  maximum_chunk_size=8, rate=0.2, maximum_iteration_count=100,
                                  maximum_warp_update_threshold=0.001, data_term_amplifier=1.0,
                                  tikhonov_strength=0.0, dump_level=3, dump_iterations=100)
    assert len(result["dump"]) == 100
    expected = literals["py_hierarchical/iteration50_warp_field"]
    assert np.allclose(result["dump"][50], expected, atol=1e-6)
    # the golden literal carries enough digits for a bit-exact comparison
    assert np.array_equal(result["dump"][50], expected)
    assert np.allclose(result["dump"][50], literals["test_data_hierarchical_optimizer/iteration50_warp_field/mat"],
                       atol=1e-6)


def test_warp_field(literals):
    """cpp/tests/test_hierarchical_optimizer.cpp:149-159, tests/test_field_warping.py:264-275"""
    g = lambda n: literals["py_hierarchical/" + n]
    out = oracle.warp(g("field_A_16x16"), g("warp_field_A_16x16"))
    assert np.allclose(out, g("fA_resampled_with_wfA"), atol=1e-6)
    out = oracle.warp_with_replacement(g("field_B_16x16"), g("warp_field_B_16x16"), 0.0)
    assert np.allclose(out, g("fB_resampled_with_wfB_replacement"), atol=1e-6)


def test_warp_python_reference(python_runs):
    """reference Python nonrigid_opt/field_warping.py:67-109 on a seeded 16x16 field"""
    f, w = python_runs["warp2d/field"], python_runs["warp2d/warp"]
    assert np.allclose(oracle.warp(f, w), python_runs["warp2d/out"], atol=1e-6)
    assert np.allclose(oracle.warp_with_replacement(f, w, 0.0), python_runs["warp2d/out_replacement0"], atol=1e-6)


def test_hierarchical_python_reference_full(python_runs):
    """reference Python HierarchicalOptimizer2d (hierarchical_optimizer2d.py:123-248) run with the Tikhonov
    term and the Sobolev kernel enabled -- modes the reference's own fixtures do not cover."""
    canonical, live = python_runs["hier2d_full/canonical"], python_runs["hier2d_full/live"]
    kernel = python_runs["kernel7"]
    modes = {
        "data_only": dict(tikhonov_term_enabled=False, gradient_kernel_enabled=False),
        "tikhonov": dict(tikhonov_term_enabled=True, gradient_kernel_enabled=False, tikhonov_strength=0.1),
        "tikhonov_kernel": dict(tikhonov_term_enabled=True, gradient_kernel_enabled=True, tikhonov_strength=0.2,
                                kernel=kernel),
    }
    for tag, kwargs in modes.items():
        result = oracle.hier_optimize(canonical, live, maximum_chunk_size=4, rate=0.1, maximum_iteration_count=25,
                                      maximum_warp_update_threshold=0.001, data_term_amplifier=1.0, **kwargs)
        expected = python_runs["hier2d_full/%s/warp" % tag]
        assert np.allclose(result["warp"], expected, atol=1e-5), (tag, np.abs(result["warp"] - expected).max())


# ----------------------------------------------------------------------------- gradient / laplacian
def test_gradient_2d(literals):
    """cpp/tests/test_gradients.cpp:43-122 (scalar_field_gradient_test01/02/05 and the 16x16 header data)"""
    g = lambda n: literals["test_gradients/" + n]
    for case in ("scalar_field_gradient_test01", "scalar_field_gradient_test02"):
        out = oracle.gradient(g(case + "/field"))
        assert np.allclose(out[..., 0], g(case + "/expected_gradient_x"), atol=1e-6)
        assert np.allclose(out[..., 1], g(case + "/expected_gradient_y"), atol=1e-6)
    out = oracle.gradient(g("scalar_field_gradient_test05/field"))
    assert np.allclose(out, g("scalar_field_gradient_test05/expected_gradient"), atol=1e-6)
    field = literals["test_data_gradients/field/field"]
    out = oracle.gradient(field)
    assert np.allclose(out[..., 0], literals["test_data_gradients/expected_gradient_x/expected_gradient_x"], atol=1e-6)
    assert np.allclose(out[..., 1], literals["test_data_gradients/expected_gradient_y/expected_gradient_y"], atol=1e-6)
    gy, gx = np.gradient(field)
    assert np.array_equal(out[..., 0], gx) and np.array_equal(out[..., 1], gy)


def test_gradient_3d(literals):
    """cpp/tests/test_gradients.cpp:167-189: 4x3x2 arange tensor -> constant (1, 4, 12)"""
    field = literals["test_gradients/test_scalar_field_graident_tensor/scalar_field"]
    out = oracle.gradient(field)
    assert np.allclose(out, np.broadcast_to(np.array([1.0, 4.0, 12.0], np.float32), out.shape), atol=1e-6)
    rng = np.random.default_rng(3)
    field = rng.standard_normal((5, 6, 7)).astype(np.float32)
    assert np.array_equal(oracle.gradient(field), np.stack(np.gradient(field), axis=-1))


def test_laplacian_2d(literals):
    """cpp/tests/test_gradients.cpp:240-268"""
    a = np.arange(1, 17, dtype=np.float32)
    a[1::2] *= -1
    a = a.reshape(4, 4)
    b = np.stack([a, a], axis=-1)
    expected = literals["test_gradients/test_laplacian_matrix/expected_b_laplacian_layer"]
    out = oracle.laplacian(b)
    assert np.allclose(out[..., 0], expected, atol=1e-6) and np.allclose(out[..., 1], expected, atol=1e-6)


def test_laplacian_3d(literals):
    """cpp/tests/test_gradients.cpp:270-294 vs test_data_gradients.hpp:350"""
    a = np.zeros((4, 4, 4), dtype=np.float32)
    value, neg = 1.0, True
    for z in range(4):
        neg = not neg
        for y in range(4):
            neg = not neg
            for x in range(4):
                a[x, y, z] = -value if neg else value
                value += 1.0
                neg = not neg
    field = np.stack([a, a, a], axis=-1)
    expected = literals["test_data_gradients/expected_tensorV3_gradient/exp"]
    assert np.allclose(oracle.laplacian(field), expected, atol=1e-6)


def test_laplacian_matches_scipy():
    """reference Python twin: scipy.ndimage.laplace(mode='nearest'), hierarchical_optimizer2d.py:200-202"""
    import scipy.ndimage
    rng = np.random.default_rng(5)
    v = rng.standard_normal((9, 11, 2)).astype(np.float32)
    out = oracle.laplacian(v)
    for c in range(2):
        assert np.allclose(out[..., c], scipy.ndimage.laplace(v[..., c], mode="nearest"), atol=2e-6)


# ----------------------------------------------------------------------------- convolution
def test_convolution_2d(literals):
    """cpp/tests/test_convolution.cpp:36-253"""
    g = lambda n: literals["test_convolution/" + n]
    field = g("test_convolve_with_kernel_preserve_zeros01/field")
    v = np.stack([field, field], axis=-1)
    expected = g("test_convolve_with_kernel_preserve_zeros01/field#1")
    out = oracle.convolve_with_kernel(v, g("test_convolve_with_kernel_preserve_zeros01/kernel"), preserve_zeros=True)
    assert np.allclose(out, np.stack([expected, expected], axis=-1), atol=1e-6)
    case = "test_convolve_with_kernel_preserve_zeros02"
    out = oracle.convolve_with_kernel(g(case + "/vector_field"), g(case + "/kernel"), preserve_zeros=True)
    assert np.allclose(out, g(case + "/expected_output"), atol=1e-6)
    case = "test_convolve_with_kernel_matrix"
    out = oracle.convolve_with_kernel(g(case + "/vector_field"), g(case + "/kernel"))
    assert np.allclose(out, g(case + "/expected_output"), atol=1e-6)


def test_convolution_3d(literals):
    """cpp/tests/test_convolution.cpp:255-262 vs test_data_convolution.hpp:30"""
    v = arange_tensor3v3f((4, 4, 5))
    out = oracle.convolve_with_kernel(v, np.array([3.0, 2.0, 1.0], np.float32))
    expected = literals["test_data_convolution/convolved_3d_vector_field/convolved_3d_vector_field"]
    assert np.allclose(out, expected, atol=1e-6, rtol=1e-6)


def test_convolution_python_reference(python_runs):
    """reference Python math_utils/convolution.py:70-132 (np.convolve based) on seeded fields"""
    k = python_runs["kernel7"]
    assert np.allclose(oracle.convolve_with_kernel(python_runs["conv2d/in"], k), python_runs["conv2d/out"], atol=2e-6)
    assert np.allclose(oracle.convolve_with_kernel(python_runs["conv3d/in"], k), python_runs["conv3d/out"], atol=2e-6)


# ----------------------------------------------------------------------------- resampling / pyramid
def test_upsampling(literals):
    """cpp/tests/test_resampling.cpp:29-269"""
    g = lambda n: literals["test_resampling/" + n]
    case = "test_upsampling_linear_matrix01"
    for suffix in ("", "2", "3"):
        out = oracle.upsample(g(case + "/input" + suffix), 2, linear=True)
        assert np.allclose(out, g(case + "/expected_output" + suffix), atol=1e-6)
    case = "test_upsampling_nearest_tensor01"
    assert np.array_equal(oracle.upsample(g(case + "/input"), 3, linear=False), g(case + "/expected_output"))
    case = "test_upsampling_linear_tensor01"
    for suffix in ("", "2"):
        out = oracle.upsample(g(case + "/input" + suffix), 3, linear=True)
        assert np.allclose(out, g(case + "/expected_output" + suffix), atol=1e-6)
    case = "test_upsampling_linear_tensor02"
    out = oracle.upsample(g(case + "/input"), 3, linear=True)
    assert np.allclose(out, g(case + "/expected_output"), atol=1e-6)


def test_downsampling(literals):
    """cpp/tests/test_resampling.cpp:271-430"""
    g = lambda n: literals["test_resampling/" + n]
    case = "test_downsampling_linear_matrix01"
    for suffix in ("", "2"):
        out = oracle.downsample(g(case + "/input" + suffix), 2, linear=True)
        assert np.allclose(out, g(case + "/expected_output" + suffix), atol=1e-6)
    case = "test_downsampling_linear_tensor01"
    for suffix in ("", "2"):
        out = oracle.downsample(g(case + "/input" + suffix), 3, linear=True)
        assert np.allclose(out, g(case + "/expected_output" + suffix), atol=1e-5)


def test_resampling_python_reference(python_runs):
    """reference Python math_utils/resampling.py:29-125 (3D linear up/down)"""
    field = python_runs["resample3d/in"]
    assert np.allclose(oracle.upsample(field, 3, linear=True), python_runs["resample3d/up_linear"], atol=2e-6)
    assert np.allclose(oracle.downsample(field, 3, linear=True), python_runs["resample3d/down_linear"], atol=2e-6)


def test_pyramid(literals, python_runs):
    """cpp/tests/test_hierarchical_optimizer.cpp:49-147, tests/test_field_pyramid.py:24-86"""
    tile = literals["test_hierarchical_optimizer/pyramid2d_test01/tile"]
    field = np.tile(tile, (16, 16))
    l2 = oracle.downsample(field, 2)
    l1 = oracle.downsample(l2, 2)
    l0 = oracle.downsample(l1, 2)
    assert l0.shape == (16, 16) and l1.shape == (32, 32) and l2.shape == (64, 64)
    assert l2[0, 0] == tile[0:2, 0:2].mean() and l2[1, 0] == tile[2:4, 0:2].mean()
    assert l2[0, 1] == tile[0:2, 2:4].mean() and l2[1, 1] == tile[2:4, 2:4].mean()
    assert l1[1, 1] == 5.0 and l0[0, 0] == 5.0 / 4.0
    # 3D: 8x8x8 arange, Eigen column-major data -> numpy [i,j,k] = data[i + 8 j + 64 k]
    data = literals["test_data_hierarchical_optimizer/pyramid3d_argument_field/data"]
    field3 = data.reshape(8, 8, 8).transpose(2, 1, 0).copy()
    m2 = oracle.downsample(field3, 3)
    m1 = oracle.downsample(m2, 3)
    m0 = oracle.downsample(m1, 3)
    assert m0.shape == (1, 1, 1) and m0[0, 0, 0] == 255.5
    assert m2[0, 0, 0] == 36.5 and m2[-1, -1, -1] == 474.5
    # reference Python ScalarFieldPyramid2d
    level = python_runs["warp2d/field"]
    expected = [python_runs["pyramid2d/level%d" % i] for i in range(4)]
    assert np.array_equal(level, expected[3])
    for i in (2, 1, 0):
        level = oracle.downsample(level, 2)
        assert np.allclose(level, expected[i], atol=1e-6)


def test_max_norm(literals):
    """cpp/tests/test_math.cpp:43-72,106-132"""
    v = literals["test_math/max_norm_test01/vector_field"]
    assert oracle.max_norm(v) == pytest.approx(float(np.sqrt((v.astype(np.float64) ** 2).sum(-1)).max()), rel=1e-6)
    v3 = literals["test_data_math/min_max_vector_field_3d/a"]
    assert oracle.max_norm(v3) == pytest.approx(float(np.linalg.norm(v3, axis=-1).max()), rel=1e-6)


def test_locate_max_norm(literals):
    """the reference's expectations: cpp/tests/test_math.cpp:67-71 (0.7614307 at (1, 2)), :110-113 (1.25495625 at (0, 0)),
    :124-127 (1.5980518 at (0, 6, 8))"""
    norm, at = oracle.locate_max_norm(literals["test_math/max_norm_test01/vector_field"])
    assert norm == pytest.approx(0.7614307, rel=1e-6) and at == (1, 2)
    norm, at = oracle.locate_max_norm(literals["test_data_math/min_max_vector_field_2d/a"])
    assert norm == pytest.approx(1.25495625, rel=1e-6) and at == (0, 0)
    norm, at = oracle.locate_max_norm(literals["test_data_math/min_max_vector_field_3d/a"])
    assert norm == pytest.approx(1.5980518, rel=1e-6) and at == (0, 6, 8)
    # equal maxima: the first one in the reference's column-major traversal stays
    field = np.zeros((4, 4, 2), np.float32)
    field[3, 0] = field[0, 2] = (3.0, 4.0)
    assert oracle.locate_max_norm(field) == (5.0, (0, 3))  # order index 3 (row 3 of column 0) -> x = 3 // 4, y = 3 % 4
    assert oracle.locate_max_norm(np.zeros((4, 4, 2), np.float32)) == (0.0, (0, 0))


# ----------------------------------------------------------------------------- 3D consistency (no reference fixture)
def test_3d_optimizer_reduces_to_2d_planewise():
    """a 3D pair constant along axis 0 must reproduce the 2D result away from that axis'
    influence: with data term only there is no coupling between planes at all."""
    rng = np.random.default_rng(11)
    yy, xx = np.mgrid[0:16, 0:16].astype(np.float32)
    canonical2 = np.clip((np.sqrt((xx - 8) ** 2 + (yy - 7) ** 2) - 4.5) / 4.0, -1, 1).astype(np.float32)
    live2 = np.clip((np.sqrt((xx - 9) ** 2 + (yy - 7.5) ** 2) - 4.8) / 4.0, -1, 1).astype(np.float32)
    canonical3 = np.broadcast_to(canonical2, (16, 16, 16)).copy()
    live3 = np.broadcast_to(live2, (16, 16, 16)).copy()
    kw = dict(tikhonov_term_enabled=False, gradient_kernel_enabled=False, maximum_chunk_size=4, rate=0.1,
              maximum_iteration_count=10, maximum_warp_update_threshold=0.001)
    r3 = oracle.hier_optimize(canonical3, live3, **kw)
    r2 = oracle.hier_optimize(canonical2, live2, **kw)
    assert r3["iterations"] == r2["iterations"]
    # 3D component c <-> axis c; 2D component 0 <-> columns (axis 1 of the plane), 1 <-> rows
    assert np.abs(r3["warp"][..., 0]).max() == 0.0
    for i in (0, 7, 15):
        assert np.allclose(r3["warp"][i, :, :, 2], r2["warp"][..., 0], atol=1e-5)
        assert np.allclose(r3["warp"][i, :, :, 1], r2["warp"][..., 1], atol=1e-5)


def test_tikhonov_strength_0_2_diverges_in_3d_and_0_1_does_not():
    """Why bench.py and the 3D parity tests run tikhonov_strength = 0.1 instead of the reference default 0.2
    (optimizer.hpp:51-65): the hierarchical "Tikhonov" term is the Laplacian of the PREVIOUS gradient
    (optimizer.tpp:195-196), so g_i = data - s * Lap(g_{i-1}) is a recurrence whose high-frequency gain is
    s * 12 * |K(pi)|^3 in 3D (12 = largest eigenvalue magnitude of the 7-point Laplacian, K = the Sobolev kernel's
    response): about 1.6 for s = 0.2, below 1 for s = 0.1. With 0.2 the reference's own algorithm blows up on the
    coarsest level -- the restatement shows it, so the setting cannot be benchmarked or parity-tested meaningfully."""
    from lsf_b200 import synthetic
    canonical, live = synthetic.sphere_plane_pair_3d(32)
    common = dict(tikhonov_term_enabled=True, gradient_kernel_enabled=True, maximum_chunk_size=4, rate=0.1,
                  maximum_iteration_count=100, maximum_warp_update_threshold=0.01, kernel=synthetic.sobolev_kernel_1d())
    stable = oracle.hier_optimize(canonical, live, tikhonov_strength=0.1, **common)
    assert float(np.abs(stable["warp"]).max()) < 5.0 and float(stable["max_updates"].max()) < 1.0
    diverged = oracle.hier_optimize(canonical, live, tikhonov_strength=0.2, **common)
    assert float(diverged["max_updates"][0]) > 1e6
    assert not float(np.nanmax(np.abs(diverged["warp"]))) < 1e6


def test_tikhonov_strength_0p2_diverges_in_3d():
    """Why bench.py and the 3D parity tests run tikhonov_strength=0.1 and not the reference default 0.2: the
    reference feeds the Laplacian of the PREVIOUS gradient back into the gradient (optimizer.tpp:191-200).
    With the 7-tap Sobolev kernel the 3D checkerboard mode is amplified by 12 * strength * |K(pi)|^3 per iteration:
    0.80 at strength 0.1 (decays), 1.59 at 0.2 (grows without bound). The CPU oracle (the restatement of the reference)
    shows exactly that on the synthetic sphere/plane pair: same inputs, same 60 iterations per level."""
    from lsf_b200 import synthetic  # numpy-only module (the package import does not load the CUDA library)
    canonical, live = synthetic.sphere_plane_pair_3d(32)
    results = {}
    for strength in (0.1, 0.2):
        r = oracle.hier_optimize(canonical, live, tikhonov_term_enabled=True, tikhonov_strength=strength,
                                 gradient_kernel_enabled=True, kernel=synthetic.sobolev_kernel_1d(), maximum_chunk_size=4,
                                 maximum_iteration_count=60, maximum_warp_update_threshold=0.0, rate=0.1)
        results[strength] = float(np.abs(r["warp"]).max())
    assert results[0.1] < 5.0, results          # a few voxels of displacement: the pair is shifted by (2.5, -1.5, 1) / 8
    assert results[0.2] > 1e3 * results[0.1], results   # diverged
