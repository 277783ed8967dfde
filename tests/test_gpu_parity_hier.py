"""GPU parity tests (run on the B200 box with `-m gpu`): the CUDA path, called through the C-ABI of
liblsf_b200.so, against the CPU oracle on the same seeded inputs and against the reference's golden vectors.

Tolerances: the kernels restate the reference's float32 operation order without FMA, so primitives and whole
optimizer runs are compared BIT-EXACTLY (np.array_equal) with the oracle; the north-star tolerances
(per-iteration warp <= 1e-5 over the first 10 iterations, final warped live <= 1e-4, identical iteration counts)
are asserted on top, and golden vectors use the reference tests' own atol (1e-6 / 10e-6).
"""
import os

import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu


def random_fields_3d(seed, shape=(12, 10, 14), warp_scale=1.5):
    rng = np.random.default_rng(seed)
    field = (rng.random(shape) * 2 - 1).astype(np.float32)
    vec = rng.standard_normal(shape + (3,)).astype(np.float32)
    warp = (rng.standard_normal(shape + (3,)) * warp_scale).astype(np.float32)
    return field, vec, warp


# ----------------------------------------------------------------------------- primitives, bit-exact vs oracle
@pytest.mark.parametrize("warp_scale", [0.3, 1.5, 8.0, 40.0])
def test_warp_3d(lsf, warp_scale):
    field, vec, warp = random_fields_3d(1, warp_scale=warp_scale)
    assert np.array_equal(lsf.ops.warp(field, warp), oracle.warp(field, warp))
    assert np.array_equal(lsf.ops.warp_with_replacement(vec, warp, 0.0), oracle.warp_with_replacement(vec, warp, 0.0))
    assert np.array_equal(lsf.ops.warp_with_replacement(field, warp, -0.25),
                          oracle.warp_with_replacement(field, warp, -0.25))


def test_warp_3d_identity_and_integer_shift(lsf):
    field, _, _ = random_fields_3d(2)
    zero = np.zeros(field.shape + (3,), np.float32)
    assert np.array_equal(lsf.ops.warp(field, zero), field)
    shift = zero.copy()
    shift[..., 1] = 2.0
    out = lsf.ops.warp(field, shift)
    assert np.array_equal(out[:, :-2, :], field[:, 2:, :])
    assert np.all(out[:, -2:, :] == 1.0)  # out-of-bounds taps -> 1.0 (field_warping.tpp:29-36)


@pytest.mark.parametrize("warp_scale", [0.3, 2.0, 30.0])
def test_warp_2d(lsf, warp_scale):
    rng = np.random.default_rng(3)
    field = (rng.random((17, 13)) * 2 - 1).astype(np.float32)
    vec = rng.standard_normal((17, 13, 2)).astype(np.float32)
    warp = (rng.standard_normal((17, 13, 2)) * warp_scale).astype(np.float32)
    assert np.array_equal(lsf.ops.warp(field, warp), oracle.warp(field, warp))
    assert np.array_equal(lsf.ops.warp_with_replacement(vec, warp, 0.0), oracle.warp_with_replacement(vec, warp, 0.0))


def test_warp_golden(lsf, literals):
    """reference cpp/tests/test_hierarchical_optimizer.cpp:149-159"""
    g = lambda n: literals["py_hierarchical/" + n]
    out = lsf.ops.warp(g("field_A_16x16"), g("warp_field_A_16x16"))
    assert np.allclose(out, g("fA_resampled_with_wfA"), atol=1e-6)
    out = lsf.ops.warp_with_replacement(g("field_B_16x16"), g("warp_field_B_16x16"), 0.0)
    assert np.allclose(out, g("fB_resampled_with_wfB_replacement"), atol=1e-6)


def test_gradient_laplacian(lsf, literals):
    field, vec, _ = random_fields_3d(4)
    assert np.array_equal(lsf.ops.gradient(field), oracle.gradient(field))
    assert np.array_equal(lsf.ops.laplacian(vec), oracle.laplacian(vec))
    f2 = field[0]
    v2 = vec[0, :, :, :2].copy()
    assert np.array_equal(lsf.ops.gradient(f2), oracle.gradient(f2))
    assert np.array_equal(lsf.ops.laplacian(v2), oracle.laplacian(v2))
    # golden: cpp/tests/test_gradients.cpp + data/test_data_gradients.hpp
    gold = literals["test_data_gradients/field/field"]
    out = lsf.ops.gradient(gold)
    assert np.allclose(out[..., 0], literals["test_data_gradients/expected_gradient_x/expected_gradient_x"], atol=1e-6)
    assert np.allclose(out[..., 1], literals["test_data_gradients/expected_gradient_y/expected_gradient_y"], atol=1e-6)


@pytest.mark.parametrize("taps", [3, 7, 11])
def test_convolution(lsf, taps):
    rng = np.random.default_rng(5)
    kernel = rng.random(taps).astype(np.float32)
    _, vec, _ = random_fields_3d(6, shape=(9, 12, 40))
    assert np.array_equal(lsf.ops.convolve_with_kernel(vec, kernel), oracle.convolve_with_kernel(vec, kernel))
    v2 = rng.standard_normal((21, 35, 2)).astype(np.float32)
    assert np.array_equal(lsf.ops.convolve_with_kernel(v2, kernel), oracle.convolve_with_kernel(v2, kernel))
    v2[rng.random((21, 35)) < 0.4] = 0.0
    assert np.array_equal(lsf.ops.convolve_with_kernel_preserve_zeros(v2, kernel),
                          oracle.convolve_with_kernel(v2, kernel, preserve_zeros=True))


def test_convolution_golden(lsf, literals):
    """reference cpp/tests/test_convolution.cpp:208-262"""
    g = lambda n: literals["test_convolution/" + n]
    case = "test_convolve_with_kernel_matrix"
    out = lsf.ops.convolve_with_kernel(g(case + "/vector_field"), g(case + "/kernel"))
    assert np.allclose(out, g(case + "/expected_output"), atol=1e-6)
    case = "test_convolve_with_kernel_preserve_zeros02"
    out = lsf.ops.convolve_with_kernel_preserve_zeros(g(case + "/vector_field"), g(case + "/kernel"))
    assert np.allclose(out, g(case + "/expected_output"), atol=1e-6)
    X, Y, Z = 4, 4, 5
    x, y, z = np.meshgrid(np.arange(X), np.arange(Y), np.arange(Z), indexing="ij")
    base = 1.0 + 3.0 * (x + X * y + X * Y * z)
    v = np.stack([base, base + 1, base + 2], axis=-1).astype(np.float32)
    out = lsf.ops.convolve_with_kernel(v, np.array([3.0, 2.0, 1.0], np.float32))
    expected = literals["test_data_convolution/convolved_3d_vector_field/convolved_3d_vector_field"]
    assert np.allclose(out, expected, atol=1e-6, rtol=1e-6)


@pytest.mark.parametrize("linear", [False, True])
def test_resampling(lsf, linear):
    rng = np.random.default_rng(7)
    s3 = rng.standard_normal((8, 12, 16)).astype(np.float32)
    v3 = rng.standard_normal((8, 12, 16, 3)).astype(np.float32)
    s2 = rng.standard_normal((16, 32)).astype(np.float32)
    v2 = rng.standard_normal((16, 32, 2)).astype(np.float32)
    for field, nd in ((s3, 3), (v3, 3), (s2, 2), (v2, 2)):
        assert np.array_equal(lsf.ops.downsample(field, nd, linear), oracle.downsample(field, nd, linear))
        assert np.array_equal(lsf.ops.upsample(field, nd, linear), oracle.upsample(field, nd, linear))


def test_resampling_golden(lsf, literals):
    """reference cpp/tests/test_resampling.cpp"""
    g = lambda n: literals["test_resampling/" + n]
    for suffix in ("", "2", "3"):
        out = lsf.ops.upsample(g("test_upsampling_linear_matrix01/input" + suffix), 2, linear=True)
        assert np.allclose(out, g("test_upsampling_linear_matrix01/expected_output" + suffix), atol=1e-6)
    out = lsf.ops.upsample(g("test_upsampling_linear_tensor02/input"), 3, linear=True)
    assert np.allclose(out, g("test_upsampling_linear_tensor02/expected_output"), atol=1e-6)
    for suffix in ("", "2"):
        out = lsf.ops.downsample(g("test_downsampling_linear_matrix01/input" + suffix), 2, linear=True)
        assert np.allclose(out, g("test_downsampling_linear_matrix01/expected_output" + suffix), atol=1e-6)
        out = lsf.ops.downsample(g("test_downsampling_linear_tensor01/input" + suffix), 3, linear=True)
        assert np.allclose(out, g("test_downsampling_linear_tensor01/expected_output" + suffix), atol=1e-5)


def test_max_norm(lsf):
    _, vec, _ = random_fields_3d(8)
    assert lsf.ops.max_norm(vec) == oracle.max_norm(vec)


def test_locate_max_norm(lsf, literals):
    """reference math::locate_max_norm (statistics.tpp:57-100): value and location, the reference's test literals
    (cpp/tests/test_math.cpp:67-71,110-113,124-127), ties, all-zero fields, host arrays and device tensors"""
    import torch
    for key, expected in (("test_math/max_norm_test01/vector_field", (1, 2)), ("test_data_math/min_max_vector_field_2d/a", (0, 0)),
                          ("test_data_math/min_max_vector_field_3d/a", (0, 6, 8))):
        assert lsf.ops.locate_max_norm(literals[key]) == oracle.locate_max_norm(literals[key])
        assert lsf.ops.locate_max_norm(literals[key])[1] == expected
    rng = np.random.default_rng(11)
    for shape in ((64, 64, 2), (33, 47, 2), (16, 20, 24, 3), (128, 128, 128, 3)):
        field = rng.standard_normal(shape).astype(np.float32)
        assert lsf.ops.locate_max_norm(field) == oracle.locate_max_norm(field)
        assert lsf.ops.locate_max_norm(torch.from_numpy(field).cuda()) == oracle.locate_max_norm(field)
        assert lsf.ops.locate_max_norm(field)[0] == lsf.ops.max_norm(field)
        # equal maxima
        flat = field.reshape(-1, shape[-1])
        for i in rng.integers(0, flat.shape[0], 5):
            flat[i] = 100.0
        assert lsf.ops.locate_max_norm(field) == oracle.locate_max_norm(field)
    assert lsf.ops.locate_max_norm(np.zeros((8, 8, 2), np.float32)) == (0.0, (0, 0))


# ----------------------------------------------------------------------------- hierarchical optimizer
def test_hier2d_golden(lsf, literals):
    """reference cpp/tests/test_hierarchical_optimizer.cpp:161-205, tests/test_hierarchical_optimizer2d.py:39-101"""
    g = lambda n: literals["py_hierarchical/" + n]
    optimizer = lsf.HierarchicalOptimizer2d(tikhonov_term_enabled=False, gradient_kernel_enabled=False,
                                            maximum_chunk_size=8, rate=0.2, maximum_iteration_count=100,
                                            maximum_warp_update_threshold=0.001, data_term_amplifier=1.0)
    warp = optimizer.optimize(g("canonical_field"), g("live_field"), capture_level=3, capture_iterations=100)
    assert np.allclose(warp, g("warp_field"), atol=10e-6)
    final_live = lsf.ops.warp(g("live_field"), warp)
    assert np.allclose(final_live, g("final_live_field"), atol=10e-6)
    assert optimizer.get_per_level_iteration_counts() == [1, 100, 100, 100]
    captured = optimizer.get_captured_warps()
    assert np.allclose(captured[50], g("iteration50_warp_field"), atol=1e-6)


HIER_MODES = {
    "data_only": dict(tikhonov_term_enabled=False, gradient_kernel_enabled=False),
    "tikhonov": dict(tikhonov_term_enabled=True, gradient_kernel_enabled=False, tikhonov_strength=0.05),
    "kernel": dict(tikhonov_term_enabled=False, gradient_kernel_enabled=True),
    "tikhonov_kernel": dict(tikhonov_term_enabled=True, gradient_kernel_enabled=True, tikhonov_strength=0.1),
}


def check_against_oracle(lsf, canonical, live, nd, mode, linear=False, chunk=8, max_iterations=30, threshold=0.01):
    from lsf_b200 import synthetic
    kwargs = dict(HIER_MODES[mode])
    kwargs.update(maximum_chunk_size=chunk, rate=0.1, maximum_iteration_count=max_iterations,
                  maximum_warp_update_threshold=threshold, data_term_amplifier=1.0,
                  kernel=synthetic.sobolev_kernel_1d(), resampling_strategy=int(linear))
    level_count = int(np.log2(chunk)) + 1
    expected = oracle.hier_optimize(canonical, live, dump_level=level_count - 1, dump_iterations=10, **kwargs)
    cls = lsf.HierarchicalOptimizer2d if nd == 2 else lsf.HierarchicalOptimizer3d
    optimizer = cls(**kwargs)
    warp = optimizer.optimize(canonical, live, capture_level=level_count - 1, capture_iterations=10)
    # identical iteration counts (north star)
    assert optimizer.get_per_level_iteration_counts() == expected["iterations"]
    # per-iteration warp fields over the first 10 iterations of the finest level: <= 1e-5 (north star); in fact equal
    captured = optimizer.get_captured_warps()
    assert len(captured) == len(expected["dump"])
    assert np.abs(captured - expected["dump"]).max() <= 1e-5
    assert np.array_equal(captured, expected["dump"])
    assert np.array_equal(warp, expected["warp"])
    # final warped live field <= 1e-4 (north star)
    assert np.abs(lsf.ops.warp(live, warp) - oracle.warp(live, expected["warp"])).max() <= 1e-4
    reports = optimizer.get_per_level_convergence_reports()
    assert np.allclose([r.max_update_length for r in reports], expected["max_updates"], rtol=0, atol=0)
    return optimizer, warp


@pytest.mark.parametrize("mode", sorted(HIER_MODES))
def test_hier3d_vs_oracle(lsf, mode):
    from lsf_b200 import synthetic
    canonical, live = synthetic.sphere_plane_pair_3d(64)
    check_against_oracle(lsf, canonical, live, 3, mode)


@pytest.mark.parametrize("mode", ["data_only", "tikhonov_kernel"])
def test_hier3d_linear_strategy_vs_oracle(lsf, mode):
    from lsf_b200 import synthetic
    canonical, live = synthetic.sphere_plane_pair_3d(64)
    check_against_oracle(lsf, canonical, live, 3, mode, linear=True, chunk=4)


def test_hier3d_non_cubic_and_early_termination(lsf):
    """ragged (non-cubic, non power-of-two) volume and a threshold that terminates levels early"""
    from lsf_b200 import synthetic
    canonical, live = synthetic.sphere_plane_pair_3d(64)
    canonical, live = canonical[:48, 8:48, :56].copy(), live[:48, 8:48, :56].copy()
    optimizer, _ = check_against_oracle(lsf, canonical, live, 3, "kernel", chunk=8, max_iterations=40, threshold=0.05)
    counts = optimizer.get_per_level_iteration_counts()
    assert any(c < 40 for c in counts), counts


@pytest.mark.parametrize("mode", sorted(HIER_MODES))
def test_hier2d_vs_oracle(lsf, mode):
    from lsf_b200 import synthetic
    canonical, live = synthetic.circle_line_pair_2d(128)
    check_against_oracle(lsf, canonical, live, 2, mode, chunk=8)
    check_against_oracle(lsf, canonical, live, 2, mode, chunk=8, linear=True)


@pytest.mark.parametrize("mode", sorted(HIER_MODES))
def test_hier2d_single_launch_levels_vs_oracle(lsf, mode):
    """without a per-iteration capture every level of a 2D run takes the single-launch path (csrc/hier2d_persistent.cu: all
    iterations of a level in one cooperative kernel, termination test inside): bit-identical to the oracle and to the
    one-launch-per-kernel path (LSF_HIER2D_PERSISTENT=0), identical iteration counts -- also where levels terminate early,
    on a non-square field and with LINEAR resampling -- with far fewer launches"""
    from lsf_b200 import _lib, synthetic
    canonical, live = synthetic.circle_line_pair_2d(128)
    cases = [(canonical, live, 30, 0.01, 0), (canonical, live, 60, 0.05, 1),
             (canonical[:64, :].copy(), live[:64, :].copy(), 25, 0.02, 0)]
    for c, l, iterations, threshold, linear in cases:
        kwargs = dict(HIER_MODES[mode])
        kwargs.update(maximum_chunk_size=8, rate=0.1, maximum_iteration_count=iterations,
                      maximum_warp_update_threshold=threshold, data_term_amplifier=1.0,
                      kernel=synthetic.sobolev_kernel_1d(), resampling_strategy=linear)
        expected = oracle.hier_optimize(c, l, **kwargs)
        optimizer = lsf.HierarchicalOptimizer2d(**kwargs)
        before = _lib.load().lsf_launch_count()
        warp = optimizer.optimize(c, l)
        launches = _lib.load().lsf_launch_count() - before
        assert optimizer.get_per_level_iteration_counts() == expected["iterations"]
        assert np.array_equal(warp, expected["warp"])
        reports = optimizer.get_per_level_convergence_reports()
        assert np.allclose([r.max_update_length for r in reports], expected["max_updates"], rtol=0, atol=0)
        os.environ["LSF_HIER2D_PERSISTENT"] = "0"
        try:
            before = _lib.load().lsf_launch_count()
            plain = optimizer.optimize(c, l)
            plain_launches = _lib.load().lsf_launch_count() - before
        finally:
            del os.environ["LSF_HIER2D_PERSISTENT"]
        assert np.array_equal(plain, warp)
        assert optimizer.get_per_level_iteration_counts() == expected["iterations"]
        assert launches < plain_launches / 2, (launches, plain_launches)
        # small levels live in the shared memory of one thread-block cluster (k_hier_level2d_strips); LSF_HIER2D_CLUSTER=1
        # keeps them in global memory (k_hier_level2d in one cluster), LSF_HIER2D_CLUSTER=0 runs that kernel as a cooperative grid
        for variant in ("1", "0"):
            os.environ["LSF_HIER2D_CLUSTER"] = variant
            try:
                other = optimizer.optimize(c, l)
            finally:
                del os.environ["LSF_HIER2D_CLUSTER"]
            assert np.array_equal(other, warp)
            assert optimizer.get_per_level_iteration_counts() == expected["iterations"]
        # the levels are enqueued back to back and their maximum slots read once at the end; LSF_HIER2D_DEFER_POLL=0 waits
        # for every level as the one-launch-per-kernel path does
        os.environ["LSF_HIER2D_DEFER_POLL"] = "0"
        try:
            other = optimizer.optimize(c, l)
        finally:
            del os.environ["LSF_HIER2D_DEFER_POLL"]
        assert np.array_equal(other, warp)
        assert optimizer.get_per_level_iteration_counts() == expected["iterations"]
        reports = optimizer.get_per_level_convergence_reports()
        assert np.allclose([r.max_update_length for r in reports], expected["max_updates"], rtol=0, atol=0)


@pytest.mark.parametrize("mode", sorted(HIER_MODES))
def test_hier2d_single_launch_long_warps(lsf, mode):
    """shapes displaced by 14 / 11 pixels: the warps of the finer levels reach beyond the rows of the live pack a block of the
    cluster keeps in its own shared memory (strip + 4 rows), so gather taps are read from other blocks' shared memory; also
    fields whose levels do not divide evenly into strips (96 x 96: 6, 3, 2 rows per block with a 7-tap filter); bit-identical
    to the oracle and to the global-memory path"""
    from lsf_b200 import synthetic
    for size in (128, 96):
        canonical, live = synthetic.circle_line_pair_2d(size, shift=(14.0, -11.0), line_shift=-9.0)
        kwargs = dict(HIER_MODES[mode])
        kwargs.update(maximum_chunk_size=8 if size == 128 else 4, rate=0.5, maximum_iteration_count=60,
                      maximum_warp_update_threshold=0.0005, data_term_amplifier=1.0, kernel=synthetic.sobolev_kernel_1d(),
                      resampling_strategy=1 if size == 96 else 0)
        expected = oracle.hier_optimize(canonical, live, **kwargs)
        optimizer = lsf.HierarchicalOptimizer2d(**kwargs)
        warp = optimizer.optimize(canonical, live)
        assert optimizer.get_per_level_iteration_counts() == expected["iterations"]
        assert np.array_equal(warp, expected["warp"])
        assert float(np.abs(warp).max()) > 6.0  # longer than the pack halo
        os.environ["LSF_HIER2D_CLUSTER"] = "1"
        try:
            other = optimizer.optimize(canonical, live)
        finally:
            del os.environ["LSF_HIER2D_CLUSTER"]
        assert np.array_equal(other, warp)


def test_hier2d_python_reference_runs(lsf, python_runs):
    """outputs of the reference's own Python HierarchicalOptimizer2d (tests/golden/make_golden.py)"""
    canonical, live = python_runs["hier2d_full/canonical"], python_runs["hier2d_full/live"]
    modes = {
        "data_only": dict(tikhonov_term_enabled=False, gradient_kernel_enabled=False),
        "tikhonov": dict(tikhonov_term_enabled=True, gradient_kernel_enabled=False, tikhonov_strength=0.1),
        "tikhonov_kernel": dict(tikhonov_term_enabled=True, gradient_kernel_enabled=True, tikhonov_strength=0.2,
                                kernel=python_runs["kernel7"]),
    }
    for tag, kwargs in modes.items():
        optimizer = lsf.HierarchicalOptimizer2d(maximum_chunk_size=4, rate=0.1, maximum_iteration_count=25,
                                                maximum_warp_update_threshold=0.001, **kwargs)
        warp = optimizer.optimize(canonical, live)
        assert np.allclose(warp, python_runs["hier2d_full/%s/warp" % tag], atol=1e-5)


def test_hier3d_device_tensors_match_host_arrays(lsf):
    import torch
    from lsf_b200 import synthetic
    canonical, live = synthetic.sphere_plane_pair_3d(32)
    optimizer = lsf.HierarchicalOptimizer3d(tikhonov_term_enabled=False, gradient_kernel_enabled=True,
                                            kernel=synthetic.sobolev_kernel_1d(), maximum_iteration_count=12,
                                            maximum_chunk_size=4)
    host = optimizer.optimize(canonical, live)
    device = optimizer.optimize(torch.from_numpy(canonical).cuda(), torch.from_numpy(live).cuda())
    assert device.is_cuda and np.array_equal(device.cpu().numpy(), host)


def test_preconditions_raise(lsf):
    """reference pyramid.tpp:53-60 / resampling.tpp:361 -> RuntimeError"""
    field = np.zeros((16, 16, 16), np.float32)
    with pytest.raises(RuntimeError):
        lsf.HierarchicalOptimizer3d(maximum_chunk_size=6).optimize(field, field)
    with pytest.raises(RuntimeError):
        lsf.HierarchicalOptimizer3d(maximum_chunk_size=32).optimize(field, field)
    with pytest.raises(RuntimeError):
        lsf.HierarchicalOptimizer2d(maximum_chunk_size=4).optimize(np.zeros((12, 12), np.float32),
                                                                   np.zeros((12, 12), np.float32))
    with pytest.raises(ValueError):
        lsf.HierarchicalOptimizer3d().optimize(field, field[:8])


# ----------------------------------------------------------------------------- full-size properties (256^3)
def test_full_size_properties_256(lsf):
    """At BASELINE.json's 256^3 the oracle is too slow for a dense comparison inside a unit test; check
    size-independent properties instead: (1) identical inputs give an exactly zero warp after one iteration
    per level (the data term vanishes), (2) a pair that is constant along axis 2 yields a warp whose axis-2
    component is zero and which is constant along axis 2, (3) a sub-sampled set of voxels of the first
    iterations agrees with the oracle run on the enclosing 64^3 crop when the crop is far from the borders'
    influence -- here simply: determinism, two runs are bit-identical."""
    import torch
    from lsf_b200 import synthetic
    canonical, live = synthetic.sphere_plane_pair_3d(256, xp=torch, device="cuda")
    kernel = synthetic.sobolev_kernel_1d()
    optimizer = lsf.HierarchicalOptimizer3d(tikhonov_term_enabled=True, tikhonov_strength=0.05,
                                            gradient_kernel_enabled=True, kernel=kernel,
                                            maximum_iteration_count=5, maximum_warp_update_threshold=0.01)
    same = optimizer.optimize(canonical, canonical)
    assert float(same.abs().max()) == 0.0
    assert optimizer.get_per_level_iteration_counts() == [1, 1, 1, 1]
    flat_c = canonical[:, :, 128:129].expand(-1, -1, 256).contiguous()
    flat_l = live[:, :, 128:129].expand(-1, -1, 256).contiguous()
    no_kernel = lsf.HierarchicalOptimizer3d(tikhonov_term_enabled=False, gradient_kernel_enabled=False,
                                            maximum_iteration_count=5, maximum_warp_update_threshold=0.01)
    warp = no_kernel.optimize(flat_c, flat_l)
    assert float(warp[..., 2].abs().max()) == 0.0
    assert bool((warp[:, :, :1, :] == warp).all())
    first = optimizer.optimize(canonical, live)
    second = optimizer.optimize(canonical, live)
    assert bool((first == second).all())
    assert float(first.abs().max()) > 0.0


@pytest.mark.parametrize("mode", ["tikhonov", "tikhonov_kernel", "kernel"])
def test_fused_kernels_match_first_generation_kernels(lsf, mode, monkeypatch):
    """A/B: the fused production kernels (4 voxels/thread stage 1, single-kernel separable filter) against the
    straightforward one-voxel-per-thread kernels (LSF_LEGACY_KERNELS=1) -- bit-identical."""
    from lsf_b200 import synthetic
    canonical, live = synthetic.sphere_plane_pair_3d(64)
    canonical, live = canonical[:, :48, :40].copy(), live[:, :48, :40].copy()  # tiles cut by the volume border
    kwargs = dict(HIER_MODES[mode])
    kwargs.update(maximum_chunk_size=4, maximum_iteration_count=12, kernel=synthetic.sobolev_kernel_1d())
    fast = lsf.HierarchicalOptimizer3d(**kwargs).optimize(canonical, live)
    # previous cut of the iteration (stage 1 | three-pass filter kernel) and odd chunk sizes of the split kernels
    monkeypatch.setenv("LSF_SPLIT_X", "0")
    assert np.array_equal(lsf.HierarchicalOptimizer3d(**kwargs).optimize(canonical, live), fast)
    monkeypatch.delenv("LSF_SPLIT_X")
    monkeypatch.setenv("LSF_XCHUNK_A", "7")
    monkeypatch.setenv("LSF_XCHUNK_B", "5")
    assert np.array_equal(lsf.HierarchicalOptimizer3d(**kwargs).optimize(canonical, live), fast)
    monkeypatch.delenv("LSF_XCHUNK_A")
    monkeypatch.delenv("LSF_XCHUNK_B")
    monkeypatch.setenv("LSF_LEGACY_KERNELS", "1")
    legacy = lsf.HierarchicalOptimizer3d(**kwargs).optimize(canonical, live)
    assert np.array_equal(fast, legacy)
    rng = np.random.default_rng(9)
    vec = rng.standard_normal((70, 37, 45, 3)).astype(np.float32)
    for taps in (3, 5, 7):
        kernel = rng.random(taps).astype(np.float32)
        slow = lsf.ops.convolve_with_kernel(vec, kernel)
        monkeypatch.delenv("LSF_LEGACY_KERNELS")
        assert np.array_equal(lsf.ops.convolve_with_kernel(vec, kernel), slow)
        monkeypatch.setenv("LSF_LEGACY_KERNELS", "1")


@pytest.mark.parametrize("mode", ["tikhonov_kernel", "kernel"])
@pytest.mark.parametrize("taps", [3, 5, 7])
def test_tma_generation_matches_previous_generations(lsf, mode, taps, monkeypatch):
    """A/B: the TMA-fed stage 1 + y-marching filter (default) against the second-generation split kernels
    (LSF_TMA=0) and the first-generation kernels, on a volume whose tiles are cut by every face (Y and Z not
    multiples of the 8 x 32 tile), with odd chunk sizes along the marching axes -- bit-identical."""
    from lsf_b200 import synthetic
    canonical, live = synthetic.sphere_plane_pair_3d(96)
    canonical, live = canonical[8:80, 4:84, 10:78].copy(), live[8:80, 4:84, 10:78].copy()  # 72 x 80 x 68
    kwargs = dict(HIER_MODES[mode])
    kwargs.update(maximum_chunk_size=4, maximum_iteration_count=8, kernel=synthetic.sobolev_kernel_1d(taps))
    fast = lsf.HierarchicalOptimizer3d(**kwargs).optimize(canonical, live)
    assert np.abs(fast).max() > 0
    monkeypatch.setenv("LSF_XCHUNK_T", "13")
    monkeypatch.setenv("LSF_YCHUNK_T", "9")
    assert np.array_equal(lsf.HierarchicalOptimizer3d(**kwargs).optimize(canonical, live), fast)
    monkeypatch.delenv("LSF_XCHUNK_T")
    monkeypatch.delenv("LSF_YCHUNK_T")
    # fourth-generation switches one at a time: warp update inside the filter kernel instead of deferred into the next
    # stage 1, one-voxel-per-thread y-marching filter, block barrier per plane in stage 1, L2 prefetch of the pack
    for switches in ({"LSF_DEFER": "0"}, {"LSF_DEFER": "0", "LSF_YMARCH2": "0"}, {"LSF_DEFER": "0", "LSF_DECOUPLE": "0"},
                     {"LSF_L2PF": "1"}):
        for name, value in switches.items():
            monkeypatch.setenv(name, value)
        assert np.array_equal(lsf.HierarchicalOptimizer3d(**kwargs).optimize(canonical, live), fast), switches
        for name in switches:
            monkeypatch.delenv(name)
    monkeypatch.setenv("LSF_TMA", "0")
    assert np.array_equal(lsf.HierarchicalOptimizer3d(**kwargs).optimize(canonical, live), fast)
    monkeypatch.setenv("LSF_LEGACY_KERNELS", "1")
    assert np.array_equal(lsf.HierarchicalOptimizer3d(**kwargs).optimize(canonical, live), fast)


@pytest.mark.parametrize("mode", ["tikhonov_kernel", "kernel"])
@pytest.mark.parametrize("taps", [3, 7])
def test_pair_generation_matches_previous_generations(lsf, mode, taps, monkeypatch):
    """A/B: the two-voxels-per-thread stage 1 (default where Z % 64 == 0) with 64 x 8 and 64 x 4 tiles and odd chunk
    sizes against the third generation (LSF_PAIR_TY=0) and the first-generation kernels, on a volume with two z tiles
    (both z faces and an interior tile edge) and ragged X -- bit-identical."""
    from lsf_b200 import synthetic
    canonical, live = synthetic.sphere_plane_pair_3d(128)
    canonical, live = canonical[20:92, 24:104, :].copy(), live[20:92, 24:104, :].copy()  # 72 x 80 x 128
    kwargs = dict(HIER_MODES[mode])
    kwargs.update(maximum_chunk_size=2, maximum_iteration_count=8, kernel=synthetic.sobolev_kernel_1d(taps))
    fast = lsf.HierarchicalOptimizer3d(**kwargs).optimize(canonical, live)
    assert np.abs(fast).max() > 0
    monkeypatch.setenv("LSF_PAIR_TY", "4")
    monkeypatch.setenv("LSF_XCHUNK_T", "13")
    assert np.array_equal(lsf.HierarchicalOptimizer3d(**kwargs).optimize(canonical, live), fast)
    monkeypatch.setenv("LSF_PAIR_TY", "8")
    monkeypatch.setenv("LSF_XCHUNK_T", "5")
    assert np.array_equal(lsf.HierarchicalOptimizer3d(**kwargs).optimize(canonical, live), fast)
    monkeypatch.delenv("LSF_XCHUNK_T")
    monkeypatch.setenv("LSF_PAIR_TY", "0")
    assert np.array_equal(lsf.HierarchicalOptimizer3d(**kwargs).optimize(canonical, live), fast)
    monkeypatch.setenv("LSF_LEGACY_KERNELS", "1")
    assert np.array_equal(lsf.HierarchicalOptimizer3d(**kwargs).optimize(canonical, live), fast)


@pytest.mark.parametrize("threshold", [0.03, 0.05, 0.08])
def test_deferred_update_with_early_termination(lsf, threshold):
    """Tikhonov + Sobolev kernel runs defer the warp update into the next stage 1 (ping-pong warp buffers); the level
    ends after an odd or an even number of iterations depending on the threshold, and iterations enqueued beyond the
    converged one are skipped on the device. Warp, iteration counts and final update lengths equal the oracle's."""
    from lsf_b200 import synthetic
    canonical, live = synthetic.sphere_plane_pair_3d(64)
    kwargs = dict(HIER_MODES["tikhonov_kernel"])
    kwargs.update(maximum_chunk_size=4, maximum_iteration_count=45, maximum_warp_update_threshold=threshold,
                  kernel=synthetic.sobolev_kernel_1d())
    expected = oracle.hier_optimize(canonical, live, **kwargs)
    optimizer = lsf.HierarchicalOptimizer3d(**kwargs)
    warp = optimizer.optimize(canonical, live)
    counts = optimizer.get_per_level_iteration_counts()
    assert counts == expected["iterations"]
    assert any(c < 45 for c in counts), counts
    assert np.array_equal(warp, expected["warp"])
    reports = optimizer.get_per_level_convergence_reports()
    assert np.allclose([r.max_update_length for r in reports], expected["max_updates"], rtol=0, atol=0)


@pytest.mark.parametrize("mode", ["tikhonov_kernel", "kernel"])
@pytest.mark.parametrize("taps", [3, 5, 7])
def test_ymarch3_filter_matches_previous_generations(lsf, mode, taps, monkeypatch):
    """A/B: the fifth-generation filter kernel k_sobolev_ymarch3 (rows of >= 256 voxels; unrolled row loop, symmetric-tap
    chain, aligned-pair window for the axis-2 operands; full-tap chain LSF_SYM=0) against k_sobolev_ymarch2
    (LSF_YMARCH3=0) and the first-generation kernels, on volumes with one z tile, with two z tiles (halo threads) and
    with a ragged last tile, odd chunk sizes along y -- bit-identical. The non-symmetric kernel exercises the full chain."""
    from lsf_b200 import synthetic
    rng = np.random.default_rng(11)
    for shape in ((40, 40, 256), (26, 24, 512), (26, 20, 328), (26, 24, 128), (26, 20, 200)):
        base_c, base_l = synthetic.sphere_plane_pair_3d(64)
        reps = [int(np.ceil(s / 64)) for s in shape]
        canonical = np.tile(base_c, reps)[:shape[0], :shape[1], :shape[2]].copy()
        live = np.tile(base_l, reps)[:shape[0], :shape[1], :shape[2]].copy()
        for kernel in (synthetic.sobolev_kernel_1d(taps), rng.random(taps).astype(np.float32) * 0.3):
            kwargs = dict(HIER_MODES[mode])
            kwargs.update(maximum_chunk_size=2, maximum_iteration_count=6, kernel=kernel)
            fast = lsf.HierarchicalOptimizer3d(**kwargs).optimize(canonical, live)
            assert np.abs(fast).max() > 0
            assert lsf._lib.load().lsf_debug_last_path() & 32, "k_sobolev_ymarch3 did not run"
            for switches in ({"LSF_SYM": "0"}, {"LSF_YCHUNK_T": "7"}, {"LSF_DEFER": "0"}, {"LSF_XCHUNK_T": "13"},
                             {"LSF_YMARCH3": "0"}, {"LSF_LEGACY_KERNELS": "1"}):
                for name, value in switches.items():
                    monkeypatch.setenv(name, value)
                assert np.array_equal(lsf.HierarchicalOptimizer3d(**kwargs).optimize(canonical, live), fast), (shape, switches)
                for name in switches:
                    monkeypatch.delenv(name)
