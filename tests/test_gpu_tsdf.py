"""GPU parity tests of the TSDF generation from depth images (run on the B200 box with `-m gpu`):
csrc/tsdf.cu, called through lsf_tsdf_generate of the C-ABI behind the reference-shaped `tsdf` scope, against
  * the reference's own goldens (cpp/tests/test_tsdf.cpp:47-333, tests/test_tsdf_ewa.py:40-235) at the reference's
    tolerances, written the way the reference's tests call the generators,
  * the CPU oracle on sub-volumes and on full 128^3 / 256^3 / 512^2 fields of the reference's depth images with a rotated
    camera: filtering NONE BIT-EXACT; the EWA methods within 1e-6 (they differ from the oracle only by the device expf,
    <= 2 ulp per weight; 1e-6 is the tolerance of the reference's own EWA tests).
"""
import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu
INTRINSICS = np.array([[700.0, 0.0, 320.0], [0.0, 700.0, 240.0], [0.0, 0.0, 1.0]], dtype=np.float32)


@pytest.fixture(scope="module")
def cpp(lsf):
    import level_set_fusion_optimization
    return level_set_fusion_optimization


def rotated_pose(angle=0.05, translation=(0.01, -0.02, 0.03)):
    c, s = np.cos(angle), np.sin(angle)
    return np.array([[c, 0, s, translation[0]], [0, 1, 0, translation[1]], [-s, 0, c, translation[2]], [0, 0, 0, 1]],
                    dtype=np.float32)


def generate(cpp, parameters, image, on_device=False):
    """the call sequence of the reference's tests (tests/test_tsdf_ewa.py:62-69)"""
    nd = parameters["nd"]
    p = cpp.tsdf.Parameters2d() if nd == 2 else cpp.tsdf.Parameters3d()
    p.interpolation_method = cpp.tsdf.FilteringMethod(parameters["filtering_method"])
    p.projection_matrix = np.array(parameters["projection_matrix"], dtype=np.float32)
    p.array_offset = (cpp.Vector2i if nd == 2 else cpp.Vector3i)(*parameters["array_offset"])
    p.field_shape = (cpp.Vector2i if nd == 2 else cpp.Vector3i)(*parameters["field_shape"])
    p.smoothing_factor = parameters["smoothing_factor"]
    p.near_clipping_distance = parameters["near_clipping_distance"]
    p.voxel_size = parameters["voxel_size"]
    p.narrow_band_width_voxels = parameters["narrow_band_width_voxels"]
    p.depth_unit_ratio = parameters["depth_unit_ratio"]
    generator = (cpp.tsdf.Generator2d if nd == 2 else cpp.tsdf.Generator3d)(p)
    pose = np.array(parameters["camera_pose"], dtype=np.float32)
    if on_device:
        import torch
        tensor = torch.from_numpy(image.view(np.int16)).cuda()
        return generator.generate(tensor, pose, parameters["image_y_coordinate"]).cpu().numpy()
    return generator.generate(image, pose, parameters["image_y_coordinate"])


def oracle_generate(parameters, image):
    return oracle.tsdf_generate(image, np.array(parameters["camera_pose"], dtype=np.float32), parameters["nd"],
                                parameters["projection_matrix"], parameters["array_offset"], parameters["field_shape"],
                                parameters["image_y_coordinate"], parameters["depth_unit_ratio"],
                                parameters["near_clipping_distance"], parameters["voxel_size"],
                                parameters["narrow_band_width_voxels"], parameters["filtering_method"],
                                parameters["smoothing_factor"])


def make(nd, method, offset, shape, pose=None, y=0, smoothing=1.0, near=0.05):
    return dict(nd=nd, filtering_method=method, array_offset=list(offset), field_shape=list(shape), image_y_coordinate=y,
                camera_pose=(np.eye(4, dtype=np.float32) if pose is None else pose).tolist(), smoothing_factor=smoothing,
                projection_matrix=INTRINSICS.tolist(), depth_unit_ratio=0.001, near_clipping_distance=near, voxel_size=0.004,
                narrow_band_width_voxels=20)


def test_reference_goldens(cpp, tsdf_cases):
    """every TSDF-generation case of the reference's C++ and Python test suites, at the reference's tolerance"""
    for parameters, image, expected in tsdf_cases.cases:
        for on_device in (False, True):
            field = generate(cpp, parameters, image, on_device)
            assert field.dtype == np.float32 and field.shape == expected.shape, parameters["source"]
            assert np.abs(field - expected).max() <= parameters["tolerance"], parameters["source"]


def test_reference_python_runs(cpp, tsdf_cases):
    """runs of the reference's Python generators (see tests/test_oracle_tsdf.py for the half-pixel ties)"""
    for parameters, image, expected in tsdf_cases.python_runs:
        field = generate(cpp, parameters, image)
        off = np.abs(field - expected) > parameters["tolerance"]
        assert off.mean() <= 0.015, parameters["source"]
        if parameters["filtering_method"] != 0:
            assert not off.any(), parameters["source"]


@pytest.mark.parametrize("shape,offset", [((128, 128, 128), (-64, -64, 64)), ((256, 256, 256), (-128, -128, 400)),
                                          ((17, 5, 13), (-46, -8, 105)), ((64, 96, 30), (-30, -40, 690))])
def test_none_3d_bit_exact_vs_oracle(cpp, tsdf_cases, shape, offset):
    """filtering NONE on whole fields (the reference's default 128^3 at offset (-64, -64, 64), experiment/dataset.py:103-118;
    256^3; ragged shapes whose z extent is not a multiple of the four voxels a thread writes), identity and rotated
    camera, both depth images: np.array_equal with the oracle"""
    band_fraction = 0.0
    for key in ("zigzag2_108", "zigzag1_064"):  # surfaces at 0.4 - 1.2 m and at 2.0 - 5.2 m
        image = tsdf_cases.images[key]
        for pose in (None, rotated_pose()):
            parameters = make(3, 0, offset, shape, pose)
            field = generate(cpp, parameters, image)
            expected = oracle_generate(parameters, image)
            assert np.array_equal(field, expected), (key, pose is None)
            band_fraction = max(band_fraction, float((np.abs(field) < 1).mean()))
    assert band_fraction > 0.01  # one of the two frames has its surface inside the field


def test_none_2d_bit_exact_vs_oracle(cpp, tsdf_cases):
    """the reference's 2D experiment geometry: 512 x 512 at offset (-256, 480), image row 200 (build_*_helper defaults),
    and a ragged field"""
    for key, offset in (("zigzag2_108", (-256, 0)), ("zigzag1_064", (-256, 480))):
        image = tsdf_cases.images[key]
        for shape in ((512, 512), (37, 37)):
            for pose in (None, rotated_pose()):
                parameters = make(2, 0, offset, shape, pose, y=200)
                assert np.array_equal(generate(cpp, parameters, image), oracle_generate(parameters, image))
    parameters = make(2, 0, (-256, 0), (512, 512), y=200)
    assert (np.abs(generate(cpp, parameters, tsdf_cases.images["zigzag2_108"])) < 1).mean() > 0.01


@pytest.mark.parametrize("method", [3, 4, 5])
def test_ewa_vs_oracle(cpp, tsdf_cases, method):
    """the three EWA methods, 3D (64 x 8 x 64 around the surface, identity and rotated camera, two covariance scales)
    and 2D (128 x 128), within 1e-6 of the oracle; voxels the generator skips agree exactly (the decisions do not
    involve expf)"""
    image = tsdf_cases.images["zigzag2_108"].copy()
    image[image == 0] = 65535
    for pose in (None, rotated_pose()):
        for smoothing in (1.0, 0.5):
            parameters = make(3, method, (-70, -8, 90), (64, 8, 64), pose, smoothing=smoothing)
            field, expected = generate(cpp, parameters, image), oracle_generate(parameters, image)
            assert np.abs(field - expected).max() <= 1e-6
            assert np.array_equal(field == 1, expected == 1) and (np.abs(field) < 1).mean() > 0.05
    for offset in ((-110, 60), (-256, 0)):
        parameters = make(2, method, offset, (128, 128), y=200)
        field, expected = generate(cpp, parameters, image), oracle_generate(parameters, image)
        assert np.abs(field - expected).max() <= 1e-6
    # image border: a volume whose projection leaves the image on the left (sampling bounds clipped / inclusive samples)
    parameters = make(3, method, (-260, -8, 200), (48, 4, 32))
    field, expected = generate(cpp, parameters, image), oracle_generate(parameters, image)
    assert np.abs(field - expected).max() <= 1e-6


def test_error_behaviour(cpp, tsdf_cases):
    """reference: bilinear methods throw "Not yet implemented" (generator_tensor.tpp:103-123) -> RuntimeError; the
    numpy converter only takes unsigned-short matrices (eigen_numpy_matrix.cpp:79-103)"""
    image = tsdf_cases.images["zigzag2_108"]
    for method in (1, 2):
        with pytest.raises(RuntimeError, match="Not yet implemented"):
            generate(cpp, make(3, method, (-8, -8, 100), (8, 8, 8)), image)
    with pytest.raises(ValueError):
        generate(cpp, make(3, 0, (-8, -8, 100), (8, 8, 8)), image.astype(np.float32))
    with pytest.raises(RuntimeError):
        generate(cpp, make(2, 0, (-8, 100), (8, 8), y=480), image)
    # the generator keeps a copy of the parameters it was built with (generator_crtp.tpp:34-36)
    p = cpp.tsdf.Parameters3d(projection_matrix=INTRINSICS, array_offset=cpp.Vector3i(-8, -8, 105),
                              field_shape=cpp.Vector3i(8, 8, 8))
    generator = cpp.tsdf.Generator3d(p)
    before = generator.generate(image, np.identity(4, dtype=np.float32), 0)
    p.array_offset = cpp.Vector3i(0, 0, 0)
    assert np.array_equal(generator.generate(image, np.identity(4, dtype=np.float32), 0), before)


def test_generated_pair_through_the_optimizer(cpp, lsf, tsdf_cases):
    """the reference's pipeline (experiment/dataset.py:160-170 -> run_hierarchical_optimizer3d.py): canonical and live
    fields generated from two depth frames, then HierarchicalOptimizer3d.optimize -- generated and optimised on the
    device without leaving it, equal to the oracle's generator followed by the oracle's optimizer"""
    import torch
    from lsf_b200 import synthetic
    first, second = tsdf_cases.images["zigzag2_108"], np.roll(tsdf_cases.images["zigzag2_108"], 3, axis=1)
    parameters = make(3, 0, (-46, -16, 105), (32, 32, 32))
    kwargs = dict(tikhonov_term_enabled=True, tikhonov_strength=0.1, gradient_kernel_enabled=True,
                  kernel=synthetic.sobolev_kernel_1d(), maximum_chunk_size=8, maximum_iteration_count=15,
                  maximum_warp_update_threshold=0.01, rate=0.1)
    p = cpp.tsdf.Parameters3d(projection_matrix=INTRINSICS, array_offset=cpp.Vector3i(-46, -16, 105),
                              field_shape=cpp.Vector3i(32, 32, 32))
    generator = cpp.tsdf.Generator3d(p)
    fields = [generator.generate(torch.from_numpy(image.view(np.int16)).cuda()) for image in (first, second)]
    assert all(f.is_cuda for f in fields)
    optimizer = cpp.HierarchicalOptimizer3d(**kwargs)
    warp = optimizer.optimize(fields[0], fields[1]).cpu().numpy()
    canonical, live = oracle_generate(parameters, first), oracle_generate(parameters, second)
    assert np.array_equal(fields[0].cpu().numpy(), canonical) and np.array_equal(fields[1].cpu().numpy(), live)
    expected = oracle.hier_optimize(canonical, live, **kwargs)
    assert optimizer.get_per_level_iteration_counts() == expected["iterations"]
    assert np.array_equal(warp, expected["warp"]) and np.abs(warp).max() > 0.01
