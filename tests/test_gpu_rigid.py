"""GPU parity tests of the rigid SDF-2-SDF tracker (run on the B200 box with `-m gpu`):
csrc/rigid.cu through lsf_sdf2sdf_optimize_2d of the C-ABI behind the reference-shaped `Sdf2SdfOptimizer2d`, against
  * the reference's Python tracker (tests/golden/reference_rigid.npz) with the reference's own assertion -- the call
    sequence and the 1e-4 tolerance of tests/test_sdf_2_sdf_optimizer.py:81-166,
  * the CPU oracle, twist after EVERY iteration: within 2e-6 of the oracle's `double_sums` mode (float32 per-voxel terms
    added up in double -- the device reduction's arithmetic; the live field of every iteration is bit-identical) and within
    1e-4 of its reference mode (sequential float32 sums like Eigen's; the 3 x 3 system is ill-conditioned, the two modes
    differ by ~4e-5 on the reference's test case, and the reference's float64 Python tracker sits closer to the double sums).
"""
import json
import os

import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_rigid.npz")


@pytest.fixture(scope="module")
def sdf2sdfo_cpp(lsf):
    import level_set_fusion_optimization
    return level_set_fusion_optimization


@pytest.fixture(scope="module")
def rigid_runs():
    data = np.load(GOLDEN)
    runs = []
    while "run/%02d/parameters" % len(runs) in data.files:
        k = len(runs)
        runs.append((json.loads(str(data["run/%02d/parameters" % k])), data["run/%02d/canonical_field" % k],
                     data["run/%02d/twist" % k], data["run/%02d/twist_matrix" % k]))
    return data["image/canonical"], data["image/live"], runs


def make_optimizer(sdf2sdfo_cpp, parameters, verbosity_parameters=None, **overrides):
    """the construction of tests/test_sdf_2_sdf_optimizer.py:87-132 / build_sdf_2_sdf_optimizer_helper.py"""
    offset, field_size = parameters["offset"], parameters["field_size"]
    tsdf_generation_parameters = sdf2sdfo_cpp.tsdf.Parameters2d(
        depth_unit_ratio=0.001,
        projection_matrix=np.array(parameters["projection_matrix"], dtype=np.float32),
        near_clipping_distance=0.05,
        array_offset=sdf2sdfo_cpp.Vector2i(int(offset[0]), int(offset[2])),
        field_shape=sdf2sdfo_cpp.Vector2i(field_size, field_size),
        voxel_size=0.004,
        narrow_band_width_voxels=parameters["narrow_band_width_voxels"],
        interpolation_method=sdf2sdfo_cpp.tsdf.FilteringMethod.NONE)
    arguments = dict(rate=parameters["rate"], maximum_iteration_count=parameters["iterations"],
                     tsdf_generation_parameters=tsdf_generation_parameters)
    arguments.update(overrides)
    if verbosity_parameters is not None:
        arguments["verbosity_parameters"] = verbosity_parameters
    return sdf2sdfo_cpp.Sdf2SdfOptimizer2d(**arguments)


def oracle_track(parameters, canonical_field, live_image, **overrides):
    n, offset = parameters["field_size"], parameters["offset"]
    arguments = dict(rate=parameters["rate"], maximum_iteration_count=parameters["iterations"], eta=parameters["eta"],
                     narrow_band_width_voxels=parameters["narrow_band_width_voxels"])
    arguments.update(overrides)
    return oracle.sdf2sdf_optimize(canonical_field, live_image, parameters["image_y_coordinate"],
                                   parameters["projection_matrix"], [offset[0], offset[2]], [n, n], **arguments)


def test_reference_test_operation_same_cpp_to_py(sdf2sdfo_cpp, rigid_runs, capsys):
    """reference tests/test_sdf_2_sdf_optimizer.py:81-166 with the Python tracker's result from the fixture"""
    _, live_depth_image, runs = rigid_runs
    for parameters, canonical_field, _, twist_matrix_py in runs:
        verbosity_parameters_cpp = sdf2sdfo_cpp.Sdf2SdfOptimizer2d.VerbosityParameters(True, True)
        optimizer_cpp = make_optimizer(sdf2sdfo_cpp, parameters, verbosity_parameters_cpp)
        twist_cpp = optimizer_cpp.optimize(image_y_coordinate=parameters["image_y_coordinate"],
                                           canonical_field=canonical_field,
                                           live_depth_image=live_depth_image,
                                           eta=parameters["eta"],
                                           initial_camera_pose=np.eye(4, dtype=np.float32))
        assert twist_cpp.shape == (3, 3) and twist_cpp.dtype == np.float32
        assert np.allclose(twist_cpp, twist_matrix_py, atol=1e-4), parameters["source"]
        printed = capsys.readouterr().out
        assert printed.count("COMPLETED]") == parameters["iterations"] and " [energy: " in printed and " [twist:" in printed


def test_every_iteration_matches_the_oracle(sdf2sdfo_cpp, rigid_runs):
    _, live_depth_image, runs = rigid_runs
    for parameters, canonical_field, _, _ in runs:
        for overrides in ({}, dict(rate=1.0, maximum_iteration_count=5), dict(rate=0.25, maximum_iteration_count=30)):
            expected = oracle_track(parameters, canonical_field, live_depth_image, double_sums=True, **overrides)
            reference_sums = oracle_track(parameters, canonical_field, live_depth_image, **overrides)
            optimizer = make_optimizer(sdf2sdfo_cpp, parameters, **overrides)
            matrix = optimizer.optimize(parameters["image_y_coordinate"], canonical_field, live_depth_image, parameters["eta"])
            assert np.abs(optimizer.get_per_iteration_twists() - expected["twists"]).max() <= 2e-6, overrides
            assert np.abs(matrix - expected["twist_matrix"]).max() <= 2e-6, overrides
            assert np.allclose(optimizer.get_per_iteration_energies(), expected["energies"], rtol=1e-6), overrides
            assert np.abs(optimizer.get_per_iteration_twists() - reference_sums["twists"]).max() <= 1e-4, overrides
            assert np.allclose(optimizer.get_per_iteration_energies(), reference_sums["energies"], rtol=5e-3), overrides


def test_larger_field_ewa_and_device_tensors(sdf2sdfo_cpp, rigid_runs):
    """a 256 x 256 field with a 20-voxel band: numpy arguments == CUDA tensors bit for bit (same kernels, same stream
    order), both within 5e-6 of the oracle (double sums); the EWA-filtered generator drives the tracker as well"""
    import torch
    canonical_image, live_depth_image, runs = rigid_runs
    base = dict(runs[0][0], field_size=256, offset=[-128, -128, 0], narrow_band_width_voxels=20, iterations=10)
    for filtering_method in (0, 3):
        canonical_field = oracle.tsdf_generate(canonical_image, np.eye(4, dtype=np.float32), 2, base["projection_matrix"],
                                               [-128, 0], [256, 256], base["image_y_coordinate"],
                                               narrow_band_width_voxels=20, filtering_method=filtering_method)
        optimizer = make_optimizer(sdf2sdfo_cpp, base)
        optimizer._tsdf_generator.parameters.interpolation_method = sdf2sdfo_cpp.tsdf.FilteringMethod(filtering_method)
        from_numpy = optimizer.optimize(base["image_y_coordinate"], canonical_field, live_depth_image, base["eta"])
        twists = optimizer.get_per_iteration_twists().copy()
        from_tensors = optimizer.optimize(base["image_y_coordinate"], torch.from_numpy(canonical_field).cuda(),
                                          torch.from_numpy(live_depth_image.view(np.int16)).cuda(), base["eta"])
        assert np.array_equal(from_numpy, from_tensors)
        assert np.array_equal(twists, optimizer.get_per_iteration_twists())
        expected = oracle_track(base, canonical_field, live_depth_image, filtering_method=filtering_method, double_sums=True)
        assert np.abs(twists - expected["twists"]).max() <= (5e-6 if filtering_method == 0 else 5e-5), filtering_method
        assert np.abs(twists[-1]).max() > 1e-3


def test_argument_checks_and_trivial_cases(sdf2sdfo_cpp, rigid_runs):
    _, live_depth_image, runs = rigid_runs
    parameters, canonical_field, _, _ = runs[0]
    identity = np.eye(3, dtype=np.float32)
    assert np.array_equal(make_optimizer(sdf2sdfo_cpp, parameters, maximum_iteration_count=0).optimize(
        parameters["image_y_coordinate"], canonical_field, live_depth_image), identity)
    assert np.array_equal(make_optimizer(sdf2sdfo_cpp, parameters, rate=0.0).optimize(
        parameters["image_y_coordinate"], canonical_field, live_depth_image), identity)
    optimizer = make_optimizer(sdf2sdfo_cpp, parameters)
    with pytest.raises(ValueError):
        optimizer.optimize(parameters["image_y_coordinate"], canonical_field[:-1], live_depth_image)
    with pytest.raises(ValueError):
        optimizer.optimize(parameters["image_y_coordinate"], canonical_field, live_depth_image.astype(np.float32))
    default = sdf2sdfo_cpp.Sdf2SdfOptimizer2d()  # sdf_2_sdf_optimizer2d.hpp:33-36 defaults
    assert default.rate == 0.5 and default.maximum_iteration_count == 60
    assert not default.verbosity_parameters.print_per_iteration_info
