"""CPU tests: the oracle's restatement of the TSDF generators (oracle/lsf_oracle_tsdf.cpp) against every golden the
reference holds for them and against runs of the reference's Python generators (tests/golden/reference_tsdf.npz)."""
import numpy as np
import pytest

import oracle


def oracle_generate(parameters, image):
    return oracle.tsdf_generate(image, np.array(parameters["camera_pose"], dtype=np.float32), parameters["nd"],
                                parameters["projection_matrix"], parameters["array_offset"], parameters["field_shape"],
                                parameters["image_y_coordinate"], parameters["depth_unit_ratio"],
                                parameters["near_clipping_distance"], parameters["voxel_size"],
                                parameters["narrow_band_width_voxels"], parameters["filtering_method"],
                                parameters["smoothing_factor"])


def test_reference_test_cases(tsdf_cases):
    """cpp/tests/test_tsdf.cpp:47-333 and the C++ halves of tests/test_tsdf_ewa.py:40-235, each at the tolerance the
    reference's own assertion states (1e-6, two cases 1e-5)"""
    assert len(tsdf_cases.cases) == 13
    methods = set()
    for parameters, image, expected in tsdf_cases.cases:
        field = oracle_generate(parameters, image)
        assert field.shape == expected.shape, parameters["source"]
        assert np.abs(field - expected).max() <= parameters["tolerance"], parameters["source"]
        assert (np.abs(expected) < 1).any(), parameters["source"]  # the golden crosses the surface
        methods.add((parameters["nd"], parameters["filtering_method"]))
    assert methods == {(2, 0), (2, 3), (2, 4), (2, 5), (3, 3)}


def test_reference_python_generators(tsdf_cases):
    """tsdf/generation.py:130-217 (2D), :356-437 (3D), tsdf/ewa.py:59-185 run here on sub-volumes of the reference's depth
    images with an identity and a rotated + translated camera. The Python twins evaluate the pixel coordinate as
    fx * x / z + cx + 0.5 in float64 where the C++ generator divides the float32 projective product by z: voxels whose
    projection is EXACTLY a half pixel (rational voxel coordinates: up to 1.3 % of a sub-volume in front of an
    axis-aligned camera, verified to be ties to 3e-14 pixels when the fixture was made) may round to the neighbouring
    pixel. They are counted and bounded; every other voxel agrees within 2e-5."""
    assert len(tsdf_cases.python_runs) == 10
    informative = 0
    for parameters, image, expected in tsdf_cases.python_runs:
        field = oracle_generate(parameters, image)
        assert field.shape == expected.shape
        off = np.abs(field - expected) > parameters["tolerance"]
        assert off.mean() <= 0.015, (parameters["source"], off.mean())
        if parameters["filtering_method"] != 0:
            assert not off.any(), parameters["source"]
        informative += bool((np.abs(expected) < 1).any())
    assert informative >= 8


def test_unsupported_methods(tsdf_cases):
    """bilinear filtering is "Not yet implemented" in the reference (generator_tensor.tpp:103-123)"""
    parameters, image, _ = tsdf_cases.cases[0]
    for method in (1, 2, 7):
        with pytest.raises(RuntimeError):
            oracle_generate(dict(parameters, filtering_method=method), image)


def test_defaults_and_clipping():
    """voxels at or behind the near clipping distance, outside the image, or looking at a pixel without a reading keep 1;
    3D NONE truncates to -1 / +1 outside the narrow band (common.hpp:32-40)"""
    image = np.full((48, 64), 1000, dtype=np.uint16)  # a wall at 1 m
    image[:, :8] = 0
    projection = [[50.0, 0, 32.0], [0, 50.0, 24.0], [0, 0, 1]]
    field = oracle.tsdf_generate(image, np.eye(4, dtype=np.float32), 3, projection, (-16, -16, -8), (32, 32, 64),
                                 voxel_size=0.02, narrow_band_width_voxels=10)
    z = (np.arange(64) - 8) * np.float32(0.02)
    assert np.all(field[:, :, z <= 0.05] == 1)                       # near clipping
    column = field[16, 16]                                          # the optical axis
    assert np.all(column[z > 1.1 + 1e-6] == -1) and np.all(column[(z > 0.05) & (z < 0.9 - 1e-6)] == 1)
    inside = (z > 0.9) & (z < 1.1)
    assert np.allclose(column[inside], (1.0 - z[inside]) / 0.1, atol=1e-5)
    # pixels without a reading (columns 0..7 of the image) leave the default
    x_voxel = (np.arange(32) - 16) * np.float32(0.02)
    image_x = 50.0 * x_voxel / 1.0 + 32.0
    plane = field[:, 16, 8 + 50]  # z = 1.0
    assert np.all(plane[np.round(image_x) < 8] == 1) and np.all(np.abs(plane[np.round(image_x) >= 9]) < 1e-5)
