"""CPU tests: the oracle's restatement of the rigid SDF-2-SDF tracker (oracle/lsf_oracle_rigid.cpp)
against runs of the reference's Python tracker (tests/golden/reference_rigid.npz, made by tests/golden/make_rigid_golden.py).
The reference holds no golden for its C++ tracker; its own test asserts C++ == Python within 1e-4 on the twist matrix
(tests/test_sdf_2_sdf_optimizer.py:81-166) -- the same assertion pins the oracle."""
import json
import os

import numpy as np
import pytest

import oracle

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_rigid.npz")


@pytest.fixture(scope="module")
def rigid_runs():
    data = np.load(GOLDEN)
    runs = []
    while "run/%02d/parameters" % len(runs) in data.files:
        k = len(runs)
        runs.append((json.loads(str(data["run/%02d/parameters" % k])), data["run/%02d/canonical_field" % k],
                     data["run/%02d/twist" % k], data["run/%02d/twist_matrix" % k]))
    return data["image/canonical"], data["image/live"], runs


def oracle_track(parameters, canonical_field, live_image, **overrides):
    n, offset = parameters["field_size"], parameters["offset"]
    arguments = dict(rate=parameters["rate"], maximum_iteration_count=parameters["iterations"], eta=parameters["eta"],
                     narrow_band_width_voxels=parameters["narrow_band_width_voxels"])
    arguments.update(overrides)
    return oracle.sdf2sdf_optimize(canonical_field, live_image, parameters["image_y_coordinate"],
                                   parameters["projection_matrix"], [offset[0], offset[2]], [n, n], **arguments)


def test_canonical_field_of_the_fixture_is_the_oracle_tsdf(rigid_runs):
    """the canonical field the reference's test builds with its Python generator (tsdf/generation.py:130-217) equals the
    oracle's generator on the canonical frame (half-pixel ties aside, see tests/test_oracle_tsdf.py)"""
    canonical_image, _, runs = rigid_runs
    for parameters, canonical_field, _, _ in runs:
        n, offset = parameters["field_size"], parameters["offset"]
        field = oracle.tsdf_generate(canonical_image, np.eye(4, dtype=np.float32), 2, parameters["projection_matrix"],
                                     [offset[0], offset[2]], [n, n], parameters["image_y_coordinate"],
                                     narrow_band_width_voxels=parameters["narrow_band_width_voxels"])
        assert (np.abs(field - canonical_field) > 2e-5).mean() <= 0.015, parameters["source"]


def test_twist_matches_reference_python_tracker(rigid_runs):
    """tests/test_sdf_2_sdf_optimizer.py:166: np.allclose(twist_cpp, twist_vector_to_matrix2d(twist_py), atol=1e-4)"""
    _, live_image, runs = rigid_runs
    assert len(runs) == 2
    for parameters, canonical_field, twist, twist_matrix in runs:
        result = oracle_track(parameters, canonical_field, live_image)
        assert result["twist_matrix"].shape == (3, 3) and result["twist_matrix"].dtype == np.float32
        assert np.allclose(result["twist_matrix"], twist_matrix, atol=parameters["tolerance"]), parameters["source"]
        assert np.allclose(result["twists"][-1], twist, atol=parameters["tolerance"]), parameters["source"]
        assert np.abs(twist).max() > 1e-3  # the fixture is not the trivial solution
        assert np.array_equal(result["twist_matrix"][2], [0, 0, 1])


def test_twist_matrix_is_a_rigid_motion_and_energy_is_reported(rigid_runs):
    _, live_image, runs = rigid_runs
    parameters, canonical_field, _, _ = runs[0]
    result = oracle_track(parameters, canonical_field, live_image)
    rotation = result["twist_matrix"][:2, :2].astype(np.float64)
    assert np.allclose(rotation @ rotation.T, np.eye(2), atol=1e-6)
    assert abs(np.linalg.det(rotation) - 1) < 1e-6
    assert np.all(np.isfinite(result["energies"])) and result["energies"][-1] < result["energies"][0]


def test_zero_iterations_and_rate_zero_give_identity(rigid_runs):
    """sdf_2_sdf_optimizer2d.cpp:69,106: the twist starts at zero and moves by rate * (optimal - twist)"""
    _, live_image, runs = rigid_runs
    parameters, canonical_field, _, _ = runs[0]
    assert np.array_equal(oracle_track(parameters, canonical_field, live_image, maximum_iteration_count=0)["twist_matrix"],
                          np.eye(3, dtype=np.float32))
    assert np.array_equal(oracle_track(parameters, canonical_field, live_image, rate=0.0)["twist_matrix"],
                          np.eye(3, dtype=np.float32))


def test_first_iteration_is_the_normal_equation_solution(rigid_runs):
    """one iteration at rate 1 from the zero twist == solve(sum g g^T, sum (canonical - live) g) with g = (gradient of the
    live field, cross term) / voxel_size (sdf_gradient_wrt_transformation2d.cpp:18-50), evaluated here in float64 numpy"""
    _, live_image, runs = rigid_runs
    parameters, canonical_field, _, _ = runs[1]
    n, offset, voxel_size = parameters["field_size"], parameters["offset"], 0.004
    live = oracle.tsdf_generate(live_image, np.eye(4, dtype=np.float32), 2, parameters["projection_matrix"],
                                [offset[0], offset[2]], [n, n], parameters["image_y_coordinate"],
                                narrow_band_width_voxels=parameters["narrow_band_width_voxels"]).astype(np.float64)
    g_rows, g_columns = np.gradient(live)
    x = (np.arange(n) + offset[0])[None, :] * voxel_size * np.ones((n, 1))
    z = (np.arange(n) + offset[2])[:, None] * voxel_size * np.ones((1, n))
    g = np.stack([g_columns, g_rows, g_columns * z - g_rows * x], axis=-1).reshape(-1, 3) / voxel_size
    A = g.T @ g
    b = g.T @ (canonical_field.astype(np.float64) - live).reshape(-1)
    expected = np.linalg.solve(A, b)
    result = oracle_track(parameters, canonical_field, live_image, rate=1.0, maximum_iteration_count=1)
    assert np.allclose(result["twists"][0], expected, rtol=2e-3, atol=2e-5)


def test_double_sums_mode_bounds_the_accumulation_error(rigid_runs):
    """the oracle's `double_sums` mode (the GPU reduction's arithmetic) stays within the reference's 1e-4 of the float32
    sums and of the reference's Python tracker (float64 numpy), and is the closer of the two to the latter"""
    _, live_image, runs = rigid_runs
    for parameters, canonical_field, twist, twist_matrix in runs:
        single = oracle_track(parameters, canonical_field, live_image)
        double = oracle_track(parameters, canonical_field, live_image, double_sums=True)
        assert np.abs(single["twists"] - double["twists"]).max() <= 1e-4
        assert np.allclose(double["twist_matrix"], twist_matrix, atol=parameters["tolerance"])
        assert np.abs(double["twists"][-1] - twist).max() <= np.abs(single["twists"][-1] - twist).max()
        assert np.allclose(single["energies"], double["energies"], rtol=2e-3)  # the twists differ by ~4e-5 from iteration 1 on
