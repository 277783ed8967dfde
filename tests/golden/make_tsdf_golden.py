#!/usr/bin/env python
"""Golden fixtures of the TSDF generation from depth images -> tests/golden/reference_tsdf.npz.

Run in the build container only (needs /root/reference and cv2; the GPU box never runs this):

    python tests/golden/make_tsdf_golden.py

Contents (only numbers; no reference source is copied)
  image/<name>                 the two depth images of the reference's tests, decoded (uint16 [480][640]):
                               tests/test_data/zigzag2_depth_00108.png, tests/test_data/zigzag1_depth_00064.png
                               (= cpp/tests/data/zigzag2_depth_00108.png, zigzag_depth_00064.png)
  case/<k>/expected, case/<k>/parameters (json)
                               one entry per test case the reference holds for the C++ generators
                               (cpp/tests/test_tsdf.cpp:47-333, tests/test_tsdf_ewa.py:40-235): the golden field with the
                               parameters of the call and the tolerance the reference's own test states.
  python/<k>/expected, python/<k>/parameters (json)
                               outputs of RUNNING the reference's Python generators here (tsdf/generation.py:130-217,356-437,
                               tsdf/ewa.py:59-185,230-600; the C++ extension import is stubbed) on sub-volumes of the
                               real depth images, incl. a rotated + translated camera.
The expected fields of the C++ cases are the literals of cpp/tests/data/test_data_tsdf.hpp and
tests/test_data/ewa_test_data.py (already in reference_literals.npz for the .hpp; the .py ones are imported here).
"""
import json
import os
import re
import sys
import types

import numpy as np

REF = os.environ.get("LSF_REFERENCE", "/root/reference")
OUT = os.path.dirname(os.path.abspath(__file__))

INTRINSICS = [[700.0, 0.0, 320.0], [0.0, 700.0, 240.0], [0.0, 0.0, 1.0]]
NONE, EWA_IMAGE, EWA_VOXEL, EWA_VOXEL_INCLUSIVE = 0, 3, 4, 5


def stub_reference_imports():
    """tsdf/generation.py and tsdf/ewa.py import the C++ extension at module load (for the FilteringMethod enum keys of
    two dispatch tables) -- give them an empty stand-in."""
    module = types.ModuleType("level_set_fusion_optimization")

    class _Enum:
        NONE, BILINEAR_IMAGE_SPACE, BILINEAR_VOXEL_SPACE, EWA_IMAGE_SPACE, EWA_VOXEL_SPACE, EWA_VOXEL_SPACE_INCLUSIVE = range(6)
        EWA_IMAGE, EWA_TSDF, EWA_TSDF_INCLUSIVE = 3, 4, 5

    tsdf_scope = types.SimpleNamespace(FilteringMethod=_Enum, InterpolationMethod=_Enum)
    module.tsdf = tsdf_scope
    sys.modules["level_set_fusion_optimization"] = module
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.patches"):  # imported for plotting helpers only
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.path.insert(0, REF)


class Camera:
    """duck-typed calib.camera.DepthCamera: .intrinsics.intrinsic_matrix, .depth_unit_ratio"""

    def __init__(self, matrix, depth_unit_ratio=0.001):
        self.intrinsics = types.SimpleNamespace(intrinsic_matrix=np.array(matrix, dtype=np.float32))
        self.depth_unit_ratio = depth_unit_ratio


def parse_hpp_tensor(text, name):
    """float data[] = {...} of `static eig::Tensor<float, 3> name`: a row-major 16 x 1 x 16 block whose layout is swapped
    into the column-major tensor (TensorLayoutSwapOp reverses the index order): field(x, y, z) = data[z][y][x]"""
    start = text.index("static eig::Tensor<float, 3> " + name)
    block = text[text.index("{", text.index("float data[]", start)) + 1:]
    block = block[:block.index("};")]
    return np.array([float(t.rstrip("f")) for t in re.findall(r"[-+]?\d+\.?\d*(?:[eE][-+]?\d+)?f?", block)],
                    dtype=np.float32).reshape(16, 1, 16).transpose(2, 1, 0).copy()


def main():
    import cv2
    stub_reference_imports()
    out = {}
    images = {}
    for key, name in (("zigzag2_108", "zigzag2_depth_00108.png"), ("zigzag1_064", "zigzag1_depth_00064.png")):
        image = cv2.imread(os.path.join(REF, "tests", "test_data", name), cv2.IMREAD_UNCHANGED)
        assert image.dtype == np.uint16 and image.shape == (480, 640)
        other = cv2.imread(os.path.join(REF, "cpp", "tests", "data",
                                        name.replace("zigzag1_depth_00064", "zigzag_depth_00064")), cv2.IMREAD_UNCHANGED)
        assert np.array_equal(image, other)
        images[key] = image
        out["image/" + key] = image
    literals = np.load(os.path.join(OUT, "reference_literals.npz"))
    hpp = open(os.path.join(REF, "cpp", "tests", "data", "test_data_tsdf.hpp")).read()
    import tests.test_data.ewa_test_data as ewa_data

    def literal(name):
        keys = [k for k in literals.files if k.startswith("test_data_tsdf/%s/" % name)]
        assert len(keys) == 1, (name, keys)
        return literals[keys[0]].astype(np.float32)

    region = literal("depth_image_region").astype(np.uint16)
    synthetic = np.full((3, 640), 65535, dtype=np.uint16)
    synthetic[:, 399:417] = region
    out["image/synthetic_region"] = synthetic
    pose_shifted = np.eye(4, dtype=np.float32)
    pose_shifted[2, 3] = 0.004
    identity = np.eye(4, dtype=np.float32)
    cases = []

    def case(source, tolerance, expected, image, nd, method, offset, shape, y=0, pose=identity, smoothing=1.0,
             zero_to_max=False):
        cases.append((dict(source=source, tolerance=tolerance, image=image, nd=nd, filtering_method=method,
                           array_offset=list(offset), field_shape=list(shape), image_y_coordinate=y,
                           camera_pose=np.asarray(pose, dtype=np.float32).tolist(), smoothing_factor=smoothing,
                           projection_matrix=INTRINSICS, depth_unit_ratio=0.001, near_clipping_distance=0.05,
                           voxel_size=0.004, narrow_band_width_voxels=20, zero_depth_to_maximum=zero_to_max),
                      np.asarray(expected, dtype=np.float32)))

    # cpp/tests/test_tsdf.cpp (raw images as read by read_image_helper)
    case("test_tsdf.cpp:47-70 no_interpolation_1", 1e-6, literal("expected_tsdf_field01"), "zigzag2_108", 2, NONE,
         (-8, 144), (16, 16), y=200)
    case("test_tsdf.cpp:72-97 no_interpolation_2", 1e-6, literal("expected_tsdf_field02"), "zigzag2_108", 2, NONE,
         (-8, 144), (16, 16), y=200, pose=pose_shifted)
    case("test_tsdf.cpp:99-125 EWA_2D_image_space_1", 1e-6, literal("out_sdf_field"), "synthetic_region", 2, EWA_IMAGE,
         (94, 804), (16, 16), y=1)
    case("test_tsdf.cpp:127-177 EWA_2D_image_space_2", 1e-6, literal("out_sdf_chunk"), "zigzag2_108", 2, EWA_IMAGE,
         (-256 + 210, 0 + 103), (16, 16), y=200)
    case("test_tsdf.cpp:179-227 EWA_2D_voxel_space_inclusive_1", 1e-6, literal("expected_tsdf_EWA_voxel_space_inclusive_1"),
         "zigzag2_108", 2, EWA_VOXEL_INCLUSIVE, (-256 + 210, 0 + 103), (16, 16), y=200)
    case("test_tsdf.cpp:229-277 EWA_2D_voxel_space_inclusive_2", 1e-6, literal("expected_tsdf_EWA_voxel_space_inclusive_2"),
         "zigzag1_064", 2, EWA_VOXEL_INCLUSIVE, (-256 + 24, 480 + 10), (16, 16), y=200)
    case("test_tsdf.cpp:279-305 EWA_3D_image_space_1", 1e-6, parse_hpp_tensor(hpp, "TSDF_slice01"), "zigzag2_108", 3,
         EWA_IMAGE, (-46, -8, 105), (16, 1, 16))
    case("test_tsdf.cpp:307-333 EWA_3D_image_space_2", 1e-5, parse_hpp_tensor(hpp, "TSDF_slice02"), "zigzag2_108", 3,
         EWA_IMAGE, (-46, -8, 105), (16, 1, 16), smoothing=0.5)
    # tests/test_tsdf_ewa.py, C++ halves (image_load_helper replaces depth 0 by 65535)
    case("test_tsdf_ewa.py:40-70 2D_ewa_tsdf_generation1", 1e-6, ewa_data.out_sdf_field01, "synthetic_region", 2, EWA_IMAGE,
         (94, 804), (16, 16), y=1)
    case("test_tsdf_ewa.py:117-144 2D_ewa_tsdf_generation3", 1e-5, ewa_data.out_sdf_field03, "zigzag1_064", 2, EWA_VOXEL,
         (-232, 490), (16, 16), y=1, smoothing=0.5, zero_to_max=True)
    case("test_tsdf_ewa.py:146-173 2D_ewa_tsdf_generation4", 1e-5, ewa_data.out_sdf_field04, "zigzag1_064", 2,
         EWA_VOXEL_INCLUSIVE, (-232, 490), (16, 16), y=1, smoothing=0.5, zero_to_max=True)
    case("test_tsdf_ewa.py:175-203 3d_ewa_tsdf_generation1", 1e-6, ewa_data.sdf_3d_slice01, "zigzag2_108", 3, EWA_IMAGE,
         (-46, -8, 105), (16, 1, 16), zero_to_max=True)
    case("test_tsdf_ewa.py:205-235 3d_ewa_tsdf_generation2", 1e-5, ewa_data.sdf_3d_slice02, "zigzag2_108", 3, EWA_IMAGE,
         (-46, -8, 105), (16, 1, 16), smoothing=0.5, zero_to_max=True)
    for k, (parameters, expected) in enumerate(cases):
        out["case/%02d/expected" % k] = expected
        out["case/%02d/parameters" % k] = np.array(json.dumps(parameters))

    # ---- runs of the reference's Python generators
    import tsdf.generation as generation
    import tsdf.ewa as ewa
    camera = Camera(INTRINSICS)
    angle = 0.05
    pose_rotated = np.array([[np.cos(angle), 0, np.sin(angle), 0.01], [0, 1, 0, -0.02], [-np.sin(angle), 0, np.cos(angle), 0.03],
                             [0, 0, 0, 1]], dtype=np.float32)
    runs = []

    def run(source, tolerance, field, image, nd, method, offset, shape, y=0, pose=identity, smoothing=1.0, near=0.05):
        runs.append((dict(source=source, tolerance=tolerance, image=image, nd=nd, filtering_method=method,
                          array_offset=[int(v) for v in offset], field_shape=[int(v) for v in shape], image_y_coordinate=y,
                          camera_pose=np.asarray(pose, dtype=np.float32).tolist(), smoothing_factor=smoothing,
                          projection_matrix=INTRINSICS, depth_unit_ratio=0.001, near_clipping_distance=near,
                          voxel_size=0.004, narrow_band_width_voxels=20, zero_depth_to_maximum=True),
                     np.asarray(field, dtype=np.float32)))

    def prepared(key):
        image = images[key].copy()
        image[image == 0] = 65535
        return image

    for key, offset3 in (("zigzag2_108", (-46, -8, 105)), ("zigzag1_064", (-10, -12, 700))):
        image = prepared(key)
        for pose in (identity, pose_rotated):
            tag = "identity" if pose is identity else "rotated"
            # the Python 3D generator indexes field[z][y][x] (generation.py:432); the C++ one [x][y][z]: transpose.
            # It clips at depth <= 0 (:405): near_clipping_distance 0 in the restatement.
            field = generation.generate_3d_tsdf_field_from_depth_image(image, camera, pose, field_size=20,
                                                                      array_offset=np.array(offset3))
            run("generation.py:356-437 %s %s" % (key, tag), 1e-5, field.transpose(2, 1, 0), key, 3, NONE, offset3,
                (20, 20, 20), pose=pose, near=0.0)
            field = generation.generate_2d_tsdf_field_from_depth_image_no_interpolation(
                image, camera, 200, pose, field_size=24, array_offset=np.array(offset3))
            run("generation.py:130-217 %s %s" % (key, tag), 1e-5, field, key, 2, NONE, (offset3[0], offset3[2]), (24, 24),
                y=200, pose=pose, near=0.0)
        field = ewa.generate_tsdf_3d_ewa_image(image, camera, identity, field_shape=np.array([10, 3, 10]),
                                               array_offset=np.array(offset3), gaussian_covariance_scale=0.5)
        run("ewa.py:59-185 %s" % key, 2e-5, field, key, 3, EWA_IMAGE, offset3, (10, 3, 10), smoothing=0.5)
    for k, (parameters, expected) in enumerate(runs):
        out["python/%02d/expected" % k] = expected
        out["python/%02d/parameters" % k] = np.array(json.dumps(parameters))
    np.savez_compressed(os.path.join(OUT, "reference_tsdf.npz"), **out)
    print("wrote reference_tsdf.npz:", len(cases), "reference test cases,", len(runs), "runs of the reference's Python;",
          os.path.getsize(os.path.join(OUT, "reference_tsdf.npz")), "bytes")


if __name__ == "__main__":
    main()
