#!/usr/bin/env python
"""Golden fixture of the rigid SDF-2-SDF tracker -> tests/golden/reference_rigid.npz.

Run in the build container only (needs /root/reference and cv2 with OpenEXR; the GPU box never runs this):

    OPENCV_IO_ENABLE_OPENEXR=1 python tests/golden/make_rigid_golden.py

The reference holds no golden for its C++ tracker: tests/test_sdf_2_sdf_optimizer.py:81-166 asserts that the C++ twist
matrix equals the Python tracker's within 1e-4. This script RUNS the reference's Python tracker
(rigid_opt/sdf_2_sdf_optimizer2d.py:62-137, with the C++ extension import, matplotlib and the visualiser stubbed, and
np.int and numpy 1's reading of one-element arrays as scalars restored for numpy 2) on the two depth frames of that test with its parameters (and a second parameter set) and
stores
  image/canonical, image/live    the frames as the test prepares them (uint16 millimetres, BGR->gray, 0 -> 65535)
  run/<k>/parameters (json), run/<k>/twist (3), run/<k>/twist_matrix (3 x 3), run/<k>/canonical_field
Only numbers are stored; no reference source is copied.
"""
import json
import os
import sys
import types

import numpy as np

os.environ.setdefault("OPENCV_IO_ENABLE_OPENEXR", "1")
REF = os.environ.get("LSF_REFERENCE", "/root/reference")
OUT = os.path.dirname(os.path.abspath(__file__))
INTRINSICS = [[570.3999633789062, 0, 320], [0, 570.3999633789062, 240], [0, 0, 1]]


def stub_reference_imports():
    module = types.ModuleType("level_set_fusion_optimization")

    class _Enum:
        NONE, BILINEAR_IMAGE_SPACE, BILINEAR_VOXEL_SPACE, EWA_IMAGE_SPACE, EWA_VOXEL_SPACE, EWA_VOXEL_SPACE_INCLUSIVE = range(6)

    module.tsdf = types.SimpleNamespace(FilteringMethod=_Enum)
    sys.modules["level_set_fusion_optimization"] = module
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.patches"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    visualizer = types.ModuleType("rigid_opt.sdf_2_sdf_visualizer")

    class Sdf2SdfVisualizer:
        class Parameters:
            def __init__(self, **kwargs):
                pass

        def __init__(self, *args, **kwargs):
            pass

        def __getattr__(self, name):
            return lambda *args, **kwargs: None

    visualizer.Sdf2SdfVisualizer = Sdf2SdfVisualizer
    sys.modules["rigid_opt.sdf_2_sdf_visualizer"] = visualizer
    if not hasattr(np, "int"):
        np.int = int  # removed in numpy 1.24; the reference writes .astype(np.int)
    # the reference's Python builds arrays from lists that mix scalars with one-element arrays (its offsets and twists are
    # 3 x 1 columns): numpy 1.x read those as scalars, numpy 2 refuses -- restore that reading for this process
    real_array = np.array

    def scalars(value):
        if isinstance(value, np.ndarray) and value.ndim > 0 and value.size == 1:
            return value.reshape(()).item()
        if isinstance(value, (list, tuple)):
            return [scalars(v) for v in value]
        return value

    def compatible_array(value, *args, **kwargs):
        try:
            return real_array(value, *args, **kwargs)
        except ValueError:
            return real_array(scalars(value), *args, **kwargs)

    np.array = compatible_array
    sys.path.insert(0, REF)


def main():
    import cv2
    stub_reference_imports()
    import rigid_opt.sdf_2_sdf_optimizer2d as tracker
    from rigid_opt.sdf_generation import ArrayBasedSingleFrameDataset
    from math_utils import transformation

    def frame(name):  # tests/test_sdf_2_sdf_optimizer.py:121-129
        image = cv2.imread(os.path.join(REF, "tests", "test_data", name), cv2.IMREAD_UNCHANGED)
        image = image.astype(np.uint16)
        image = cv2.cvtColor(image, cv2.COLOR_BGR2GRAY)
        image[image == 0] = np.iinfo(np.uint16).max
        return image

    canonical_image, live_image = frame("depth_000000.exr"), frame("depth_000003.exr")
    camera = types.SimpleNamespace(intrinsics=types.SimpleNamespace(intrinsic_matrix=np.array(INTRINSICS, dtype=np.float32)),
                                   depth_unit_ratio=0.001)
    out = {"image/canonical": canonical_image, "image/live": live_image}
    runs = [dict(source="test_sdf_2_sdf_optimizer.py:81-166 test_operation_same_cpp_to_py", rate=0.5, iterations=8, eta=0.01,
                 narrow_band_width_voxels=2, field_size=32, offset=[-16, -16, 93], image_y_coordinate=240, tolerance=1e-4),
            dict(source="same frames, 20-voxel band, 12 iterations at rate 0.3", rate=0.3, iterations=12, eta=0.01,
                 narrow_band_width_voxels=20, field_size=48, offset=[-24, -24, 85], image_y_coordinate=240, tolerance=1e-4)]
    for k, run in enumerate(runs):
        data = ArrayBasedSingleFrameDataset(canonical_image, live_image, run["image_y_coordinate"], run["field_size"],
                                            np.array(run["offset"], dtype=np.int32).reshape(3, 1), camera)
        optimizer = tracker.Sdf2SdfOptimizer2d(rate=run["rate"])
        twist = optimizer.optimize(data, voxel_size=0.004, narrow_band_width_voxels=run["narrow_band_width_voxels"],
                                   iteration=run["iterations"], eta=run["eta"])
        run["projection_matrix"] = INTRINSICS
        out["run/%02d/parameters" % k] = np.array(json.dumps(run))
        out["run/%02d/twist" % k] = np.asarray(twist, dtype=np.float64).reshape(3)
        out["run/%02d/twist_matrix" % k] = np.asarray(transformation.twist_vector_to_matrix2d(twist), dtype=np.float64)
        out["run/%02d/canonical_field" % k] = np.asarray(
            data.generate_2d_canonical_field(narrow_band_width_voxels=run["narrow_band_width_voxels"]), dtype=np.float32)
        print(run["source"], "twist", np.asarray(twist).ravel())
    np.savez_compressed(os.path.join(OUT, "reference_rigid.npz"), **out)
    print("wrote reference_rigid.npz", os.path.getsize(os.path.join(OUT, "reference_rigid.npz")), "bytes")


if __name__ == "__main__":
    main()
