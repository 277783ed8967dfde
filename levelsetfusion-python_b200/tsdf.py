"""TSDF generation from depth images -- host-side mirror of the reference's `level_set_fusion_optimization.tsdf` scope
(reference cpp/src/python_export/tsdf.cpp:41-121): FilteringMethod, Parameters2d / Parameters3d, Generator2d / Generator3d
with the same constructor arguments, attribute names and defaults (reference cpp/src/tsdf/parameters.hpp:31-56), backed by
lsf_tsdf_generate (csrc/tsdf.cu). This is the step right before the optimisation path: the
reference's experiment code builds its canonical / live fields with exactly these calls
(experiment/dataset.py:103-170).

Besides numpy depth images (uint16, as cv2.imread(..., IMREAD_UNCHANGED) returns them) the generators accept CUDA torch
tensors (int16 / uint16 storage) and then return a CUDA tensor, so that a generated pair can be handed to an optimizer
without leaving the device.
"""
import ctypes
import enum

import numpy as np

from . import _lib
from .telemetry import Vector2i, Vector3i


class FilteringMethod(enum.IntEnum):
    """reference tsdf::FilteringMethod, cpp/src/tsdf/interpolation_method.hpp:39-46"""
    NONE = 0
    BILINEAR_IMAGE_SPACE = 1
    BILINEAR_VOXEL_SPACE = 2
    EWA_IMAGE_SPACE = 3
    EWA_VOXEL_SPACE = 4
    EWA_VOXEL_SPACE_INCLUSIVE = 5


class _Parameters:
    _nd = 0

    def __init__(self, depth_unit_ratio=0.001, projection_matrix=None, near_clipping_distance=0.05, array_offset=None,
                 field_shape=None, voxel_size=0.004, narrow_band_width_voxels=20,
                 interpolation_method=FilteringMethod.NONE, smoothing_factor=1.0):
        vector = Vector2i if self._nd == 2 else Vector3i
        self.depth_unit_ratio = depth_unit_ratio
        self.projection_matrix = np.identity(3, dtype=np.float32) if projection_matrix is None else projection_matrix
        self.near_clipping_distance = near_clipping_distance
        self.array_offset = vector(-64) if array_offset is None else array_offset
        self.field_shape = vector(128) if field_shape is None else field_shape
        self.voxel_size = voxel_size
        self.narrow_band_width_voxels = narrow_band_width_voxels
        self.interpolation_method = interpolation_method
        self.smoothing_factor = smoothing_factor

    def _raw(self):
        raw = _lib.TsdfParams()
        raw.depth_unit_ratio = float(self.depth_unit_ratio)
        matrix = np.asarray(self.projection_matrix)
        if matrix.shape != (3, 3):
            raise ValueError("projection_matrix must be 3 x 3, got shape %s" % (matrix.shape,))
        raw.projection_matrix = (ctypes.c_float * 9)(*[float(v) for v in matrix.reshape(9)])
        raw.near_clipping_distance = float(self.near_clipping_distance)
        offset = [int(v) for v in self.array_offset]
        shape = [int(v) for v in self.field_shape]
        if len(offset) != self._nd or len(shape) != self._nd:
            raise ValueError("array_offset and field_shape need %d coordinates" % self._nd)
        raw.array_offset = (ctypes.c_int * 3)(*(offset + [0] * (3 - self._nd)))
        raw.field_shape = (ctypes.c_int * 3)(*(shape + [1] * (3 - self._nd)))
        raw.voxel_size = float(self.voxel_size)
        raw.narrow_band_width_voxels = int(self.narrow_band_width_voxels)
        raw.filtering_method = int(self.interpolation_method)
        raw.smoothing_factor = float(self.smoothing_factor)
        return raw, shape


class Parameters2d(_Parameters):
    """reference tsdf.Parameters2d (python_export/tsdf.cpp:52-66): array_offset / field_shape are Vector2i(x, y) with
    x along the image's x direction and y along the depth direction"""
    _nd = 2


class Parameters3d(_Parameters):
    """reference tsdf.Parameters3d (python_export/tsdf.cpp:68-82)"""
    _nd = 3


class _Generator:
    _nd = 0

    def __init__(self, parameters):
        if getattr(parameters, "_nd", None) != self._nd:
            raise TypeError("expected tsdf.Parameters%dd" % self._nd)
        # the reference's generator copies the parameters at construction (generator_crtp.tpp:34-36)
        vector = Vector2i if self._nd == 2 else Vector3i
        self.parameters = type(parameters)(parameters.depth_unit_ratio, np.array(parameters.projection_matrix, dtype=np.float32),
                                           parameters.near_clipping_distance, vector(*[int(v) for v in parameters.array_offset]),
                                           vector(*[int(v) for v in parameters.field_shape]), parameters.voxel_size,
                                           parameters.narrow_band_width_voxels, parameters.interpolation_method,
                                           parameters.smoothing_factor)

    def generate(self, depth_image, camera_pose=None, image_y_coordinate=0):
        """reference Generator{2d,3d}::generate(depth_image, camera_pose, image_y_coordinate), generator_crtp.tpp:40-71.
        depth_image: uint16 [rows][cols]; camera_pose: 4 x 4 float32 (identity when omitted); image_y_coordinate: the
        image row a 2D field is generated from. Returns float32 [shape.y][shape.x] (2D) or [shape.x][shape.y][shape.z]."""
        raw, shape = self.parameters._raw()
        pose = np.identity(4, dtype=np.float32) if camera_pose is None else \
            np.ascontiguousarray(np.asarray(camera_pose), dtype=np.float32)
        if pose.shape != (4, 4):
            raise ValueError("camera_pose must be 4 x 4, got shape %s" % (pose.shape,))
        out_shape = (shape[1], shape[0]) if self._nd == 2 else tuple(shape)
        if _lib.is_torch_cuda(depth_image):
            import torch
            _lib.check_device(depth_image)
            if depth_image.dtype not in (torch.int16, torch.uint16) or depth_image.dim() != 2:
                raise ValueError("depth_image must be a 2D 16-bit integer tensor")
            image = depth_image.contiguous()
            field = torch.empty(out_shape, dtype=torch.float32, device=image.device)
            depth_pointer = ctypes.cast(image.data_ptr(), ctypes.POINTER(ctypes.c_ushort))
            field_pointer = ctypes.cast(field.data_ptr(), _lib.c_float_p)
            kind, stream = _lib.LSF_DEVICE, _lib.current_stream_handle()
            rows, cols = int(image.shape[0]), int(image.shape[1])
        else:
            image = np.asarray(depth_image)
            if image.dtype != np.uint16 or image.ndim != 2:
                # reference converter: only unsigned-short matrices are convertible (eigen_numpy_matrix.cpp:79-103)
                raise ValueError("depth_image must be a 2D uint16 array, got %s with %d dimensions" % (image.dtype, image.ndim))
            image = np.ascontiguousarray(image)
            field = _lib.result_array(out_shape)
            depth_pointer = image.ctypes.data_as(ctypes.POINTER(ctypes.c_ushort))
            field_pointer = _lib.fptr(field)
            kind, stream = _lib.LSF_HOST, _lib.host_stream_handle()
            rows, cols = image.shape
        _lib.check(_lib.load().lsf_tsdf_generate(ctypes.byref(raw), depth_pointer, int(rows), int(cols), _lib.fptr(pose),
                                                 int(image_y_coordinate), self._nd, field_pointer, kind, stream))
        return field


class Generator2d(_Generator):
    """reference tsdf.Generator2d (python_export/tsdf.cpp:84-89)"""
    _nd = 2


class Generator3d(_Generator):
    """reference tsdf.Generator3d (python_export/tsdf.cpp:91-96)"""
    _nd = 3
