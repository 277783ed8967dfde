"""Rigid SDF-2-SDF tracker in 2D -- host-side mirror of the reference's `Sdf2SdfOptimizer2d`
(reference cpp/src/python_export/sdf_2_sdf_optimizer.cpp:22-58, cpp/src/rigid_optimization/sdf_2_sdf_optimizer2d.hpp:22-53),
backed by lsf_sdf2sdf_optimize_2d (csrc/rigid.cu). SURVEY.md 8f, row f4: the step before the non-rigid alignment."""
import ctypes

import numpy as np

from . import _lib
from . import tsdf as _tsdf


class Sdf2SdfOptimizer2d:
    """reference Sdf2SdfOptimizer2d(rate=0.5, maximum_iteration_count=60, tsdf_generation_parameters=tsdf.Parameters2d(),
    verbosity_parameters=VerbosityParameters())"""

    class VerbosityParameters:
        """reference Sdf2SdfOptimizer2d::VerbosityParameters (sdf_2_sdf_optimizer2d.cpp:43-54): read-only flags"""

        def __init__(self, print_iteration_max_warp_update=False, print_iteration_energy=False):
            self._print_iteration_max_warp_update = bool(print_iteration_max_warp_update)
            self._print_iteration_energy = bool(print_iteration_energy)

        print_iteration_max_warp_update = property(lambda self: self._print_iteration_max_warp_update)
        print_iteration_energy = property(lambda self: self._print_iteration_energy)
        print_per_iteration_info = property(
            lambda self: self._print_iteration_max_warp_update or self._print_iteration_energy)

    def __init__(self, rate=0.5, maximum_iteration_count=60, tsdf_generation_parameters=None, verbosity_parameters=None):
        self.rate = float(rate)
        self.maximum_iteration_count = int(maximum_iteration_count)
        parameters = _tsdf.Parameters2d() if tsdf_generation_parameters is None else tsdf_generation_parameters
        # the reference builds its generator (a copy of the parameters) in the constructor
        self._tsdf_generator = _tsdf.Generator2d(parameters)
        self.verbosity_parameters = verbosity_parameters or Sdf2SdfOptimizer2d.VerbosityParameters()
        self._twists = self._optimal_twists = self._energies = None

    def optimize(self, image_y_coordinate, canonical_field, live_depth_image, eta=0.01, initial_camera_pose=None):
        """Find the twist that maps the live depth frame's TSDF onto canonical_field; returns the 3 x 3 twist matrix
        (float32), reference sdf_2_sdf_optimizer2d.cpp:63-124. canonical_field: float32 [shape.y][shape.x];
        live_depth_image: uint16 [rows][cols]; both numpy arrays, or both CUDA torch tensors (16-bit integer image)."""
        raw, shape = self._tsdf_generator.parameters._raw()
        expected_shape = (shape[1], shape[0])
        iterations = self.maximum_iteration_count
        matrix = np.zeros((3, 3), dtype=np.float32)
        twists = np.zeros((max(iterations, 1), 3), dtype=np.float32)
        optimal = np.zeros((max(iterations, 1), 3), dtype=np.float32)
        energies = np.zeros(max(iterations, 1), dtype=np.float32)
        pose = None if initial_camera_pose is None else np.ascontiguousarray(initial_camera_pose, dtype=np.float32)
        if _lib.is_torch_cuda(canonical_field) or _lib.is_torch_cuda(live_depth_image):
            import torch
            _lib.check_device(canonical_field, live_depth_image)
            if not (_lib.is_torch_cuda(canonical_field) and _lib.is_torch_cuda(live_depth_image)):
                raise ValueError("canonical_field and live_depth_image must both be CUDA tensors or both numpy arrays")
            if live_depth_image.dtype not in (torch.int16, torch.uint16) or live_depth_image.dim() != 2:
                raise ValueError("live_depth_image must be a 2D 16-bit integer tensor")
            canonical = canonical_field.contiguous().float()
            image = live_depth_image.contiguous()
            canonical_pointer = ctypes.cast(canonical.data_ptr(), _lib.c_float_p)
            depth_pointer = ctypes.cast(image.data_ptr(), ctypes.POINTER(ctypes.c_ushort))
            kind, stream = _lib.LSF_DEVICE, _lib.current_stream_handle()
        else:
            canonical = _lib.as_f32(canonical_field)
            image = np.asarray(live_depth_image)
            if image.dtype != np.uint16 or image.ndim != 2:
                raise ValueError("live_depth_image must be a 2D uint16 array, got %s with %d dimensions"
                                 % (image.dtype, image.ndim))
            image = np.ascontiguousarray(image)
            canonical_pointer = _lib.fptr(canonical)
            depth_pointer = image.ctypes.data_as(ctypes.POINTER(ctypes.c_ushort))
            kind, stream = _lib.LSF_HOST, _lib.host_stream_handle()
        if tuple(canonical.shape) != expected_shape:
            raise ValueError("canonical_field has shape %s, the TSDF parameters say %s" % (tuple(canonical.shape), expected_shape))
        _lib.check(_lib.load().lsf_sdf2sdf_optimize_2d(
            ctypes.byref(raw), ctypes.c_float(self.rate), iterations, int(image_y_coordinate), canonical_pointer, depth_pointer,
            int(image.shape[0]), int(image.shape[1]), ctypes.c_float(eta), None if pose is None else _lib.fptr(pose),
            _lib.fptr(matrix), _lib.fptr(twists), _lib.fptr(optimal), _lib.fptr(energies), kind, stream))
        self._twists, self._optimal_twists, self._energies = twists[:iterations], optimal[:iterations], energies[:iterations]
        verbosity = self.verbosity_parameters
        if verbosity.print_per_iteration_info:  # reference sdf_2_sdf_optimizer2d.cpp:107-117
            for k in range(iterations):
                print("[ITERATION %d COMPLETED]" % k)
                if verbosity.print_iteration_max_warp_update:
                    print(" [optimize twist:%s]" % " ".join("%g" % v for v in optimal[k]))
                    print(" [twist:%s]" % " ".join("%g" % v for v in twists[k]))
                if verbosity.print_iteration_energy:
                    print(" [energy: %g]\n\n" % energies[k])
        return matrix

    # extensions (not in the reference): the per-iteration numbers of the last optimize() call
    def get_per_iteration_twists(self):
        return self._twists

    def get_per_iteration_energies(self):
        return self._energies
