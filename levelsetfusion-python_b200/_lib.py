"""ctypes binding of liblsf_b200.so (the C-ABI declared in include/lsf_b200.h).

There is NO CPU fallback: if the library is missing or no CUDA device is usable, calls raise.
"""
import ctypes
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liblsf_b200.so")

LSF_HOST = 0
LSF_DEVICE = 1
LSF_MAX_LEVELS = 16
LSF_SEMANTICS_CPP, LSF_SEMANTICS_PY_DIRECT, LSF_SEMANTICS_PY_VECTORIZED = 0, 1, 2

c_float_p = ctypes.POINTER(ctypes.c_float)
c_int_p = ctypes.POINTER(ctypes.c_int)


class HierParams(ctypes.Structure):
    """lsf_hier_params"""
    _fields_ = [
        ("tikhonov_term_enabled", ctypes.c_int),
        ("gradient_kernel_enabled", ctypes.c_int),
        ("maximum_chunk_size", ctypes.c_int),
        ("rate", ctypes.c_float),
        ("maximum_iteration_count", ctypes.c_int),
        ("maximum_warp_update_threshold", ctypes.c_float),
        ("data_term_amplifier", ctypes.c_float),
        ("tikhonov_strength", ctypes.c_float),
        ("kernel", c_float_p),
        ("kernel_size", ctypes.c_int),
        ("resampling_strategy", ctypes.c_int),
    ]


class LevelReport(ctypes.Structure):
    """lsf_level_report"""
    _fields_ = [
        ("iteration_count", ctypes.c_int),
        ("iteration_limit_reached", ctypes.c_int),
        ("max_update_length", ctypes.c_float),
        ("dims", ctypes.c_int * 3),
        ("warp_ratio_above_min_threshold", ctypes.c_float),
        ("warp_length_min", ctypes.c_float),
        ("warp_length_max", ctypes.c_float),
        ("warp_length_mean", ctypes.c_float),
        ("warp_length_std", ctypes.c_float),
        ("warp_longest_location", ctypes.c_int * 3),
        ("warp_is_largest_below_min_threshold", ctypes.c_int),
        ("warp_is_largest_above_max_threshold", ctypes.c_int),
        ("diff_min", ctypes.c_float),
        ("diff_max", ctypes.c_float),
        ("diff_mean", ctypes.c_float),
        ("diff_std", ctypes.c_float),
        ("diff_biggest_location", ctypes.c_int * 3),
    ]


class IterationCapture(ctypes.Structure):
    """lsf_iteration_capture"""
    _fields_ = [
        ("level", ctypes.c_int),
        ("max_iterations", ctypes.c_int),
        ("buffer", c_float_p),
        ("count", ctypes.c_int),
    ]


class IterationRecord(ctypes.Structure):
    """lsf_iteration_record (include/lsf_b200.h)"""
    _fields_ = [
        ("level", ctypes.c_int),
        ("iteration", ctypes.c_int),
        ("dims", ctypes.c_int * 3),
        ("max_update_length", ctypes.c_float),
        ("mean_tsdf_difference", ctypes.c_float),
        ("std_tsdf_difference", ctypes.c_float),
        ("normalized_data_energy", ctypes.c_float),
        ("normalized_tikhonov_energy", ctypes.c_float),
        ("live_field", c_float_p),
        ("warp_field", c_float_p),
        ("data_term_gradient", c_float_p),
        ("tikhonov_term_gradient", c_float_p),
    ]


ITERATION_CALLBACK = ctypes.CFUNCTYPE(None, ctypes.c_void_p, ctypes.POINTER(IterationRecord))


class IterationSink(ctypes.Structure):
    """lsf_iteration_sink (include/lsf_b200.h)"""
    _fields_ = [
        ("callback", ITERATION_CALLBACK),
        ("user", ctypes.c_void_p),
        ("want_fields", ctypes.c_int),
        ("want_statistics", ctypes.c_int),
    ]


class SlabLevel(ctypes.Structure):
    """lsf_slab_level"""
    _fields_ = [
        ("planes", ctypes.c_int), ("Y", ctypes.c_int), ("Z", ctypes.c_int),
        ("own_begin", ctypes.c_int), ("own_end", ctypes.c_int),
        ("x_origin", ctypes.c_int), ("X_global", ctypes.c_int),
        ("pack", ctypes.c_void_p), ("pack_planes", ctypes.c_int), ("pack_origin", ctypes.c_int),
        ("pack_interior_low", ctypes.c_int), ("pack_interior_high", ctypes.c_int),
        ("canonical", ctypes.c_void_p), ("warp", ctypes.c_void_p), ("g_post", ctypes.c_void_p),
        ("g_pre", ctypes.c_void_p), ("max_sq_bits", ctypes.c_void_p), ("violation", ctypes.c_void_p),
    ]


SLAB_MAX_PEERS = 8          # LSF_SLAB_MAX_PEERS
SLAB_MAILBOX_BYTES = 512    # LSF_SLAB_MAILBOX_BYTES
PEER_HANDLE_BYTES = 64      # LSF_PEER_HANDLE_BYTES


class SlabPeers(ctypes.Structure):
    """lsf_slab_peers"""
    _fields_ = [
        ("rank", ctypes.c_int), ("world_size", ctypes.c_int),
        ("base", ctypes.c_void_p * SLAB_MAX_PEERS),
        ("mailbox_offset", ctypes.c_size_t),
    ]


class SlabLink(ctypes.Structure):
    """lsf_slab_link"""
    _fields_ = [
        ("pre_offset", ctypes.c_size_t), ("post_offset", ctypes.c_size_t),
        ("low_planes", ctypes.c_int), ("low_own_end", ctypes.c_int),
        ("high_planes", ctypes.c_int), ("high_own_begin", ctypes.c_int),
    ]


class SlavchevaParams(ctypes.Structure):
    """lsf_slavcheva_params"""
    _fields_ = [
        ("semantics", ctypes.c_int),
        ("data_term_method", ctypes.c_int),
        ("smoothing_term_method", ctypes.c_int),
        ("level_set_term_enabled", ctypes.c_int),
        ("sobolev_smoothing_enabled", ctypes.c_int),
        ("gradient_descent_rate", ctypes.c_float),
        ("data_term_weight", ctypes.c_float),
        ("smoothing_term_weight", ctypes.c_float),
        ("isomorphic_enforcement_factor", ctypes.c_float),
        ("level_set_term_weight", ctypes.c_float),
        ("maximum_warp_length_lower_threshold", ctypes.c_float),
        ("maximum_warp_length_upper_threshold", ctypes.c_float),
        ("maximum_iteration_count", ctypes.c_int),
        ("minimum_iteration_count", ctypes.c_int),
        ("sobolev_kernel", c_float_p),
        ("sobolev_kernel_size", ctypes.c_int),
    ]


class WarpDeltaStatisticsRaw(ctypes.Structure):
    """lsf_warp_delta_statistics_t"""
    _fields_ = [
        ("ratio_above_min_threshold", ctypes.c_float),
        ("length_min", ctypes.c_float),
        ("length_max", ctypes.c_float),
        ("length_mean", ctypes.c_float),
        ("length_standard_deviation", ctypes.c_float),
        ("longest_warp_location", ctypes.c_int * 3),
        ("is_largest_below_min_threshold", ctypes.c_int),
        ("is_largest_above_max_threshold", ctypes.c_int),
    ]


class TsdfDifferenceStatisticsRaw(ctypes.Structure):
    """lsf_tsdf_difference_statistics_t"""
    _fields_ = [
        ("difference_min", ctypes.c_float),
        ("difference_max", ctypes.c_float),
        ("difference_mean", ctypes.c_float),
        ("difference_standard_deviation", ctypes.c_float),
        ("biggest_difference_location", ctypes.c_int * 3),
    ]


class SlavchevaReport(ctypes.Structure):
    """lsf_slavcheva_report"""
    _fields_ = [
        ("iteration_count", ctypes.c_int),
        ("iteration_limit_reached", ctypes.c_int),
        ("last_max_warp_length", ctypes.c_float),
        ("has_statistics", ctypes.c_int),
        ("warp_delta_statistics", WarpDeltaStatisticsRaw),
        ("tsdf_difference_statistics", TsdfDifferenceStatisticsRaw),
    ]


class TsdfParams(ctypes.Structure):
    """lsf_tsdf_params"""
    _fields_ = [
        ("depth_unit_ratio", ctypes.c_float),
        ("projection_matrix", ctypes.c_float * 9),
        ("near_clipping_distance", ctypes.c_float),
        ("array_offset", ctypes.c_int * 3),
        ("field_shape", ctypes.c_int * 3),
        ("voxel_size", ctypes.c_float),
        ("narrow_band_width_voxels", ctypes.c_int),
        ("filtering_method", ctypes.c_int),
        ("smoothing_factor", ctypes.c_float),
    ]


# every symbol include/lsf_b200.h declares (tests check that the library exports all of them)
EXPORTED_SYMBOLS = [
    "lsf_last_error", "lsf_version", "lsf_launch_count", "lsf_trim",
    "lsf_hier_optimize_3d", "lsf_hier_optimize_2d", "lsf_hier_optimize_3d_batch", "lsf_hier_iterate_3d",
    "lsf_warp_3d", "lsf_warp_2d", "lsf_gradient_3d", "lsf_gradient_2d", "lsf_laplacian_3d", "lsf_laplacian_2d",
    "lsf_convolve_3d", "lsf_convolve_2d", "lsf_downsample_3d", "lsf_upsample_3d", "lsf_downsample_2d",
    "lsf_upsample_2d", "lsf_max_norm", "lsf_locate_max_norm",
    "lsf_hier_slab_iteration", "lsf_slab_pack_finest", "lsf_slab_restrict", "lsf_slab_prolong_nearest",
    "lsf_debug_last_path", "lsf_hier_optimize_3d_telemetry", "lsf_hier_optimize_2d_telemetry",
    "lsf_slavcheva_optimize", "lsf_slavcheva_optimize_logged", "lsf_warp_advanced", "lsf_warp_delta_statistics", "lsf_tsdf_difference_statistics",
    "lsf_tsdf_generate", "lsf_sdf2sdf_optimize_2d",
    "lsf_peer_alloc", "lsf_peer_open", "lsf_peer_close", "lsf_peer_free", "lsf_slab_exchange", "lsf_slab_exchange_error",
    "lsf_hier_slab_iterations",
]

_lib = None


class LsfError(RuntimeError):
    """Raised for every failing library call (the reference raises RuntimeError through Boost.Python's
    default translator for its AssertionFailureException, error_handling/throw_assert.hpp:67-74)."""


def load():
    """Loads liblsf_b200.so; raises if it has not been built (no fallback path exists)."""
    global _lib
    if _lib is None:
        path = os.environ.get("LSF_B200_LIBRARY", LIB_PATH)  # experiment builds of the same sources (build.py -D... -o ...)
        if not os.path.exists(path):
            raise LsfError("liblsf_b200.so is missing at %s: build it with `python __graft_entry__.py` "
                           "(nvcc, sm_100a). There is no CPU fallback." % path)
        lib = ctypes.CDLL(path)
        lib.lsf_last_error.restype = ctypes.c_char_p
        lib.lsf_launch_count.restype = ctypes.c_longlong
        for name in EXPORTED_SYMBOLS:
            getattr(lib, name)  # AttributeError if the header and the library ever diverge
        _lib = lib
    return _lib


def check(status):
    if status < 0:
        raise LsfError(load().lsf_last_error().decode("utf-8", "replace"))
    return status


# The reference's Boost.Python converters only accept float32, C-contiguous, aligned numpy arrays
# (python_export/eigen_numpy_tensor.cpp:120-156, eigen_numpy_matrix.cpp): any other argument fails overload resolution with
# Boost.Python.ArgumentError, a TypeError. By default this package is lenient and converts; set_strict_inputs(True) (or
# LSF_STRICT_INPUTS=1) reproduces the reference's refusal, e.g. to find silent float64 copies in a pipeline.
_strict_inputs = os.environ.get("LSF_STRICT_INPUTS", "0") == "1"


def set_strict_inputs(enabled):
    """True: field arguments must be float32 C-contiguous aligned numpy arrays, as for the reference's extension module
    (TypeError otherwise); False (default): anything array-like is converted. Returns the previous setting."""
    global _strict_inputs
    previous, _strict_inputs = _strict_inputs, bool(enabled)
    return previous


def as_f32(array, name="array"):
    """float32 C-contiguous view or copy of a field argument (see set_strict_inputs)"""
    if _strict_inputs:
        if not isinstance(array, np.ndarray):
            raise TypeError("%s must be a numpy array, got %s" % (name, type(array).__name__))
        if array.dtype != np.float32 or not array.flags.c_contiguous or not array.flags.aligned:
            raise TypeError("%s must be a float32, C-contiguous, aligned numpy array (got dtype %s, C-contiguous %s): the "
                            "reference's converters accept nothing else (python_export/eigen_numpy_tensor.cpp:120-156)"
                            % (name, array.dtype, array.flags.c_contiguous))
        return array
    return np.ascontiguousarray(array, dtype=np.float32)


def fptr(array):
    return array.ctypes.data_as(c_float_p)


def is_torch_cuda(obj):
    return type(obj).__module__.startswith("torch") and hasattr(obj, "is_cuda") and obj.is_cuda


def current_stream_handle():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def host_stream_handle():
    """Stream for calls with numpy (LSF_HOST) arguments: torch's current stream when torch.cuda is in use in this
    process -- worker threads that set their own stream (multigpu.optimize_pairs, multipair.run_multipair) then really
    run side by side -- else the default stream."""
    import sys
    torch = sys.modules.get("torch")
    if torch is not None and torch.cuda.is_available() and torch.cuda.is_initialized():
        return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    return ctypes.c_void_p(0)


PINNED_RESULT_BYTES = 8 << 20


def result_array(shape):
    """A fresh float32 numpy array for a result (the reference returns new arrays, SURVEY.md 8b "Ownership"). Large
    results are backed by page-locked memory from torch's caching host allocator when torch is loaded: the device-to-host
    copy then needs no bounce through a staging buffer and no first-touch page faults (201 MB for a 256^3 warp field),
    and the block goes back to the allocator when the array is garbage-collected. LSF_PINNED_RESULTS=0: plain np.empty."""
    import sys
    count = int(np.prod(shape)) if len(shape) else 1
    torch = sys.modules.get("torch")
    if (count * 4 >= PINNED_RESULT_BYTES and torch is not None and os.environ.get("LSF_PINNED_RESULTS", "1") != "0"
            and torch.cuda.is_available()):
        try:
            return torch.empty(tuple(shape), dtype=torch.float32, pin_memory=True).numpy()
        except RuntimeError:
            pass
    return np.empty(tuple(shape), dtype=np.float32)


def check_device(*tensors):
    """torch CUDA tensors of one call must live on one GPU and that GPU must be the current device: the library
    launches on the current device's stream and allocates its scratch there. Raises ValueError otherwise (wrap the
    call in `with torch.cuda.device(tensor.device):`)."""
    import torch
    devices = {t.device for t in tensors if t is not None and is_torch_cuda(t)}
    if len(devices) != 1:
        raise ValueError("all fields of a call must live on the same CUDA device, got %s" % sorted(map(str, devices)))
    device = devices.pop()
    if device.index != torch.cuda.current_device():
        raise ValueError("the fields live on %s but the current CUDA device is cuda:%d; wrap the call in "
                         "`with torch.cuda.device(%r):`" % (device, torch.cuda.current_device(), str(device)))
    return device
