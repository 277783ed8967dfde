"""Telemetry value types and builders of the reference's extension module (cpp/src/python_export/telemetry.tpp:50-145,
cpp/src/python_export/math.cpp:36-70): Vector2i/3i/2f, WarpDeltaStatistics{2d,3d}, TsdfDifferenceStatistics{2d,3d},
ConvergenceReport{2d,3d}, build_warp_delta_statistics_{2d,3d}, build_tsdf_difference_statistics_{2d,3d},
mean_vector_length. Equality is the reference's `almost_equal` (absolute tolerance 3e-6,
cpp/src/math/almost_equal.tpp:56); the statistics are computed by GPU reductions (csrc/slavcheva.cuh)."""
import ctypes

import numpy as np

from . import _lib

_TOLERANCE = 3e-6


def _close(a, b):
    return abs(float(a) - float(b)) < _TOLERANCE


class Vector2i:
    def __init__(self, x=0, y=None):
        self.x = int(x)
        self.y = int(x if y is None else y)

    u = property(lambda self: self.x, lambda self, value: setattr(self, "x", int(value)))
    v = property(lambda self: self.y, lambda self, value: setattr(self, "y", int(value)))

    def __eq__(self, other):
        return isinstance(other, Vector2i) and (self.x, self.y) == (other.x, other.y)

    def __ne__(self, other):
        return not self == other

    def __iter__(self):
        return iter((self.x, self.y))

    def __repr__(self):
        return "Vector2i(%d, %d)" % (self.x, self.y)

    def __str__(self):
        return "%d, %d" % (self.x, self.y)


class Vector3i:
    def __init__(self, x=0, y=None, z=None):
        self.x = int(x)
        self.y = int(x if y is None else y)
        self.z = int(x if z is None else z)

    def __eq__(self, other):
        return isinstance(other, Vector3i) and (self.x, self.y, self.z) == (other.x, other.y, other.z)

    def __ne__(self, other):
        return not self == other

    def __iter__(self):
        return iter((self.x, self.y, self.z))

    def __repr__(self):
        return "Vector3i(%d, %d, %d)" % (self.x, self.y, self.z)

    def __str__(self):
        return "%d, %d, %d" % (self.x, self.y, self.z)


class Vector2f:
    def __init__(self, x=0.0, y=None):
        self.x = float(x)
        self.y = float(x if y is None else y)

    u = property(lambda self: self.x, lambda self, value: setattr(self, "x", float(value)))
    v = property(lambda self: self.y, lambda self, value: setattr(self, "y", float(value)))

    def __repr__(self):
        return "Vector2f(%g, %g)" % (self.x, self.y)


def _coordinates(nd, values=None):
    if values is None:
        return Vector2i(0) if nd == 2 else Vector3i(0)
    values = [int(v) for v in values]
    return Vector2i(values[0], values[1]) if nd == 2 else Vector3i(values[0], values[1], values[2])


class _WarpDeltaStatistics:
    _nd = 0

    def __init__(self, ratio_above_min_threshold=0.0, length_min=0.0, length_max=0.0, length_mean=0.0,
                 length_standard_deviation=0.0, longest_warp_location=None, is_largest_below_min_threshold=False,
                 is_largest_above_max_threshold=False):
        self.ratio_above_min_threshold = float(ratio_above_min_threshold)
        self.length_min = float(length_min)
        self.length_max = float(length_max)
        self.length_mean = float(length_mean)
        self.length_standard_deviation = float(length_standard_deviation)
        self.longest_warp_location = longest_warp_location if longest_warp_location is not None \
            else _coordinates(self._nd)
        self.is_largest_below_min_threshold = bool(is_largest_below_min_threshold)
        self.is_largest_above_max_threshold = bool(is_largest_above_max_threshold)

    @classmethod
    def _from_raw(cls, raw):
        return cls(raw.ratio_above_min_threshold, raw.length_min, raw.length_max, raw.length_mean,
                   raw.length_standard_deviation, _coordinates(cls._nd, raw.longest_warp_location),
                   bool(raw.is_largest_below_min_threshold), bool(raw.is_largest_above_max_threshold))

    def to_array(self):
        """reference WarpDeltaStatistics::to_array (warp_delta_statistics.tpp:55-68): the vector is declared with 7
        entries and nine values are streamed into it; all nine are returned"""
        return np.array([self.ratio_above_min_threshold, self.length_min, self.length_max, self.length_mean,
                         self.length_standard_deviation, float(self.longest_warp_location.x),
                         float(self.longest_warp_location.y), float(self.is_largest_below_min_threshold),
                         float(self.is_largest_above_max_threshold)], dtype=np.float32)

    def __eq__(self, other):
        return (isinstance(other, _WarpDeltaStatistics)
                and _close(self.ratio_above_min_threshold, other.ratio_above_min_threshold)
                and _close(self.length_min, other.length_min) and _close(self.length_max, other.length_max)
                and _close(self.length_mean, other.length_mean)
                and _close(self.length_standard_deviation, other.length_standard_deviation)
                and self.longest_warp_location == other.longest_warp_location
                and self.is_largest_below_min_threshold == other.is_largest_below_min_threshold
                and self.is_largest_above_max_threshold == other.is_largest_above_max_threshold)

    def __ne__(self, other):
        return not self == other

    def __str__(self):
        return ("[warp delta stats]\n  ratio above min threshold: %g\n  min: %g\n  max: %g\n  mean: %g\n  std: %g\n"
                "  longest warp at: (%s)\n  largest below min threshold: %d\n  largest above max threshold: %d"
                % (self.ratio_above_min_threshold, self.length_min, self.length_max, self.length_mean,
                   self.length_standard_deviation, self.longest_warp_location, self.is_largest_below_min_threshold,
                   self.is_largest_above_max_threshold))

    __repr__ = __str__


class WarpDeltaStatistics2d(_WarpDeltaStatistics):
    _nd = 2


class WarpDeltaStatistics3d(_WarpDeltaStatistics):
    _nd = 3


class _TsdfDifferenceStatistics:
    _nd = 0

    def __init__(self, difference_min=0.0, difference_max=0.0, difference_mean=0.0, difference_standard_deviation=0.0,
                 biggest_difference_location=None):
        self.difference_min = float(difference_min)
        self.difference_max = float(difference_max)
        self.difference_mean = float(difference_mean)
        self.difference_standard_deviation = float(difference_standard_deviation)
        self.biggest_difference_location = biggest_difference_location if biggest_difference_location is not None \
            else _coordinates(self._nd)

    @classmethod
    def _from_raw(cls, raw):
        return cls(raw.difference_min, raw.difference_max, raw.difference_mean, raw.difference_standard_deviation,
                   _coordinates(cls._nd, raw.biggest_difference_location))

    def to_array(self):
        return np.array([self.difference_min, self.difference_max, self.difference_mean,
                         self.difference_standard_deviation, float(self.biggest_difference_location.x),
                         float(self.biggest_difference_location.y)], dtype=np.float32)

    def __eq__(self, other):
        return (isinstance(other, _TsdfDifferenceStatistics) and _close(self.difference_min, other.difference_min)
                and _close(self.difference_max, other.difference_max)
                and _close(self.difference_mean, other.difference_mean)
                and _close(self.difference_standard_deviation, other.difference_standard_deviation)
                and self.biggest_difference_location == other.biggest_difference_location)

    def __ne__(self, other):
        return not self == other

    def __str__(self):
        return ("[tsdf diff stats]\n  min: %g\n  max: %g\n  mean: %g\n  std: %g\n  greatest diff at: (%s)"
                % (self.difference_min, self.difference_max, self.difference_mean, self.difference_standard_deviation,
                   self.biggest_difference_location))

    __repr__ = __str__


class TsdfDifferenceStatistics2d(_TsdfDifferenceStatistics):
    _nd = 2


class TsdfDifferenceStatistics3d(_TsdfDifferenceStatistics):
    _nd = 3


class _ConvergenceReport:
    _nd = 0
    _warp_class = None
    _diff_class = None

    def __init__(self, iteration_count=0, iteration_limit_reached=False, warp_delta_statistics=None,
                 tsdf_difference_statistics=None):
        self.iteration_count = int(iteration_count)
        self.iteration_limit_reached = bool(iteration_limit_reached)
        self.warp_delta_statistics = warp_delta_statistics if warp_delta_statistics is not None \
            else self._warp_class()
        self.tsdf_difference_statistics = tsdf_difference_statistics if tsdf_difference_statistics is not None \
            else self._diff_class()
        # extensions of this implementation (not part of the reference type, ignored by ==)
        self.max_update_length = float("nan")
        self.dims = ()

    def __eq__(self, other):
        return (isinstance(other, _ConvergenceReport) and self.iteration_count == other.iteration_count
                and self.iteration_limit_reached == other.iteration_limit_reached
                and self.warp_delta_statistics == other.warp_delta_statistics
                and self.tsdf_difference_statistics == other.tsdf_difference_statistics)

    def __ne__(self, other):
        return not self == other

    def __str__(self):
        return ("===[convergence report]===\n  iter count: %d\n  limit reached: %d\n--------------------------\n%s\n"
                "--------------------------\n%s\n==========================)"
                % (self.iteration_count, self.iteration_limit_reached, self.warp_delta_statistics,
                   self.tsdf_difference_statistics))

    __repr__ = __str__


class ConvergenceReport2d(_ConvergenceReport):
    _nd = 2
    _warp_class = WarpDeltaStatistics2d
    _diff_class = TsdfDifferenceStatistics2d


class ConvergenceReport3d(_ConvergenceReport):
    _nd = 3
    _warp_class = WarpDeltaStatistics3d
    _diff_class = TsdfDifferenceStatistics3d


def _fields(nd, *arrays):
    """numpy (host) or torch CUDA (device) arguments -> (memory kind, stream, pointers, dims)"""
    on_device = any(_lib.is_torch_cuda(a) for a in arrays)
    if on_device:
        if not all(_lib.is_torch_cuda(a) for a in arrays):
            raise ValueError("all fields of a call must be torch CUDA tensors, or all numpy arrays")
        _lib.check_device(*arrays)
        converted = [a.contiguous().float() for a in arrays]
        pointers = [ctypes.cast(ctypes.c_void_p(a.data_ptr()), _lib.c_float_p) for a in converted]
        kind, stream = _lib.LSF_DEVICE, _lib.current_stream_handle()
    else:
        converted = [_lib.as_f32(a) for a in arrays]
        pointers = [_lib.fptr(a) for a in converted]
        kind, stream = _lib.LSF_HOST, _lib.host_stream_handle()
    shape = tuple(int(d) for d in converted[-1].shape[:nd])
    return kind, stream, pointers, (ctypes.c_int * nd)(*shape), converted


def _build_warp_delta_statistics(nd, warp_field, canonical_field, live_field, min_threshold, max_threshold):
    kind, stream, (warp, canonical, live), dims, keep = _fields(nd, warp_field, canonical_field, live_field)
    if tuple(keep[0].shape) != tuple(keep[1].shape) + (nd,) or tuple(keep[1].shape) != tuple(keep[2].shape):
        raise RuntimeError("Dimensions of one of the input matrices don't appear to match.")
    raw = _lib.WarpDeltaStatisticsRaw()
    _lib.check(_lib.load().lsf_warp_delta_statistics(warp, canonical, live, nd, dims, ctypes.c_float(min_threshold),
                                                     ctypes.c_float(max_threshold), ctypes.byref(raw), kind, stream))
    return (WarpDeltaStatistics2d if nd == 2 else WarpDeltaStatistics3d)._from_raw(raw)


def _build_tsdf_difference_statistics(nd, canonical_tsdf, live_tsdf):
    kind, stream, (canonical, live), dims, keep = _fields(nd, canonical_tsdf, live_tsdf)
    if tuple(keep[0].shape) != tuple(keep[1].shape):
        raise RuntimeError("Dimensions of one of the input matrices don't appear to match.")
    raw = _lib.TsdfDifferenceStatisticsRaw()
    _lib.check(_lib.load().lsf_tsdf_difference_statistics(canonical, live, nd, dims, ctypes.byref(raw), kind, stream))
    return (TsdfDifferenceStatistics2d if nd == 2 else TsdfDifferenceStatistics3d)._from_raw(raw)


def build_warp_delta_statistics_2d(warp_field, canonical_field, live_field, min_threshold, max_threshold):
    """reference telemetry::build_warp_delta_statistics (warp_delta_statistics.tpp:88-116)"""
    return _build_warp_delta_statistics(2, warp_field, canonical_field, live_field, min_threshold, max_threshold)


def build_warp_delta_statistics_3d(warp_field, canonical_field, live_field, min_threshold, max_threshold):
    return _build_warp_delta_statistics(3, warp_field, canonical_field, live_field, min_threshold, max_threshold)


def build_tsdf_difference_statistics_2d(canonical_tsdf, live_tsdf):
    """reference telemetry::build_tsdf_difference_statistics (tsdf_difference_statistics.tpp:86-97)"""
    return _build_tsdf_difference_statistics(2, canonical_tsdf, live_tsdf)


def build_tsdf_difference_statistics_3d(canonical_tsdf, live_tsdf):
    return _build_tsdf_difference_statistics(3, canonical_tsdf, live_tsdf)


def mean_vector_length(vector_field):
    """reference math::mean_vector_length (cpp/src/math/statistics.cpp:100-109): mean length over ALL vectors. Computed
    with the band-union reduction against two all-zero scalar fields (no voxel is masked)."""
    vector_field = _lib.as_f32(vector_field) if not _lib.is_torch_cuda(vector_field) else vector_field
    nd = vector_field.ndim - 1
    if _lib.is_torch_cuda(vector_field):
        import torch
        zeros = torch.zeros(tuple(vector_field.shape[:nd]), dtype=torch.float32, device=vector_field.device)
    else:
        zeros = np.zeros(vector_field.shape[:nd], dtype=np.float32)
    return _build_warp_delta_statistics(nd, vector_field, zeros, zeros, 0.0, float("inf")).length_mean
