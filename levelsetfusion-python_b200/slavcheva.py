"""SobolevFusion / KillingFusion optimizers -- host-side mirror of the reference's interfaces for this path:

* ``SobolevOptimizer2d`` with the process-wide parameter singletons ``SharedParameters`` / ``SobolevParameters``
  (reference C++ class exported by cpp/src/python_export/slavcheva_optimizer.cpp:90-137; parameters
  cpp/src/nonrigid_optimization/slavcheva/optimizer2d.hpp:59-78, sobolev_optimizer2d.hpp:39-70);
* ``SlavchevaOptimizer2d`` (reference Python class nonrigid_opt/slavcheva/slavcheva_optimizer2d.py:72-430) with its enums
  ``ComputeMethod``, ``AdaptiveLearningRateMethod``, ``DataTermMethod``, ``SmoothingTermMethod``;
* ``SlavchevaOptimizer3d`` -- the dimensional generalisation used for 3D KillingFusion volumes (the reference has no 3D
  slavcheva optimizer; definition in DESIGN.md, restated by the CPU checker under oracle/);
* free functions ``warp_field_advanced``, ``warp_field_advanced_no_warp_change``, ``data_term_at_location``
  (python_export/slavcheva_optimizer.cpp:64-89).

``optimize(live_field, canonical_field)`` returns the warped live field (note the argument order, opposite to the
hierarchical optimizer's). numpy arguments are staged through the device by the library; torch CUDA tensors are used in
place. Everything runs in liblsf_b200.so (csrc/slavcheva.cu); there is no CPU path.
"""
import ctypes
import enum

import numpy as np

from . import _lib
from . import telemetry


class ComputeMethod(enum.Enum):
    """reference slavcheva_optimizer2d.py:67-69"""
    DIRECT = 0
    VECTORIZED = 1


class AdaptiveLearningRateMethod(enum.Enum):
    """reference slavcheva_optimizer2d.py:43-45 (RMS_PROP is declared but never used by the reference's iteration)"""
    NONE = 0
    RMS_PROP = 1


class DataTermMethod(enum.Enum):
    """reference nonrigid_opt/slavcheva/data_term.py:44-47"""
    BASIC = 0
    THRESHOLDED_FDM = 1
    BASIC_CPP = 2


class SmoothingTermMethod(enum.Enum):
    """reference nonrigid_opt/slavcheva/smoothing_term.py:27-29"""
    TIKHONOV = 0
    KILLING = 1


# reference math_utils/convolution.py:20-26 == sobolev_optimizer2d.hpp:51-61 (size 7, lambda 0.1)
DEFAULT_SOBOLEV_KERNEL = np.array([2.995900285895913839e-04, 4.410949535667896271e-03, 6.571318954229354858e-02,
                                   9.956527948379516602e-01, 6.571318954229354858e-02, 4.410949535667896271e-03,
                                   2.995900285895913839e-04], dtype=np.float32)


class SharedParameters:
    """reference Optimizer2d::SharedParameters singleton (optimizer2d.hpp:31-82)"""
    _instance = None

    def __init__(self):
        self.gradient_descent_rate = 0.1
        self.maximum_warp_length_lower_threshold = 0.1
        self.maximum_warp_length_upper_threshold = 10000.0
        self.maximum_iteration_count = 100
        self.minimum_iteration_count = 1
        self.enable_focus_spot_analytics = False
        self.enable_convergence_reporting = False
        self.enable_live_sdf_progression_logging = False
        self.enable_gradient_logging = False
        self.enable_gradient_component_logging = False
        self.enable_warp_statistics_logging = False
        self.focus_spot = telemetry.Vector2i(0, 0)

    @classmethod
    def get_instance(cls):
        if cls._instance is None:
            cls._instance = cls()
        return cls._instance


class SobolevParameters:
    """reference SobolevOptimizer2d::SobolevParameters singleton (sobolev_optimizer2d.hpp:39-70)"""
    _instance = None

    def __init__(self):
        self._sobolev_kernel = DEFAULT_SOBOLEV_KERNEL.copy()
        self.smoothing_term_weight = 0.2

    @classmethod
    def get_instance(cls):
        if cls._instance is None:
            cls._instance = cls()
        return cls._instance

    def get_sobolev_kernel(self):
        return self._sobolev_kernel.copy()

    def set_sobolev_kernel(self, sobolev_kernel):
        self._sobolev_kernel = np.ascontiguousarray(np.asarray(sobolev_kernel).ravel(), dtype=np.float32)


def _pointer(array_or_tensor):
    if isinstance(array_or_tensor, np.ndarray):
        return _lib.fptr(array_or_tensor)
    return ctypes.cast(ctypes.c_void_p(array_or_tensor.data_ptr()), _lib.c_float_p)


class _Result:
    pass


def _run(nd, live_field, canonical_field, semantics, data_term_method, smoothing_term_method, level_set_term_enabled,
         sobolev_smoothing_enabled, gradient_descent_rate, data_term_weight, smoothing_term_weight,
         isomorphic_enforcement_factor, level_set_term_weight, lower, upper, maximum_iteration_count,
         minimum_iteration_count, sobolev_kernel, collect_statistics=False, capture_iterations=0,
         log_iteration_statistics=False, log_energies=False):
    """One lsf_slavcheva_optimize[_logged] call. Returns an object with live, warp, report, max_warps, captured,
    iteration_statistics (one WarpDeltaStatistics per iteration when log_iteration_statistics is set) and energies
    ([iterations][3] float64: data, smoothing, level-set energy of every iteration when log_energies is set)."""
    on_device = _lib.is_torch_cuda(live_field) or _lib.is_torch_cuda(canonical_field)
    if on_device:
        import torch
        if not (_lib.is_torch_cuda(live_field) and _lib.is_torch_cuda(canonical_field)):
            raise ValueError("live_field and canonical_field must live on the same device")
        _lib.check_device(live_field, canonical_field)
        live = live_field.contiguous().float()
        canonical = canonical_field.contiguous().float()
    else:
        live = _lib.as_f32(live_field)
        canonical = _lib.as_f32(canonical_field)
    shape = tuple(int(d) for d in live.shape)
    if len(shape) != nd or tuple(canonical.shape) != shape or (nd == 2 and shape[0] != shape[1]):
        # reference slavcheva_optimizer2d.py:157-161
        raise ValueError("warp field, warped live field, and canonical field all need to be square arrays of the same "
                         "size, got %s and %s" % (tuple(live.shape), tuple(canonical.shape)))
    params = _lib.SlavchevaParams()
    params.semantics = int(semantics)
    params.data_term_method = int(data_term_method)
    params.smoothing_term_method = int(smoothing_term_method)
    params.level_set_term_enabled = int(bool(level_set_term_enabled))
    params.sobolev_smoothing_enabled = int(bool(sobolev_smoothing_enabled))
    params.gradient_descent_rate = gradient_descent_rate
    params.data_term_weight = data_term_weight
    params.smoothing_term_weight = smoothing_term_weight
    params.isomorphic_enforcement_factor = isomorphic_enforcement_factor
    params.level_set_term_weight = level_set_term_weight
    params.maximum_warp_length_lower_threshold = lower
    params.maximum_warp_length_upper_threshold = min(float(upper), 3.0e38)
    params.maximum_iteration_count = int(maximum_iteration_count)
    params.minimum_iteration_count = int(minimum_iteration_count)
    kernel = None
    if sobolev_kernel is not None and len(sobolev_kernel) > 0:
        kernel = np.ascontiguousarray(np.asarray(sobolev_kernel).ravel(), dtype=np.float32)
        params.sobolev_kernel = _lib.fptr(kernel)
        params.sobolev_kernel_size = int(kernel.size)
    else:
        params.sobolev_kernel = None
        params.sobolev_kernel_size = 0
    capacity = max(int(maximum_iteration_count), int(minimum_iteration_count), 1)
    max_warps = np.zeros(capacity, dtype=np.float32)
    capture = _lib.IterationCapture()
    capture.level = 0
    capture.max_iterations = int(capture_iterations)
    capture_buffer = None
    if on_device:
        import torch
        live_out = torch.empty(shape, dtype=torch.float32, device=live.device)
        warp_out = torch.empty(shape + (nd,), dtype=torch.float32, device=live.device)
        if capture_iterations > 0:
            capture_buffer = torch.zeros((capture_iterations,) + shape + (nd,), dtype=torch.float32, device=live.device)
        kind, stream = _lib.LSF_DEVICE, _lib.current_stream_handle()
    else:
        live_out = _lib.result_array(shape)
        warp_out = _lib.result_array(shape + (nd,))
        if capture_iterations > 0:
            capture_buffer = np.zeros((capture_iterations,) + shape + (nd,), dtype=np.float32)
        kind, stream = _lib.LSF_HOST, _lib.host_stream_handle()
    if capture_buffer is not None:
        capture.buffer = _pointer(capture_buffer)
    report = _lib.SlavchevaReport()
    dims = (ctypes.c_int * nd)(*shape)
    iteration_statistics = (_lib.WarpDeltaStatisticsRaw * capacity)() if log_iteration_statistics else None
    energies = np.zeros((capacity, 3), dtype=np.float64) if log_energies else None
    _lib.check(_lib.load().lsf_slavcheva_optimize_logged(
        ctypes.byref(params), _pointer(live), _pointer(canonical), nd, dims, _pointer(live_out), _pointer(warp_out), kind,
        ctypes.byref(report), int(bool(collect_statistics)), _lib.fptr(max_warps), capacity, ctypes.byref(capture),
        iteration_statistics, capacity if log_iteration_statistics else 0,
        None if energies is None else energies.ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
        capacity if log_energies else 0, stream))
    result = _Result()
    result.energies = None if energies is None else energies[:int(report.iteration_count)].copy()
    statistics_class = telemetry.WarpDeltaStatistics2d if nd == 2 else telemetry.WarpDeltaStatistics3d
    result.iteration_statistics = [] if iteration_statistics is None else \
        [statistics_class._from_raw(iteration_statistics[i]) for i in range(int(report.iteration_count))]
    result.live = live_out
    result.warp = warp_out
    result.iteration_count = int(report.iteration_count)
    result.iteration_limit_reached = bool(report.iteration_limit_reached)
    result.max_warps = max_warps[:report.iteration_count].copy()
    result.captured = None if capture_buffer is None else capture_buffer[:capture.count]
    report_class = telemetry.ConvergenceReport2d if nd == 2 else telemetry.ConvergenceReport3d
    if report.has_statistics:
        result.report = report_class(report.iteration_count, bool(report.iteration_limit_reached),
                                     report_class._warp_class._from_raw(report.warp_delta_statistics),
                                     report_class._diff_class._from_raw(report.tsdf_difference_statistics))
    else:
        result.report = report_class(report.iteration_count, bool(report.iteration_limit_reached))
    result.report.max_update_length = float(report.last_max_warp_length)
    result.report.dims = shape
    return result


class SobolevOptimizer2d:
    """reference `level_set_fusion_optimization.SobolevOptimizer2d` (sobolev_optimizer2d.cpp:71-138): data term +
    Tikhonov term inside the narrow-band union, Sobolev filter, masked re-warp; parameters come from the
    SharedParameters / SobolevParameters singletons at the time of the call."""
    _nd = 2

    def __init__(self):
        self._report = telemetry.ConvergenceReport2d()
        self._warp_statistics = []
        self._last = None

    def optimize(self, live_field, canonical_field, capture_iterations=0):
        shared = SharedParameters.get_instance()
        sobolev = SobolevParameters.get_instance()
        self._report = telemetry.ConvergenceReport2d()  # clean_out_logs, sobolev_optimizer2d.cpp:164-167
        self._warp_statistics = []
        result = _run(self._nd, live_field, canonical_field, _lib.LSF_SEMANTICS_CPP, DataTermMethod.BASIC.value,
                      SmoothingTermMethod.TIKHONOV.value, False, True, shared.gradient_descent_rate, 1.0,
                      sobolev.smoothing_term_weight, 0.1, 0.0, shared.maximum_warp_length_lower_threshold,
                      shared.maximum_warp_length_upper_threshold, shared.maximum_iteration_count,
                      shared.minimum_iteration_count, sobolev.get_sobolev_kernel(),
                      collect_statistics=shared.enable_convergence_reporting,
                      capture_iterations=max(int(capture_iterations), 0),
                      log_iteration_statistics=shared.enable_warp_statistics_logging)
        if shared.enable_convergence_reporting:
            self._report = result.report
        self._warp_statistics = result.iteration_statistics
        self._last = result
        return result.live

    def get_convergence_report(self):
        return self._report

    def get_warp_statistics_as_matrix(self):
        """reference sobolev_optimizer2d.cpp:144-160: one row of WarpDeltaStatistics2d.to_array() per iteration of the
        last optimize() call, filled when SharedParameters.enable_warp_statistics_logging is set. The reference sizes the
        matrix for 6 columns, to_array() declares 7 and streams 9 values (warp_delta_statistics.tpp:55-68): the nine
        values are returned here -- ratio above the minimum threshold, minimum, maximum, mean and standard deviation of
        the warp lengths, location of the longest warp (x, y), is-largest-below-minimum, is-largest-above-maximum."""
        rows = [s.to_array() for s in self._warp_statistics]
        return np.array(rows, dtype=np.float32).reshape(len(rows), 9)

    # extensions used by the parity tests
    def get_last_warp_field(self):
        return None if self._last is None else self._last.warp

    def get_iteration_count(self):
        return 0 if self._last is None else self._last.iteration_count

    def get_max_warps(self):
        return None if self._last is None else self._last.max_warps

    def get_captured_warps(self):
        return None if self._last is None else self._last.captured


class OptimizationLog:
    """reference slavcheva_optimizer2d.py:58-64: per-iteration maximum warp lengths and the data / smoothing / level-set
    energy aggregates (:370-374; GPU reductions, lsf_slavcheva_optimize_logged)"""

    def __init__(self):
        self.data_energies = []
        self.smoothing_energies = []
        self.level_set_energies = []
        self.max_warps = []
        self.convergence_report = telemetry.ConvergenceReport2d()


class SlavchevaOptimizer2d:
    """reference Python class `SlavchevaOptimizer2d` (slavcheva_optimizer2d.py:72-430). compute_method selects which
    of the reference's two iterations is reproduced (DIRECT: per-voxel loop with Killing / level-set /
    thresholded-FDM support; VECTORIZED: Tikhonov only). Like the reference, ``optimize`` returns the warped live
    field AND overwrites a numpy ``live_field`` argument with it. Visualisation / plotting arguments are accepted and
    ignored (out of scope)."""
    _nd = 2

    def __init__(self, out_path="out2D", field_size=128, default_value=1.0, compute_method=ComputeMethod.DIRECT,
                 level_set_term_enabled=False, sobolev_smoothing_enabled=False,
                 data_term_method=DataTermMethod.BASIC, smoothing_term_method=SmoothingTermMethod.TIKHONOV,
                 adaptive_learning_rate_method=AdaptiveLearningRateMethod.NONE, gradient_descent_rate=0.1,
                 data_term_weight=1.0, smoothing_term_weight=0.2, isomorphic_enforcement_factor=0.1,
                 level_set_term_weight=0.2, maximum_warp_length_lower_threshold=0.1,
                 maximum_warp_length_upper_threshold=10000, max_iterations=100, min_iterations=1, sobolev_kernel=None,
                 visualization_settings=None, enable_convergence_status_logging=True, log_energies=True):
        # log_energies (extension): the reference always fills OptimizationLog's energy lists; False skips the per-iteration
        # energy reductions and lets small fields take the single-launch path
        self.log_energies = bool(log_energies)
        self.out_path = out_path
        self.field_size = field_size
        self.default_value = default_value
        self.compute_method = compute_method
        self.level_set_term_enabled = level_set_term_enabled
        self.sobolev_smoothing_enabled = sobolev_smoothing_enabled
        self.data_term_method = data_term_method
        self.smoothing_term_method = smoothing_term_method
        self.adaptive_learning_rate_method = adaptive_learning_rate_method
        self.gradient_descent_rate = gradient_descent_rate
        self.data_term_weight = data_term_weight
        self.smoothing_term_weight = smoothing_term_weight
        self.isomorphic_enforcement_factor = isomorphic_enforcement_factor
        self.level_set_term_weight = level_set_term_weight
        self.maximum_warp_length_lower_threshold = maximum_warp_length_lower_threshold
        self.maximum_warp_length_upper_threshold = maximum_warp_length_upper_threshold
        self.max_iterations = max_iterations
        self.min_iterations = min_iterations
        self.sobolev_kernel = sobolev_kernel
        self.visualization_settings = visualization_settings
        self.enable_convergence_status_logging = enable_convergence_status_logging
        self.log = None
        self._last = None

    def _semantics(self):
        name = getattr(self.compute_method, "name", str(self.compute_method))
        return _lib.LSF_SEMANTICS_PY_VECTORIZED if name == "VECTORIZED" else _lib.LSF_SEMANTICS_PY_DIRECT

    @staticmethod
    def _enum_value(member, basic_cpp_as=None):
        name = getattr(member, "name", None)
        if name == "BASIC_CPP":
            return DataTermMethod.BASIC.value if basic_cpp_as is None else basic_cpp_as
        return int(getattr(member, "value", member))

    def optimize(self, live_field, canonical_field, capture_iterations=0):
        kernel = self.sobolev_kernel if self.sobolev_kernel is not None else DEFAULT_SOBOLEV_KERNEL
        result = _run(self._nd, live_field, canonical_field, self._semantics(),
                      self._enum_value(self.data_term_method), self._enum_value(self.smoothing_term_method),
                      self.level_set_term_enabled, self.sobolev_smoothing_enabled, self.gradient_descent_rate,
                      self.data_term_weight, self.smoothing_term_weight, self.isomorphic_enforcement_factor,
                      self.level_set_term_weight, self.maximum_warp_length_lower_threshold,
                      self.maximum_warp_length_upper_threshold, self.max_iterations, self.min_iterations, kernel,
                      collect_statistics=self.enable_convergence_status_logging, capture_iterations=capture_iterations,
                      log_energies=self.log_energies)
        self.log = OptimizationLog()
        self.log.max_warps = [float(v) for v in result.max_warps]
        if result.energies is not None:  # reference :370-374
            self.log.data_energies = [float(v) for v in result.energies[:, 0]]
            self.log.smoothing_energies = [float(v) for v in result.energies[:, 1]]
            self.log.level_set_energies = [float(v) for v in result.energies[:, 2]]
        if self.enable_convergence_status_logging:
            self.log.convergence_report = result.report
        self._last = result
        if isinstance(live_field, np.ndarray) and live_field.dtype == np.float32:
            np.copyto(live_field, result.live)  # the reference warps its argument in place (:324-328,408)
            return live_field
        return result.live

    def get_convergence_report(self):
        return self.log.convergence_report

    def get_last_warp_field(self):
        return None if self._last is None else self._last.warp

    def get_iteration_count(self):
        return 0 if self._last is None else self._last.iteration_count

    def get_captured_warps(self):
        return None if self._last is None else self._last.captured


class SlavchevaOptimizer3d:
    """3D SobolevFusion / KillingFusion optimizer: the reference's C++ SobolevOptimizer2d loop (band-union masks,
    per-pass preserve-zeros Sobolev filter, masked trilinear re-warp with truncation snap, maximum warp measured after
    the re-warp) generalised to volumes, with the Killing and level-set terms of the reference's Python code available
    as options. Component c of the warp displaces along array axis c. Keyword names follow SlavchevaOptimizer2d."""
    _nd = 3

    def __init__(self, level_set_term_enabled=False, sobolev_smoothing_enabled=True,
                 data_term_method=DataTermMethod.BASIC, smoothing_term_method=SmoothingTermMethod.TIKHONOV,
                 gradient_descent_rate=0.1, data_term_weight=1.0, smoothing_term_weight=0.2,
                 isomorphic_enforcement_factor=0.1, level_set_term_weight=0.2, maximum_warp_length_lower_threshold=0.1,
                 maximum_warp_length_upper_threshold=10000, max_iterations=100, min_iterations=1, sobolev_kernel=None,
                 enable_convergence_status_logging=False):
        self.level_set_term_enabled = level_set_term_enabled
        self.sobolev_smoothing_enabled = sobolev_smoothing_enabled
        self.data_term_method = data_term_method
        self.smoothing_term_method = smoothing_term_method
        self.gradient_descent_rate = gradient_descent_rate
        self.data_term_weight = data_term_weight
        self.smoothing_term_weight = smoothing_term_weight
        self.isomorphic_enforcement_factor = isomorphic_enforcement_factor
        self.level_set_term_weight = level_set_term_weight
        self.maximum_warp_length_lower_threshold = maximum_warp_length_lower_threshold
        self.maximum_warp_length_upper_threshold = maximum_warp_length_upper_threshold
        self.max_iterations = max_iterations
        self.min_iterations = min_iterations
        self.sobolev_kernel = sobolev_kernel
        self.enable_convergence_status_logging = enable_convergence_status_logging
        self._last = None

    def optimize(self, live_field, canonical_field, capture_iterations=0):
        kernel = self.sobolev_kernel if self.sobolev_kernel is not None else DEFAULT_SOBOLEV_KERNEL
        self._last = _run(self._nd, live_field, canonical_field, _lib.LSF_SEMANTICS_CPP,
                          SlavchevaOptimizer2d._enum_value(self.data_term_method),
                          SlavchevaOptimizer2d._enum_value(self.smoothing_term_method), self.level_set_term_enabled,
                          self.sobolev_smoothing_enabled, self.gradient_descent_rate, self.data_term_weight,
                          self.smoothing_term_weight, self.isomorphic_enforcement_factor, self.level_set_term_weight,
                          self.maximum_warp_length_lower_threshold, self.maximum_warp_length_upper_threshold,
                          self.max_iterations, self.min_iterations, kernel,
                          collect_statistics=self.enable_convergence_status_logging,
                          capture_iterations=capture_iterations)
        return self._last.live

    def get_convergence_report(self):
        return self._last.report

    def get_last_warp_field(self):
        return self._last.warp

    def get_iteration_count(self):
        return self._last.iteration_count

    def get_max_warps(self):
        return self._last.max_warps

    def get_captured_warps(self):
        return self._last.captured


class SlavchevaOptimizer2dCpp(SlavchevaOptimizer3d):
    """2D twin of SlavchevaOptimizer3d: the C++ loop semantics with every term option (used by the parity tests to
    tie the 3D generalisation to the 2D code path; SobolevOptimizer2d is this class with the singleton parameters)."""
    _nd = 2


# ------------------------------------------------------------------------------------------------ free functions
def _warp_advanced(warped_live_field, canonical_field, warp_field_u, warp_field_v, band_union_only, known_values_only,
                   substitute_original, truncation_float_threshold, modify_warp):
    live = _lib.as_f32(warped_live_field)
    canonical = _lib.as_f32(canonical_field)
    warp = np.ascontiguousarray(np.stack([np.asarray(warp_field_u), np.asarray(warp_field_v)], axis=-1),
                                dtype=np.float32)
    if live.ndim != 2 or live.shape != canonical.shape or warp.shape[:2] != live.shape:
        raise ValueError("fields do not match: %s %s %s" % (live.shape, canonical.shape, warp.shape))
    out = np.empty_like(live)
    dims = (ctypes.c_int * 2)(*live.shape)
    _lib.check(_lib.load().lsf_warp_advanced(_lib.fptr(live), _lib.fptr(canonical), _lib.fptr(warp), 2, dims,
                                             int(bool(band_union_only)), int(bool(known_values_only)),
                                             int(bool(substitute_original)),
                                             ctypes.c_float(truncation_float_threshold), int(modify_warp),
                                             _lib.fptr(out), _lib.LSF_HOST, _lib.host_stream_handle()))
    return out, warp


def warp_field_advanced(warped_live_field, canonical_field, warp_field_u, warp_field_v, band_union_only=False,
                        known_values_only=False, substitute_original=False, truncation_float_threshold=1e-6):
    """reference py_warp_field_advanced (field_warping.cpp:156-170) -> (new live field, (u, v))"""
    out, warp = _warp_advanced(warped_live_field, canonical_field, warp_field_u, warp_field_v, band_union_only,
                               known_values_only, substitute_original, truncation_float_threshold, 1)
    return out, (np.ascontiguousarray(warp[..., 0]), np.ascontiguousarray(warp[..., 1]))


def warp_field_advanced_no_warp_change(warped_live_field, canonical_field, warp_field_u, warp_field_v,
                                       band_union_only=False, known_values_only=False, substitute_original=False,
                                       truncation_float_threshold=1e-6):
    """reference py_warp_field_advanced_no_warp_change (field_warping.cpp:172-183) -> new live field"""
    out, _ = _warp_advanced(warped_live_field, canonical_field, warp_field_u, warp_field_v, band_union_only,
                            known_values_only, substitute_original, truncation_float_threshold, 0)
    return out


def warp_advanced(live_field, canonical_field, warp_field, band_union_only=False, known_values_only=False,
                  substitute_original=False, truncation_float_threshold=1e-6, modify_warp=True):
    """Dimension-generic form (2D [H,W,2] / 3D [X,Y,Z,3] interleaved warp) -> (new live field, warp field)"""
    live = _lib.as_f32(live_field)
    canonical = _lib.as_f32(canonical_field)
    warp = _lib.as_f32(warp_field).copy()
    nd = live.ndim
    out = np.empty_like(live)
    dims = (ctypes.c_int * nd)(*live.shape)
    _lib.check(_lib.load().lsf_warp_advanced(_lib.fptr(live), _lib.fptr(canonical), _lib.fptr(warp), nd, dims,
                                             int(bool(band_union_only)), int(bool(known_values_only)),
                                             int(bool(substitute_original)),
                                             ctypes.c_float(truncation_float_threshold), int(bool(modify_warp)),
                                             _lib.fptr(out), _lib.LSF_HOST, _lib.host_stream_handle()))
    return out, warp


def data_term_at_location(warped_live_field, canonical_field, x, y, live_gradient_x_field, live_gradient_y_field):
    """reference py_data_term_at_location (data_term.cpp:135-150, compute_local_data_term_gradient :41-60): a scalar
    host-side convenience of the reference API -> (array [gradient x, gradient y], energy contribution). Like the
    reference it addresses the matrices as (x, y) = (row, column) (data_term.cpp:49-54)."""
    live = np.float32(warped_live_field[x, y])
    canonical = np.float32(canonical_field[x, y])
    difference = np.float32(live - canonical)
    scaling_factor = np.float32(10.0)
    gradient_x = np.float32(np.float32(difference * np.float32(live_gradient_x_field[x, y])) * scaling_factor)
    gradient_y = np.float32(np.float32(difference * np.float32(live_gradient_y_field[x, y])) * scaling_factor)
    energy = np.float32(np.float32(np.float32(0.5) * difference) * difference)
    return np.array([gradient_x, gradient_y], dtype=np.float32), float(energy)
