"""HierarchicalOptimizer2d / HierarchicalOptimizer3d -- host-side mirror of the reference's classes
(cpp/src/python_export/hierarchical_optimizer.tpp:40-135; constructor defaults
cpp/src/nonrigid_optimization/hierarchical/optimizer.hpp:51-65; Python twin
nonrigid_opt/hierarchical/hierarchical_optimizer2d.py:62-121).

``optimize(canonical_field, live_field)`` returns the warp field ([H,W,2] / [X,Y,Z,3] float32), exactly like the
reference. numpy arguments are staged through the device by the library (LSF_HOST); torch CUDA tensors are used
in place (LSF_DEVICE) on torch's current stream and a torch tensor is returned.
"""
import ctypes
import enum

import numpy as np

from . import _lib
from . import telemetry


def _report_from_raw(nd, raw, with_statistics):
    """lsf_level_report -> telemetry.ConvergenceReport{2d,3d} (reference telemetry::ConvergenceReport,
    cpp/src/telemetry/convergence_report.hpp:40-44). max_update_length / dims are extensions."""
    cls = telemetry.ConvergenceReport2d if nd == 2 else telemetry.ConvergenceReport3d
    if with_statistics:
        warp = cls._warp_class(raw.warp_ratio_above_min_threshold, raw.warp_length_min, raw.warp_length_max,
                               raw.warp_length_mean, raw.warp_length_std,
                               telemetry._coordinates(nd, raw.warp_longest_location),
                               bool(raw.warp_is_largest_below_min_threshold),
                               bool(raw.warp_is_largest_above_max_threshold))
        diff = cls._diff_class(raw.diff_min, raw.diff_max, raw.diff_mean, raw.diff_std,
                               telemetry._coordinates(nd, raw.diff_biggest_location))
        report = cls(raw.iteration_count, bool(raw.iteration_limit_reached), warp, diff)
    else:
        report = cls(raw.iteration_count, bool(raw.iteration_limit_reached))
    report.max_update_length = float(raw.max_update_length)
    report.dims = tuple(int(d) for d in raw.dims)
    return report


class OptimizationIterationData:
    """reference telemetry::OptimizationIterationData (cpp/src/telemetry/optimization_iteration_data.tpp, exported as
    OptimizationIterationData2d / 3d, python_export/telemetry.tpp:136-145): what one pyramid level's iterations left
    behind when LoggingParameters.collect_per_level_iteration_data is set. Frame i of a level holds the live pyramid
    level, the warp field after iteration i, and the data-term / Tikhonov-term gradients of that iteration (level 0
    starts with one extra frame of zero fields, optimizer_with_telemetry.tpp:90-99)."""

    def __init__(self):
        self._live_fields = []
        self._warp_fields = []
        self._data_term_gradients = []
        self._tikhonov_term_gradients = []

    def add_iteration_result(self, live_field, warp_field, data_term_gradients, tikhonov_term_gradients):
        self._live_fields.append(live_field)
        self._warp_fields.append(warp_field)
        self._data_term_gradients.append(data_term_gradients)
        self._tikhonov_term_gradients.append(tikhonov_term_gradients)

    def get_live_fields(self):
        return list(self._live_fields)

    def get_warp_fields(self):
        return list(self._warp_fields)

    def get_data_term_gradients(self):
        return list(self._data_term_gradients)

    def get_tikhonov_term_gradients(self):
        return list(self._tikhonov_term_gradients)

    def get_frame_count(self):
        return len(self._live_fields)


class OptimizationIterationData2d(OptimizationIterationData):
    pass


class OptimizationIterationData3d(OptimizationIterationData):
    pass


class _HierarchicalOptimizer:
    _nd = 0

    class ResamplingStrategy(enum.IntEnum):
        NEAREST_AND_AVERAGE = 0
        LINEAR = 1

    class VerbosityParameters:
        def __init__(self, print_max_warp_update=False, print_iteration_mean_tsdf_difference=False,
                     print_iteration_std_tsdf_difference=False, print_iteration_data_energy=False,
                     print_iteration_tikhonov_energy=False):
            self.print_iteration_max_warp_update = print_max_warp_update
            self.print_iteration_mean_tsdf_difference = print_iteration_mean_tsdf_difference
            self.print_iteration_std_tsdf_difference = print_iteration_std_tsdf_difference
            self.print_iteration_data_energy = print_iteration_data_energy
            self.print_iteration_tikhonov_energy = print_iteration_tikhonov_energy
            self.print_per_iteration_info = any((print_max_warp_update, print_iteration_mean_tsdf_difference,
                                                 print_iteration_std_tsdf_difference, print_iteration_data_energy,
                                                 print_iteration_tikhonov_energy))
            self.print_per_level_info = self.print_per_iteration_info

    class LoggingParameters:
        def __init__(self, collect_per_level_convergence_reports=False, collect_per_level_iteration_data=False):
            self.collect_per_level_convergence_reports = collect_per_level_convergence_reports
            self.collect_per_level_iteration_data = collect_per_level_iteration_data

    def __init__(self,
                 tikhonov_term_enabled=True,
                 gradient_kernel_enabled=True,
                 maximum_chunk_size=8,
                 rate=0.1,
                 maximum_iteration_count=100,
                 maximum_warp_update_threshold=0.001,
                 data_term_amplifier=1.0,
                 tikhonov_strength=0.2,
                 kernel=None,
                 resampling_strategy=None,
                 verbosity_parameters=None,
                 logging_parameters=None):
        self.tikhonov_term_enabled = bool(tikhonov_term_enabled)
        self.gradient_kernel_enabled = bool(gradient_kernel_enabled)
        self.maximum_chunk_size = int(maximum_chunk_size)
        self.rate = float(rate)
        self.maximum_iteration_count = int(maximum_iteration_count)
        self.maximum_warp_update_threshold = float(maximum_warp_update_threshold)
        self.data_term_amplifier = float(data_term_amplifier)
        self.tikhonov_strength = float(tikhonov_strength)
        self.kernel = None if kernel is None else np.ascontiguousarray(np.asarray(kernel).ravel(), dtype=np.float32)
        if resampling_strategy is None:
            resampling_strategy = self.ResamplingStrategy.NEAREST_AND_AVERAGE
        self.resampling_strategy = self.ResamplingStrategy(int(resampling_strategy))
        self.verbosity_parameters = verbosity_parameters or self.VerbosityParameters()
        self.logging_parameters = logging_parameters or self.LoggingParameters()
        self._reports = []
        self._iteration_data = []

    # ------------------------------------------------------------------ C-ABI plumbing
    def _params(self):
        p = _lib.HierParams()
        p.tikhonov_term_enabled = int(self.tikhonov_term_enabled)
        p.gradient_kernel_enabled = int(self.gradient_kernel_enabled)
        p.maximum_chunk_size = self.maximum_chunk_size
        p.rate = self.rate
        p.maximum_iteration_count = self.maximum_iteration_count
        p.maximum_warp_update_threshold = self.maximum_warp_update_threshold
        p.data_term_amplifier = self.data_term_amplifier
        p.tikhonov_strength = self.tikhonov_strength
        if self.kernel is not None and self.kernel.size > 0:
            p.kernel = _lib.fptr(self.kernel)
            p.kernel_size = int(self.kernel.size)
        else:
            p.kernel = None
            p.kernel_size = 0
        p.resampling_strategy = int(self.resampling_strategy)
        return p

    def optimize(self, canonical_field, live_field, capture_level=-1, capture_iterations=0, out=None):
        """Find the warp that maps `live_field` onto `canonical_field` (reference optimizer.tpp:83-131).

        capture_level / capture_iterations (extension used by the parity tests): additionally keeps the warp
        field after each of the first `capture_iterations` iterations of pyramid level `capture_level`;
        retrieve with get_captured_warps(). `out` (extension): preallocated float32 C-contiguous result buffer
        (e.g. a pinned numpy view) to write the warp field into."""
        nd = self._nd
        on_device = _lib.is_torch_cuda(canonical_field) or _lib.is_torch_cuda(live_field)
        if on_device:
            import torch
            if not (_lib.is_torch_cuda(canonical_field) and _lib.is_torch_cuda(live_field)):
                raise ValueError("canonical_field and live_field must live on the same device")
            _lib.check_device(canonical_field, live_field)
            canonical = canonical_field.contiguous().float()
            live = live_field.contiguous().float()
            shape = tuple(int(d) for d in canonical.shape)
        else:
            canonical = _lib.as_f32(canonical_field)
            live = _lib.as_f32(live_field)
            shape = canonical.shape
        if len(shape) != nd or tuple(live.shape) != tuple(shape):
            raise ValueError("expected two %dD fields of equal shape, got %s and %s"
                             % (nd, tuple(canonical.shape), tuple(live.shape)))
        lib = _lib.load()
        params = self._params()
        reports = (_lib.LevelReport * _lib.LSF_MAX_LEVELS)()
        capture = _lib.IterationCapture()
        capture.level = int(capture_level)
        capture.max_iterations = int(capture_iterations)
        capture_buffer = None
        if capture_level >= 0 and capture_iterations > 0:
            level_count = int(np.log2(self.maximum_chunk_size)) + 1
            shrink = 2 ** (level_count - 1 - capture_level)
            level_shape = tuple(d // shrink for d in shape)
        if on_device:
            import torch
            warp = torch.empty(shape + (nd,), dtype=torch.float32, device=canonical.device)
            kind, stream = _lib.LSF_DEVICE, _lib.current_stream_handle()
            ptr = lambda t: ctypes.cast(ctypes.c_void_p(t.data_ptr()), _lib.c_float_p)
            if capture_level >= 0 and capture_iterations > 0:
                capture_buffer = torch.zeros((capture_iterations,) + level_shape + (nd,), dtype=torch.float32,
                                             device=canonical.device)
                capture.buffer = ptr(capture_buffer)
        else:
            if out is not None:
                if not (isinstance(out, np.ndarray) and out.dtype == np.float32 and out.flags.c_contiguous
                        and out.shape == tuple(shape) + (nd,)):
                    raise ValueError("out must be a C-contiguous float32 array of shape %s" % (tuple(shape) + (nd,),))
                warp = out
            else:
                warp = _lib.result_array(tuple(shape) + (nd,))
            kind, stream = _lib.LSF_HOST, _lib.host_stream_handle()
            ptr = _lib.fptr
            if capture_level >= 0 and capture_iterations > 0:
                capture_buffer = np.zeros((capture_iterations,) + level_shape + (nd,), dtype=np.float32)
                capture.buffer = ptr(capture_buffer)
        collect = int(self.logging_parameters.collect_per_level_convergence_reports)
        want_fields = bool(self.logging_parameters.collect_per_level_iteration_data)
        want_prints = bool(self.verbosity_parameters.print_per_iteration_info)
        self._iteration_data = []
        if want_fields or want_prints:
            # reference OptimizerWithTelemetry (optimizer_with_telemetry.tpp:83-182): the library runs one iteration at a
            # time and reports every iteration through a callback
            levels = self._optimize_with_telemetry(lib, params, ptr, canonical, live, shape, warp, kind, reports, collect,
                                                   stream, want_fields, want_prints)
            self._captured = None
        else:
            fn = lib.lsf_hier_optimize_2d if nd == 2 else lib.lsf_hier_optimize_3d
            levels = _lib.check(fn(ctypes.byref(params), ptr(canonical), ptr(live), *[ctypes.c_int(d) for d in shape],
                                   ptr(warp), kind, reports, collect, ctypes.byref(capture), stream))
            self._captured = None if capture_buffer is None else capture_buffer[:capture.count]
        self._reports = [_report_from_raw(nd, reports[i], bool(collect)) for i in range(levels)]
        return warp

    def _optimize_with_telemetry(self, lib, params, ptr, canonical, live, shape, warp, kind, reports, collect, stream,
                                 want_fields, want_prints):
        nd = self._nd
        verbosity = self.verbosity_parameters
        data_class = OptimizationIterationData2d if nd == 2 else OptimizationIterationData3d
        per_level = self._iteration_data
        state = {"level": -1}

        def level_done(level):
            if verbosity.print_per_level_info:
                print("[LEVEL %d COMPLETED]" % level)  # reference optimizer_with_telemetry.tpp:102-105

        def on_iteration(_user, record_pointer):
            record = record_pointer.contents
            if record.level != state["level"]:
                if state["level"] >= 0:
                    level_done(state["level"])
                state["level"] = record.level
                if want_fields:
                    per_level.append(data_class())
            if want_fields:
                dims = tuple(record.dims[:nd])
                count = int(np.prod(dims))

                def field(pointer, channels):
                    if not pointer:
                        return np.zeros((0,) * (nd + 1), np.float32)  # the reference stores an empty container
                    flat = np.ctypeslib.as_array(pointer, shape=(count * channels,))
                    return flat.reshape(dims + ((channels,) if channels > 1 else ())).copy()

                per_level[-1].add_iteration_result(field(record.live_field, 1), field(record.warp_field, nd),
                                                   field(record.data_term_gradient, nd),
                                                   field(record.tikhonov_term_gradient, nd))
            if want_prints and record.iteration >= 0:
                # reference optimizer_with_telemetry.tpp:161-181 (current_iteration is printed before its increment)
                line = "[ITERATION %d COMPLETED]" % record.iteration
                if verbosity.print_iteration_max_warp_update:
                    line += " [max upd. l.: %g]" % record.max_update_length
                if verbosity.print_iteration_mean_tsdf_difference:
                    line += " [mean diff.: %g]" % record.mean_tsdf_difference
                if verbosity.print_iteration_std_tsdf_difference:
                    line += " [std diff.: %g]" % record.std_tsdf_difference
                if verbosity.print_iteration_data_energy:
                    line += " [norm. data energy: %g]" % record.normalized_data_energy
                if verbosity.print_iteration_tikhonov_energy and self.tikhonov_term_enabled:
                    line += " [norm. tikhonov energy: %g]" % record.normalized_tikhonov_energy
                print(line)
            self._last_iteration_statistics = (record.max_update_length, record.mean_tsdf_difference,
                                               record.std_tsdf_difference, record.normalized_data_energy,
                                               record.normalized_tikhonov_energy)
            if record.iteration >= 0:
                self._iteration_statistics.append((record.level, record.iteration) + self._last_iteration_statistics)

        self._iteration_statistics = []
        sink = _lib.IterationSink()
        callback = _lib.ITERATION_CALLBACK(on_iteration)
        sink.callback = callback
        sink.user = None
        sink.want_fields = int(want_fields)
        sink.want_statistics = 1
        fn = lib.lsf_hier_optimize_2d_telemetry if nd == 2 else lib.lsf_hier_optimize_3d_telemetry
        levels = _lib.check(fn(ctypes.byref(params), ptr(canonical), ptr(live), *[ctypes.c_int(d) for d in shape],
                               ptr(warp), kind, reports, collect, ctypes.byref(sink), stream))
        if state["level"] >= 0:
            level_done(state["level"])
        return levels

    def optimize_batch(self, canonical_fields, live_fields):
        """(extension; 3D) optimize() for a batch of independent pairs in one call -- the reference's multi-pair loop
        (run_hierarchical_optimizer3d_multipair.py:403-406): `canonical_fields` / `live_fields` are [P, X, Y, Z] arrays or
        CUDA tensors, the result is the [P, X, Y, Z, 3] warp fields. The pairs move through the pyramid levels together
        (one kernel launch sequence per iteration for the whole batch, per-pair termination); every pair's result equals
        optimize() of that pair bit for bit. Iteration counts: get_per_pair_iteration_counts()."""
        if self._nd != 3:
            raise ValueError("optimize_batch exists for 3D pairs only")
        on_device = _lib.is_torch_cuda(canonical_fields) or _lib.is_torch_cuda(live_fields)
        if on_device:
            import torch
            if not (_lib.is_torch_cuda(canonical_fields) and _lib.is_torch_cuda(live_fields)):
                raise ValueError("canonical_fields and live_fields must live on the same device")
            if canonical_fields.device != live_fields.device:
                raise ValueError("canonical_fields and live_fields must live on the same device")
            canonical = canonical_fields.contiguous().float()
            live = live_fields.contiguous().float()
        else:
            canonical = _lib.as_f32(canonical_fields)
            live = _lib.as_f32(live_fields)
        shape = tuple(int(d) for d in canonical.shape)
        if len(shape) != 4 or tuple(live.shape) != shape:
            raise ValueError("expected two [P, X, Y, Z] batches of equal shape, got %s and %s"
                             % (tuple(canonical.shape), tuple(live.shape)))
        lib = _lib.load()
        params = self._params()
        pairs = shape[0]
        counts = (ctypes.c_int * (max(pairs, 1) * _lib.LSF_MAX_LEVELS))()
        if on_device:
            import torch
            with torch.cuda.device(canonical.device):
                warp = torch.empty(shape + (3,), dtype=torch.float32, device=canonical.device)
                ptr = lambda t: ctypes.cast(ctypes.c_void_p(t.data_ptr()), _lib.c_float_p)
                stream = ctypes.c_void_p(torch.cuda.current_stream(canonical.device).cuda_stream)
                levels = _lib.check(lib.lsf_hier_optimize_3d_batch(ctypes.byref(params), ptr(canonical), ptr(live), pairs,
                                                                   shape[1], shape[2], shape[3], ptr(warp),
                                                                   _lib.LSF_DEVICE, counts, stream))
        else:
            warp = _lib.result_array(shape + (3,))
            levels = _lib.check(lib.lsf_hier_optimize_3d_batch(ctypes.byref(params), _lib.fptr(canonical), _lib.fptr(live),
                                                               pairs, shape[1], shape[2], shape[3], _lib.fptr(warp),
                                                               _lib.LSF_HOST, counts, _lib.host_stream_handle()))
        self._pair_iteration_counts = [[counts[p * _lib.LSF_MAX_LEVELS + l] for l in range(levels)] for p in range(pairs)]
        return warp

    def get_per_pair_iteration_counts(self):
        """(extension) per-level iteration counts of every pair of the last optimize_batch() call"""
        return [list(c) for c in getattr(self, "_pair_iteration_counts", [])]

    def get_per_iteration_statistics(self):
        """(extension) the numbers behind the reference's per-iteration prints, one tuple per iteration of the last
        optimize() call that ran with telemetry: (level, iteration, max update length, mean diff, std diff, normalised
        data energy, normalised Tikhonov energy)."""
        return list(getattr(self, "_iteration_statistics", []))

    def get_per_level_convergence_reports(self):
        return list(self._reports)

    def get_per_level_iteration_counts(self):
        return [r.iteration_count for r in self._reports]

    def get_captured_warps(self):
        return self._captured

    def get_per_level_iteration_data(self):
        return list(self._iteration_data)


class HierarchicalOptimizer2d(_HierarchicalOptimizer):
    """reference `level_set_fusion_optimization.HierarchicalOptimizer2d`"""
    _nd = 2


class HierarchicalOptimizer3d(_HierarchicalOptimizer):
    """reference `level_set_fusion_optimization.HierarchicalOptimizer3d`"""
    _nd = 3
