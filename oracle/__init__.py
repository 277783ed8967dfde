"""CPU oracle (TEST INFRASTRUCTURE ONLY).

ctypes front-end of ``liblsf_oracle.so`` (``oracle/lsf_oracle.cpp``), the from-scratch CPU restatement of
the reference's warp-field optimisation path. Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import this package; the product package
(``levelsetfusion-python_b200`` a.k.a. ``lsf_b200``) never does.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liblsf_oracle.so")
_lib = None

c_float_p = ctypes.POINTER(ctypes.c_float)
c_int_p = ctypes.POINTER(ctypes.c_int)


class HierParams(ctypes.Structure):
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


class IterationDump(ctypes.Structure):
    _fields_ = [
        ("level", ctypes.c_int),
        ("max_iterations", ctypes.c_int),
        ("buffer", c_float_p),
        ("count", ctypes.c_int),
    ]


def build(force=False):
    """Compile the oracle with the recipe in oracle/Makefile (g++ -O3 -fopenmp -ffp-contract=off)."""
    sources = [os.path.join(_HERE, name) for name in ("lsf_oracle.cpp", "lsf_oracle_slavcheva.cpp", "lsf_oracle_tsdf.cpp", "lsf_oracle_rigid.cpp",
                                                          "lsf_oracle.h")]
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < max(map(os.path.getmtime, sources)):
        subprocess.run(["make", "-C", _HERE, "-B", "liblsf_oracle.so"], check=True, capture_output=True)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.orc_max_norm2d.restype = ctypes.c_float
        _lib.orc_hier_time_iterations3d.restype = ctypes.c_double
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a):
    return a.ctypes.data_as(c_float_p)


def num_threads():
    return int(lib().orc_num_threads())


def use_all_cores():
    """torch.distributed.run exports OMP_NUM_THREADS=1; the CPU-baseline timings use every host core instead"""
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    lib().orc_set_num_threads(int(cores))
    return num_threads()


# ------------------------------------------------------------------ primitives
def warp(field, warp_field):
    """reference `warp` (OOB -> 1.0): field_warping.tpp:68-140,145-194"""
    field, warp_field = _f32(field), _f32(warp_field)
    out = np.empty_like(field)
    if field.ndim == 2:
        lib().orc_warp2d(_p(field), _p(warp_field), *map(ctypes.c_int, field.shape), _p(out))
    else:
        lib().orc_warp3d(_p(field), _p(warp_field), *map(ctypes.c_int, field.shape), _p(out))
    return out


def warp_with_replacement(field, warp_field, replacement=0.0):
    """reference `warp_with_replacement`; `field` may be scalar [..] or vector [..,D]"""
    field, warp_field = _f32(field), _f32(warp_field)
    nd = warp_field.ndim - 1
    C = 1 if field.ndim == nd else field.shape[-1]
    out = np.empty_like(field)
    dims = list(map(ctypes.c_int, warp_field.shape[:nd]))
    if nd == 2:
        lib().orc_warp2d_replacement(_p(field), C, _p(warp_field), *dims, ctypes.c_float(replacement), _p(out))
    else:
        lib().orc_warp3d_replacement(_p(field), C, _p(warp_field), *dims, ctypes.c_float(replacement), _p(out))
    return out


def gradient(field):
    field = _f32(field)
    out = np.empty(field.shape + (field.ndim,), dtype=np.float32)
    if field.ndim == 2:
        lib().orc_gradient2d(_p(field), *map(ctypes.c_int, field.shape), _p(out))
    else:
        lib().orc_gradient3d(_p(field), *map(ctypes.c_int, field.shape), _p(out))
    return out


def laplacian(vfield):
    vfield = _f32(vfield)
    out = np.empty_like(vfield)
    C = vfield.shape[-1]
    dims = list(map(ctypes.c_int, vfield.shape[:-1]))
    if vfield.ndim == 3:
        lib().orc_laplacian2d(_p(vfield), C, *dims, _p(out))
    else:
        lib().orc_laplacian3d(_p(vfield), C, *dims, _p(out))
    return out


def convolve_with_kernel(vfield, kernel, preserve_zeros=False):
    """returns a filtered copy (the reference filters in place)"""
    out = _f32(vfield).copy()
    kernel = _f32(kernel)
    C = out.shape[-1]
    dims = list(map(ctypes.c_int, out.shape[:-1]))
    if out.ndim == 3:
        lib().orc_convolve2d(_p(out), C, *dims, _p(kernel), int(kernel.size), int(preserve_zeros))
    else:
        assert not preserve_zeros
        lib().orc_convolve3d(_p(out), C, *dims, _p(kernel), int(kernel.size))
    return out


def _resample(field, nd, linear, up):
    field = _f32(field)
    C = 1 if field.ndim == nd else field.shape[-1]
    sdims = field.shape[:nd]
    odims = tuple(d * 2 for d in sdims) if up else tuple(d // 2 for d in sdims)
    out = np.empty(odims + field.shape[nd:], dtype=np.float32)
    fn = {(2, True): lib().orc_upsample2d, (2, False): lib().orc_downsample2d,
          (3, True): lib().orc_upsample3d, (3, False): lib().orc_downsample3d}[(nd, up)]
    status = fn(_p(field), C, *map(ctypes.c_int, sdims), int(linear), _p(out))
    if status != 0:
        raise RuntimeError("oracle resampling precondition failed (status %d)" % status)
    return out


def downsample(field, nd, linear=False):
    return _resample(field, nd, linear, False)


def upsample(field, nd, linear=False):
    return _resample(field, nd, linear, True)


def max_norm(vfield):
    vfield = _f32(vfield)
    C = vfield.shape[-1]
    return float(lib().orc_max_norm2d(_p(vfield), C, ctypes.c_long(vfield.size // C)))


def locate_max_norm(vfield):
    """reference math::locate_max_norm (cpp/src/math/statistics.tpp:57-100) on a [H][W][C] / [X][Y][Z][C] array:
    (max_norm, coordinates). Squared lengths summed component by component in float32; the traversal is column-major over the
    elements with a strict `>` (the first maximum stays); 2D coordinates decoded like statistics.tpp:70-71
    (x = i / column_count, y = i % column_count), 3D ones by unravel_3d_index (x fastest)."""
    vfield = _f32(vfield)
    nd = vfield.ndim - 1
    sq = np.zeros(vfield.shape[:-1], np.float32)
    for c in range(vfield.shape[-1]):
        sq = sq + vfield[..., c] * vfield[..., c]
    order = sq.ravel(order="F")  # the reference's element order
    best = int(np.argmax(order)) if order.max() > 0 else 0  # np.argmax: first occurrence
    value = float(np.sqrt(np.float32(order[best]) if order.max() > 0 else np.float32(0)))
    if nd == 2:
        columns = vfield.shape[1]
        return value, (best // columns, best % columns)
    X, Y = vfield.shape[0], vfield.shape[1]
    return value, (best % X, (best // X) % Y, best // (X * Y))


# ------------------------------------------------------------------ hierarchical optimizer
def make_hier_params(tikhonov_term_enabled=True, gradient_kernel_enabled=True, maximum_chunk_size=8, rate=0.1,
                     maximum_iteration_count=100, maximum_warp_update_threshold=0.001, data_term_amplifier=1.0,
                     tikhonov_strength=0.2, kernel=None, resampling_strategy=0):
    p = HierParams()
    p.tikhonov_term_enabled = int(tikhonov_term_enabled)
    p.gradient_kernel_enabled = int(gradient_kernel_enabled)
    p.maximum_chunk_size = int(maximum_chunk_size)
    p.rate = rate
    p.maximum_iteration_count = int(maximum_iteration_count)
    p.maximum_warp_update_threshold = maximum_warp_update_threshold
    p.data_term_amplifier = data_term_amplifier
    p.tikhonov_strength = tikhonov_strength
    keep = None
    if kernel is not None and len(kernel) > 0:
        keep = _f32(kernel)
        p.kernel = _p(keep)
        p.kernel_size = int(keep.size)
    else:
        p.kernel = None
        p.kernel_size = 0
    p.resampling_strategy = int(resampling_strategy)
    p._keep = keep  # keep the kernel array alive
    return p


def hier_optimize(canonical, live, dump_level=-1, dump_iterations=0, **kwargs):
    """reference Optimizer::optimize (optimizer.tpp:83-131). Returns dict(warp, iterations, max_updates, dump)."""
    canonical, live = _f32(canonical), _f32(live)
    assert canonical.shape == live.shape
    nd = canonical.ndim
    p = make_hier_params(**kwargs)
    warp_out = np.zeros(canonical.shape + (nd,), dtype=np.float32)
    iters = np.zeros(32, dtype=np.int32)
    maxes = np.zeros(32, dtype=np.float32)
    dump = IterationDump()
    dump.level = dump_level
    dump.max_iterations = dump_iterations
    dump_buf = None
    if dump_level >= 0 and dump_iterations > 0:
        level_count = int(np.log2(p.maximum_chunk_size)) + 1
        shrink = 2 ** (level_count - 1 - dump_level)
        ldims = tuple(d // shrink for d in canonical.shape)
        dump_buf = np.zeros((dump_iterations,) + ldims + (nd,), dtype=np.float32)
        dump.buffer = _p(dump_buf)
    fn = lib().orc_hier_optimize2d if nd == 2 else lib().orc_hier_optimize3d
    levels = fn(ctypes.byref(p), _p(canonical), _p(live), *map(ctypes.c_int, canonical.shape), _p(warp_out),
                iters.ctypes.data_as(c_int_p), _p(maxes), ctypes.byref(dump))
    if levels < 0:
        raise RuntimeError("oracle hierarchical optimizer precondition failed (status %d)" % levels)
    return dict(warp=warp_out, iterations=iters[:levels].tolist(), max_updates=maxes[:levels].copy(),
                dump=None if dump_buf is None else dump_buf[:dump.count])


def hier_time_iterations3d(canonical, live, iterations, **kwargs):
    canonical, live = _f32(canonical), _f32(live)
    p = make_hier_params(**kwargs)
    return float(lib().orc_hier_time_iterations3d(ctypes.byref(p), _p(canonical), _p(live),
                                                  *map(ctypes.c_int, canonical.shape), int(iterations)))


# ------------------------------------------------------------------ slavcheva (SobolevFusion / KillingFusion) optimizers
SEMANTICS_CPP, SEMANTICS_PY_DIRECT, SEMANTICS_PY_VECTORIZED = 0, 1, 2
DATA_TERM_BASIC, DATA_TERM_THRESHOLDED_FDM = 0, 1
SMOOTHING_TIKHONOV, SMOOTHING_KILLING = 0, 1


class SlavchevaParams(ctypes.Structure):
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
        ("max_iterations", ctypes.c_int),
        ("min_iterations", ctypes.c_int),
        ("kernel", c_float_p),
        ("kernel_size", ctypes.c_int),
    ]


class WarpDeltaStatistics(ctypes.Structure):
    _fields_ = [("ratio_above_min_threshold", ctypes.c_float), ("length_min", ctypes.c_float),
                ("length_max", ctypes.c_float), ("length_mean", ctypes.c_float),
                ("length_standard_deviation", ctypes.c_float), ("longest_warp_location", ctypes.c_int * 3),
                ("is_largest_below_min_threshold", ctypes.c_int), ("is_largest_above_max_threshold", ctypes.c_int)]


class TsdfDifferenceStatistics(ctypes.Structure):
    _fields_ = [("difference_min", ctypes.c_float), ("difference_max", ctypes.c_float),
                ("difference_mean", ctypes.c_float), ("difference_standard_deviation", ctypes.c_float),
                ("biggest_difference_location", ctypes.c_int * 3)]


def make_slavcheva_params(semantics=SEMANTICS_CPP, data_term_method=DATA_TERM_BASIC,
                          smoothing_term_method=SMOOTHING_TIKHONOV, level_set_term_enabled=False,
                          sobolev_smoothing_enabled=True, gradient_descent_rate=0.1, data_term_weight=1.0,
                          smoothing_term_weight=0.2, isomorphic_enforcement_factor=0.1, level_set_term_weight=0.2,
                          maximum_warp_length_lower_threshold=0.1, maximum_warp_length_upper_threshold=10000.0,
                          max_iterations=100, min_iterations=1, sobolev_kernel=None):
    p = SlavchevaParams()
    p.semantics = int(semantics)
    p.data_term_method = int(data_term_method)
    p.smoothing_term_method = int(smoothing_term_method)
    p.level_set_term_enabled = int(bool(level_set_term_enabled))
    p.sobolev_smoothing_enabled = int(bool(sobolev_smoothing_enabled))
    p.gradient_descent_rate = gradient_descent_rate
    p.data_term_weight = data_term_weight
    p.smoothing_term_weight = smoothing_term_weight
    p.isomorphic_enforcement_factor = isomorphic_enforcement_factor
    p.level_set_term_weight = level_set_term_weight
    p.maximum_warp_length_lower_threshold = maximum_warp_length_lower_threshold
    p.maximum_warp_length_upper_threshold = maximum_warp_length_upper_threshold
    p.max_iterations = int(max_iterations)
    p.min_iterations = int(min_iterations)
    keep = None
    if sobolev_kernel is not None and len(sobolev_kernel) > 0:
        keep = _f32(sobolev_kernel)
        p.kernel = _p(keep)
        p.kernel_size = int(keep.size)
    else:
        p.kernel = None
        p.kernel_size = 0
    p._keep = keep
    return p


def _dims(shape):
    return (ctypes.c_int * len(shape))(*[int(d) for d in shape])


def slavcheva_optimize(live, canonical, dump_iterations=0, **kwargs):
    """reference SobolevOptimizer2d.optimize(live, canonical) / SlavchevaOptimizer2d.optimize (+ 3D generalisation).
    Returns dict(live, warp, iterations, max_warps, dump, energies); energies [iterations][3] = the {data, smoothing,
    level set} energy aggregates the reference's Python optimizer appends to its log every iteration (2D, Python semantics;
    zeros otherwise)."""
    live, canonical = _f32(live), _f32(canonical)
    assert live.shape == canonical.shape
    nd = live.ndim
    p = make_slavcheva_params(**kwargs)
    live_out = np.empty_like(live)
    warp_out = np.empty(live.shape + (nd,), dtype=np.float32)
    iterations = ctypes.c_int(0)
    capacity = max(int(p.max_iterations), int(p.min_iterations), 1)
    max_warps = np.zeros(capacity, dtype=np.float32)
    dump = IterationDump()
    dump.level = 0
    dump.max_iterations = int(dump_iterations)
    dump_buffer = None
    if dump_iterations > 0:
        dump_buffer = np.zeros((dump_iterations,) + live.shape + (nd,), dtype=np.float32)
        dump.buffer = _p(dump_buffer)
    energies = np.zeros((capacity, 3), dtype=np.float64)
    status = lib().orc_slavcheva_optimize_energies(ctypes.byref(p), _p(live), _p(canonical), nd, _dims(live.shape),
                                                   _p(live_out), _p(warp_out), ctypes.byref(iterations), _p(max_warps),
                                                   capacity, ctypes.byref(dump),
                                                   energies.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), capacity)
    if status != 0:
        raise RuntimeError("oracle slavcheva optimizer precondition failed (status %d)" % status)
    return dict(live=live_out, warp=warp_out, iterations=int(iterations.value),
                max_warps=max_warps[:iterations.value].copy(),
                dump=None if dump_buffer is None else dump_buffer[:dump.count],
                energies=energies[:iterations.value].copy())


def slavcheva_energies(live, canonical, warp, **kwargs):
    """{data, smoothing, level set} energy aggregates the reference's Python optimizer logs for an iteration that starts
    from (live, warp) (slavcheva_optimizer2d.py:163-175,236-300); float64 array of 3."""
    live, canonical, warp = _f32(live), _f32(canonical), _f32(warp)
    p = make_slavcheva_params(**kwargs)
    out = np.zeros(3, dtype=np.float64)
    lib().orc_slavcheva_energies(ctypes.byref(p), _p(live), _p(canonical), _p(warp), live.ndim, _dims(live.shape),
                                 out.ctypes.data_as(ctypes.POINTER(ctypes.c_double)))
    return out


def slavcheva_data_term(live, canonical, band_union_only=False, **kwargs):
    live, canonical = _f32(live), _f32(canonical)
    p = make_slavcheva_params(**kwargs)
    out = np.empty(live.shape + (live.ndim,), dtype=np.float32)
    lib().orc_slavcheva_data_term(ctypes.byref(p), _p(live), _p(canonical), live.ndim, _dims(live.shape),
                                  int(band_union_only), _p(out))
    return out


def slavcheva_smoothing_term(warp_field, live=None, canonical=None, band_union_only=False, **kwargs):
    warp_field = _f32(warp_field)
    nd = warp_field.ndim - 1
    shape = warp_field.shape[:nd]
    live = np.zeros(shape, np.float32) if live is None else _f32(live)
    canonical = np.zeros(shape, np.float32) if canonical is None else _f32(canonical)
    p = make_slavcheva_params(**kwargs)
    out = np.empty_like(warp_field)
    lib().orc_slavcheva_smoothing_term(ctypes.byref(p), _p(warp_field), _p(live), _p(canonical), nd, _dims(shape),
                                       int(band_union_only), _p(out))
    return out


def slavcheva_level_set_term(live, **kwargs):
    live = _f32(live)
    p = make_slavcheva_params(**kwargs)
    out = np.empty(live.shape + (live.ndim,), dtype=np.float32)
    lib().orc_slavcheva_level_set_term(ctypes.byref(p), _p(live), live.ndim, _dims(live.shape), _p(out))
    return out


def warp_advanced(live, canonical, warp_field, band_union_only=False, known_values_only=False,
                  substitute_original=False, truncation_float_threshold=1e-6, modify_warp=True):
    """reference warp_2d_advanced (field_warping.cpp:64-154) / its 3D generalisation -> (new live, warp)"""
    live, canonical = _f32(live), _f32(canonical)
    warp_field = _f32(warp_field).copy()
    new_live = np.empty_like(live)
    lib().orc_warp_advanced(_p(live), _p(canonical), _p(warp_field), live.ndim, _dims(live.shape), int(band_union_only),
                            int(known_values_only), int(substitute_original),
                            ctypes.c_float(truncation_float_threshold), int(modify_warp), _p(new_live))
    return new_live, warp_field


def warp_delta_statistics(warp_field, canonical, live, min_threshold, max_threshold):
    warp_field, canonical, live = _f32(warp_field), _f32(canonical), _f32(live)
    out = WarpDeltaStatistics()
    lib().orc_warp_delta_statistics(_p(warp_field), _p(canonical), _p(live), live.ndim, _dims(live.shape),
                                    ctypes.c_float(min_threshold), ctypes.c_float(max_threshold), ctypes.byref(out))
    return out


def tsdf_difference_statistics(canonical, live):
    canonical, live = _f32(canonical), _f32(live)
    out = TsdfDifferenceStatistics()
    lib().orc_tsdf_difference_statistics(_p(canonical), _p(live), live.ndim, _dims(live.shape), ctypes.byref(out))
    return out


# ------------------------------------------------------------------------------------------------ TSDF generation
class TsdfParams(ctypes.Structure):
    """orc_tsdf_params (oracle/lsf_oracle.h)"""
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


def _tsdf_params(nd, projection_matrix, array_offset, field_shape, depth_unit_ratio, near_clipping_distance, voxel_size,
                 narrow_band_width_voxels, filtering_method, smoothing_factor):
    p = TsdfParams()
    p.depth_unit_ratio = depth_unit_ratio
    p.projection_matrix = (ctypes.c_float * 9)(*np.asarray(projection_matrix, dtype=np.float32).reshape(9))
    p.near_clipping_distance = near_clipping_distance
    offset = list(array_offset) + [0] * (3 - len(array_offset))
    shape = list(field_shape) + [1] * (3 - len(field_shape))
    p.array_offset = (ctypes.c_int * 3)(*[int(v) for v in offset])
    p.field_shape = (ctypes.c_int * 3)(*[int(v) for v in shape])
    p.voxel_size = voxel_size
    p.narrow_band_width_voxels = int(narrow_band_width_voxels)
    p.filtering_method = int(filtering_method)
    p.smoothing_factor = smoothing_factor
    return p, shape


def sdf2sdf_optimize(canonical_field, live_depth_image, image_y_coordinate, projection_matrix, array_offset, field_shape,
                     rate=0.5, maximum_iteration_count=60, eta=0.01, depth_unit_ratio=0.001, near_clipping_distance=0.05,
                     voxel_size=0.004, narrow_band_width_voxels=20, filtering_method=0, smoothing_factor=1.0, double_sums=False):
    """reference Sdf2SdfOptimizer2d(rate, maximum_iteration_count, tsdf_generation_parameters).optimize(image_y_coordinate,
    canonical_field, live_depth_image, eta, initial_camera_pose) (sdf_2_sdf_optimizer2d.cpp:63-124). Returns a dict with the
    3 x 3 twist matrix, the twist after every iteration and the energy of every iteration. `double_sums`: float32 terms
    added up in double instead of the reference's sequential float32 sums (the arithmetic of the GPU reduction)."""
    p, _ = _tsdf_params(2, projection_matrix, array_offset, field_shape, depth_unit_ratio, near_clipping_distance, voxel_size,
                        narrow_band_width_voxels, filtering_method, smoothing_factor)
    canonical = np.ascontiguousarray(canonical_field, dtype=np.float32)
    depth = np.ascontiguousarray(live_depth_image, dtype=np.uint16)
    matrix = np.zeros((3, 3), dtype=np.float32)
    twists = np.zeros((maximum_iteration_count, 3), dtype=np.float32)
    energies = np.zeros(maximum_iteration_count, dtype=np.float32)
    status = lib().orc_sdf2sdf_optimize(ctypes.byref(p), ctypes.c_float(rate), int(maximum_iteration_count),
                                        int(image_y_coordinate), _p(canonical),
                                        depth.ctypes.data_as(ctypes.POINTER(ctypes.c_ushort)), int(depth.shape[0]),
                                        int(depth.shape[1]), ctypes.c_float(eta), int(bool(double_sums)), _p(matrix),
                                        _p(twists), _p(energies))
    if status != 0:
        raise RuntimeError("oracle rigid tracker: unsupported parameters (status %d)" % status)
    return {"twist_matrix": matrix, "twists": twists, "energies": energies}


def tsdf_generate(depth_image, camera_pose, nd, projection_matrix, array_offset, field_shape, image_y_coordinate=0,
                  depth_unit_ratio=0.001, near_clipping_distance=0.05, voxel_size=0.004, narrow_band_width_voxels=20,
                  filtering_method=0, smoothing_factor=1.0):
    """reference tsdf::Generator{2d,3d}::generate (generator_crtp.tpp:40-71) with FilteringMethod NONE (0),
    EWA_IMAGE_SPACE (3), EWA_VOXEL_SPACE (4) or EWA_VOXEL_SPACE_INCLUSIVE (5) (generator_tensor.tpp:40-270,
    generator_matrix.tpp:33-238). Returns [x][y][z] (3D) or [y][x] (2D)."""
    depth = np.ascontiguousarray(depth_image, dtype=np.uint16)
    pose = np.ascontiguousarray(camera_pose, dtype=np.float32).reshape(16)
    p = TsdfParams()
    p.depth_unit_ratio = depth_unit_ratio
    p.projection_matrix = (ctypes.c_float * 9)(*np.asarray(projection_matrix, dtype=np.float32).reshape(9))
    p.near_clipping_distance = near_clipping_distance
    offset = list(array_offset) + [0] * (3 - len(array_offset))
    shape = list(field_shape) + [1] * (3 - len(field_shape))
    p.array_offset = (ctypes.c_int * 3)(*[int(v) for v in offset])
    p.field_shape = (ctypes.c_int * 3)(*[int(v) for v in shape])
    p.voxel_size = voxel_size
    p.narrow_band_width_voxels = int(narrow_band_width_voxels)
    p.filtering_method = int(filtering_method)
    p.smoothing_factor = smoothing_factor
    out = np.empty((shape[0], shape[1], shape[2]) if nd == 3 else (shape[1], shape[0]), dtype=np.float32)
    status = lib().orc_tsdf_generate(ctypes.byref(p), depth.ctypes.data_as(ctypes.POINTER(ctypes.c_ushort)),
                                     int(depth.shape[0]), int(depth.shape[1]), _p(pose), int(image_y_coordinate), nd, _p(out))
    if status != 0:
        raise RuntimeError("oracle TSDF generator: unsupported parameters (status %d)" % status)
    return out
