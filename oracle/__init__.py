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
    src = os.path.join(_HERE, "lsf_oracle.cpp")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
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
