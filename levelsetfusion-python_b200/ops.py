"""Field-math primitives of the path, numpy in / numpy out (or torch CUDA tensors in / out, zero-copy).

Mirrors the functions the reference exposes or tests for this path:
  warp / warp_with_replacement   cpp/src/nonrigid_optimization/field_warping.tpp:212-225,
                                 nonrigid_opt/field_warping.py:67-109
  gradient / laplacian           cpp/src/math/gradients.tpp
  convolve_with_kernel[_preserve_zeros]   cpp/src/math/convolution.cpp, math_utils/convolution.py:70-132
  downsample / upsample          cpp/src/math/resampling.tpp
  max_norm                       cpp/src/math/statistics.tpp:57-100
"""
import ctypes

import numpy as np

from . import _lib


def _prepare(*arrays):
    """Returns (memory_kind, stream, converted arrays, maker of outputs)."""
    if any(_lib.is_torch_cuda(a) for a in arrays):
        if not all(a is None or _lib.is_torch_cuda(a) for a in arrays):
            raise ValueError("all fields of a call must be torch CUDA tensors, or all numpy arrays")
        _lib.check_device(*arrays)
        converted = [a.contiguous().float() if a is not None else None for a in arrays]
        return _lib.LSF_DEVICE, _lib.current_stream_handle(), converted, "torch"
    converted = [_lib.as_f32(a) if a is not None else None for a in arrays]
    return _lib.LSF_HOST, _lib.host_stream_handle(), converted, "numpy"


def _ptr(a):
    if isinstance(a, np.ndarray):
        return _lib.fptr(a)
    return ctypes.cast(ctypes.c_void_p(a.data_ptr()), _lib.c_float_p)


def _empty(shape, like, flavour):
    if flavour == "torch":
        import torch
        return torch.empty(shape, dtype=torch.float32, device=like.device)
    return np.empty(shape, dtype=np.float32)


def _dims(shape):
    return [ctypes.c_int(int(d)) for d in shape]


def warp(field, warp_field, oob_value=1.0):
    """Trilinear / bilinear resample of `field` at p + warp(p); out-of-bounds taps read `oob_value`
    (reference `warp`: 1.0)."""
    kind, stream, (field, warp_field), flavour = _prepare(field, warp_field)
    nd = warp_field.ndim - 1
    if warp_field.shape[-1] != nd or tuple(field.shape[:nd]) != tuple(warp_field.shape[:nd]):
        raise ValueError("field %s and warp field %s do not match" % (tuple(field.shape), tuple(warp_field.shape)))
    channels = 1 if field.ndim == nd else int(field.shape[-1])
    out = _empty(tuple(field.shape), field, flavour)
    fn = _lib.load().lsf_warp_2d if nd == 2 else _lib.load().lsf_warp_3d
    _lib.check(fn(_ptr(field), channels, _ptr(warp_field), *_dims(warp_field.shape[:nd]), ctypes.c_float(oob_value),
                  _ptr(out), kind, stream))
    return out


def warp_with_replacement(field, warp_field, replacement=0.0):
    return warp(field, warp_field, oob_value=replacement)


def gradient(field):
    kind, stream, (field,), flavour = _prepare(field)
    nd = field.ndim
    out = _empty(tuple(field.shape) + (nd,), field, flavour)
    fn = _lib.load().lsf_gradient_2d if nd == 2 else _lib.load().lsf_gradient_3d
    _lib.check(fn(_ptr(field), *_dims(field.shape), _ptr(out), kind, stream))
    return out


def laplacian(vector_field):
    kind, stream, (vector_field,), flavour = _prepare(vector_field)
    nd = vector_field.ndim - 1
    out = _empty(tuple(vector_field.shape), vector_field, flavour)
    fn = _lib.load().lsf_laplacian_2d if nd == 2 else _lib.load().lsf_laplacian_3d
    _lib.check(fn(_ptr(vector_field), *_dims(vector_field.shape[:nd]), _ptr(out), kind, stream))
    return out


def convolve_with_kernel(vector_field, kernel, preserve_zeros=False):
    """Returns the filtered field (the reference filters its argument in place)."""
    kind, stream, (vector_field,), flavour = _prepare(vector_field)
    out = vector_field.clone() if flavour == "torch" else vector_field.copy()
    kernel = _lib.as_f32(np.asarray(kernel))
    nd = out.ndim - 1
    if nd == 2:
        _lib.check(_lib.load().lsf_convolve_2d(_ptr(out), *_dims(out.shape[:2]), _lib.fptr(kernel), int(kernel.size),
                                               int(bool(preserve_zeros)), kind, stream))
    else:
        if preserve_zeros:
            raise ValueError("preserve_zeros is a 2D-only variant in the reference (convolution.cpp:69-145)")
        _lib.check(_lib.load().lsf_convolve_3d(_ptr(out), *_dims(out.shape[:3]), _lib.fptr(kernel), int(kernel.size),
                                               kind, stream))
    return out


def convolve_with_kernel_preserve_zeros(vector_field, kernel):
    return convolve_with_kernel(vector_field, kernel, preserve_zeros=True)


def _resample(field, nd, linear, up):
    kind, stream, (field,), flavour = _prepare(field)
    channels = 1 if field.ndim == nd else int(field.shape[-1])
    sdims = tuple(int(d) for d in field.shape[:nd])
    odims = tuple(d * 2 for d in sdims) if up else tuple(d // 2 for d in sdims)
    out = _empty(odims + tuple(field.shape[nd:]), field, flavour)
    lib = _lib.load()
    fn = {(2, True): lib.lsf_upsample_2d, (2, False): lib.lsf_downsample_2d,
          (3, True): lib.lsf_upsample_3d, (3, False): lib.lsf_downsample_3d}[(nd, up)]
    _lib.check(fn(_ptr(field), channels, *_dims(sdims), int(bool(linear)), _ptr(out), kind, stream))
    return out


def downsample(field, nd, linear=False):
    return _resample(field, nd, linear, False)


def upsample(field, nd, linear=False):
    return _resample(field, nd, linear, True)


def max_norm(vector_field):
    kind, stream, (vector_field,), _ = _prepare(vector_field)
    channels = int(vector_field.shape[-1])
    count = 1
    for d in vector_field.shape[:-1]:
        count *= int(d)
    out = ctypes.c_float(0.0)
    _lib.check(_lib.load().lsf_max_norm(_ptr(vector_field), channels, ctypes.c_longlong(count), ctypes.byref(out),
                                        kind, stream))
    return float(out.value)


def locate_max_norm(vector_field):
    """(max_norm, coordinates) of a [H][W][C] or [X][Y][Z][C] vector field (reference math::locate_max_norm,
    cpp/src/math/statistics.tpp:57-100): coordinates = (x, y) = (column, row) for a 2D field, (x, y, z) for a 3D field; among
    equal maxima the element the reference's column-major traversal meets first"""
    kind, stream, (vector_field,), _ = _prepare(vector_field)
    nd = len(vector_field.shape) - 1
    if nd not in (2, 3):
        raise ValueError("expected a [H][W][C] or [X][Y][Z][C] vector field, got shape %s" % (tuple(vector_field.shape),))
    dims = (ctypes.c_int * 3)(*[int(d) for d in vector_field.shape[:nd]])
    out = ctypes.c_float(0.0)
    coordinates = (ctypes.c_int * 3)()
    _lib.check(_lib.load().lsf_locate_max_norm(_ptr(vector_field), int(vector_field.shape[-1]), nd, dims, ctypes.byref(out),
                                               coordinates, kind, stream))
    return float(out.value), tuple(int(c) for c in coordinates[:nd])
