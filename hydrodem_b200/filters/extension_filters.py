"""Wrapped numpy / scipy filters (reference: ``filters/extension_filters.py``),
re-implemented as CUDA kernels behind the same class names.

The reference delegates these to ``numpy``, ``scipy.ndimage`` and
``scipy.fftpack``; here they are launches of libhydrodem_b200 (see
include/hydrodem_b200.h for the contract of each entry point).
"""
import copy
import ctypes

import numpy as np

from . import DeviceFilter
from .. import _lib, device as dev
from ..exceptions import DeviceError
from .simple_filters import _operand


class BitwiseXOR(DeviceFilter):
    """``np.bitwise_xor(operand, image)`` (extension_filters.py:12-60)."""

    def __init__(self, *, operand):
        self.operand = operand if isinstance(operand, dev.DeviceRaster) else copy.deepcopy(operand)

    def run_device(self, raster):
        b, scalar, proto = _operand(self.operand, raster)
        ref = np.bitwise_xor(proto, np.zeros(0, dtype=raster.ref_dtype)).dtype
        a = dev.convert(raster, dev.hd_dtype_of(raster.ref_dtype))
        if b is not None:
            b = dev.convert(b, dev.hd_dtype_of(b.ref_dtype))
        out = dev.empty(raster.ny, raster.nx, dev.hd_dtype_of(ref), ref)
        return dev.elementwise(_lib.OP_XOR, a, b, scalar, out)


class AbsoluteValues(DeviceFilter):
    """``np.abs(image)`` (extension_filters.py:63-95); complex -> real."""

    def run_device(self, raster):
        ref = np.abs(np.zeros(0, dtype=raster.ref_dtype)).dtype
        src = raster
        if np.dtype(raster.ref_dtype).kind == "c" and raster.dtype not in (_lib.C64, _lib.C128):
            raise DeviceError("complex raster stored as a real dtype")
        out = dev.empty(raster.ny, raster.nx, dev.hd_dtype_of(ref), ref)
        return dev.elementwise(_lib.OP_ABS, src, None, 0.0, out)


class Around(DeviceFilter):
    """``np.around(image)`` -- round half to even (extension_filters.py:98-130)."""

    def run_device(self, raster):
        ref = np.around(np.zeros(0, dtype=raster.ref_dtype)).dtype
        out = dev.empty(raster.ny, raster.nx, dev.hd_dtype_of(ref), ref)
        return dev.elementwise(_lib.OP_RINT, raster, None, 0.0, out)


class Convolve(DeviceFilter):
    """``scipy.ndimage.convolve(image, weights) / weights.size`` with the default
    mode='reflect' (extension_filters.py:133-184).  3x3 weights only (the
    pipeline uses ones((3, 3)))."""

    def __init__(self, weights=np.ones((3, 3))):
        self.weights = weights

    def run_device(self, raster, do_round=False, copy32=None):
        """``copy32``: optional F32 device raster that receives a float32 copy of the result (chain plumbing)."""
        w = np.asarray(self.weights, dtype=np.float64)
        if w.shape != (3, 3):
            raise DeviceError(f"Convolve: only 3x3 weights are implemented on the device (got {w.shape})")
        ref = np.dtype(raster.ref_dtype)
        work = np.float32 if ref == np.float32 else np.float64       # ndimage keeps float dtypes, else float64
        src = dev.convert(raster, dev.hd_dtype_of(work))
        out = dev.empty(raster.ny, raster.nx, src.dtype, work)
        corr = np.ascontiguousarray(w[::-1, ::-1])                    # convolution = correlation with the flipped kernel
        cw = (ctypes.c_double * 9)(*corr.ravel())
        _lib.check(_lib.load().hd_convolve3(src.ptr, src.pitch, out.ptr, out.pitch, src.dtype, src.ny, src.nx, cw,
                                            float(w.size), int(do_round), copy32.ptr if copy32 is not None else None,
                                            copy32.pitch if copy32 is not None else 0, dev.stream_ptr()))
        return out


class BinaryErosion(DeviceFilter):
    """``scipy.ndimage.binary_erosion(image, iterations=n)`` (extension_filters.py:187-235):
    cross structuring element, border_value 0, bool result."""

    def __init__(self, *, iterations):
        self.iterations = iterations

    def run_device(self, raster):
        return _morph(raster, _lib.MORPH_ERODE, None, self.iterations)


class BinaryClosing(DeviceFilter):
    """``scipy.ndimage.binary_closing(image, structure=...)`` (extension_filters.py:238-293)."""

    def __init__(self, *, structure=None):
        self.structure = structure

    def run_device(self, raster):
        return _morph(raster, _lib.MORPH_CLOSE, self.structure, 1)


def _morph(raster, op, structure, iterations):
    if structure is None:
        full = 0
    else:
        s = np.asarray(structure) != 0
        cross = np.array([[0, 1, 0], [1, 1, 1], [0, 1, 0]], dtype=bool)
        if s.shape == (3, 3) and s.all():
            full = 1
        elif s.shape == (3, 3) and (s == cross).all():
            full = 0
        else:
            raise DeviceError("binary morphology: only the 3x3 cross and the 3x3 square are implemented on the device")
    src = raster if raster.dtype in (_lib.U8, _lib.F32) else _nonzero_u8(raster)
    out = dev.empty(raster.ny, raster.nx, _lib.U8, np.bool_)
    _lib.check(_lib.load().hd_binary_morph(src.ptr, src.dtype, src.pitch, out.ptr, out.pitch, src.ny, src.nx, op, full,
                                           int(iterations), dev.stream_ptr()))
    return out


def _nonzero_u8(raster):
    """(raster != 0) as uint8 for dtypes the morphology kernel does not stage directly."""
    lt = dev.empty(raster.ny, raster.nx, _lib.U8, np.bool_)
    gt = dev.empty(raster.ny, raster.nx, _lib.U8, np.bool_)
    dev.elementwise(_lib.OP_LT, raster, None, 0.0, lt)
    dev.elementwise(_lib.OP_GT, raster, None, 0.0, gt)
    return dev.elementwise(_lib.OP_ADD, lt, gt, 0.0, dev.empty(raster.ny, raster.nx, _lib.U8, np.bool_))


class GreyDilation(DeviceFilter):
    """``scipy.ndimage.grey_dilation(image, size=(s, s))`` (extension_filters.py:296-345):
    flat square maximum filter, mode='reflect'."""

    def __init__(self, *, size):
        self.size = size

    def run_device(self, raster):
        size = self.size if np.ndim(self.size) else (self.size, self.size)
        if len(size) != 2 or size[0] != size[1] or size[0] % 2 != 1:
            raise DeviceError(f"GreyDilation: only odd square sizes are implemented on the device (got {self.size})")
        ref = np.dtype(raster.ref_dtype)
        if ref == np.float64 and raster.dtype == _lib.F32:
            src = raster                                              # narrow-exact storage (e.g. majority values)
        else:
            src = dev.convert(raster, _lib.F32 if ref == np.float32 else _lib.F64)
        out = dev.empty(raster.ny, raster.nx, src.dtype, ref)
        _lib.check(_lib.load().hd_max_filter(src.ptr, src.pitch, out.ptr, out.pitch, src.dtype, src.ny, src.nx,
                                             int(size[0]), dev.stream_ptr()))
        return out


def _complex_ref(ref_dtype):
    """Result dtype of scipy.fftpack.fft2 / ifft2: single precision stays single, everything else complex128."""
    return np.dtype(np.complex64) if np.dtype(ref_dtype) in (np.float32, np.complex64) else np.dtype(np.complex128)


def _fft2(raster, inverse):
    ref = np.dtype(raster.ref_dtype)
    if ref.kind == "c":
        src = dev.convert(raster, _lib.C64)
    else:
        src = dev.convert(raster, _lib.F32)
    out = dev.empty(raster.ny, raster.nx, _lib.C64, _complex_ref(ref))
    lib = _lib.load()
    plan = dev.fft_plan(raster.ny, raster.nx)
    nbytes = lib.hd_fft2_workspace_bytes(raster.ny, raster.nx)
    work = dev.scratch(nbytes)
    _lib.check(lib.hd_fft2_c2c(plan, src.ptr, src.dtype, src.pitch, out.ptr, out.pitch, int(inverse),
                               ctypes.c_void_p(work.data_ptr()), nbytes, dev.stream_ptr()))
    return out


class FourierTransform(DeviceFilter):
    """``scipy.fftpack.fft2(image)`` (extension_filters.py:348-379).  Computed in complex64 on the
    device (float32 input gives complex64 exactly like scipy; wider inputs are returned as
    complex128 but carry single-precision accuracy -- tolerance class)."""

    def run_device(self, raster):
        return _fft2(raster, inverse=False)


class FourierITransform(DeviceFilter):
    """``scipy.fftpack.ifft2(image)`` (extension_filters.py:382-414); see FourierTransform."""

    def run_device(self, raster):
        return _fft2(raster, inverse=True)


def _shift(raster, inverse):
    out = dev.empty(raster.ny, raster.nx, raster.dtype, raster.ref_dtype)
    src = raster
    if raster.itemsize not in (4, 8, 16):
        src = dev.convert(raster, _lib.F32)
        out = dev.empty(raster.ny, raster.nx, _lib.F32, raster.ref_dtype)
    _lib.check(_lib.load().hd_fftshift2(src.ptr, src.pitch, out.ptr, out.pitch, src.dtype, src.ny, src.nx, int(inverse),
                                        dev.stream_ptr()))
    return out


class FourierShift(DeviceFilter):
    """``scipy.fftpack.fftshift(image)`` (extension_filters.py:417-447)."""

    def run_device(self, raster):
        return _shift(raster, inverse=False)


class FourierIShift(DeviceFilter):
    """``scipy.fftpack.ifftshift(image)`` (extension_filters.py:450-480)."""

    def run_device(self, raster):
        return _shift(raster, inverse=True)
