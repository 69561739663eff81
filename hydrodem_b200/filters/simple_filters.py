"""Elementwise filters (reference: ``filters/simple_filters.py``), device backed.

Same class names and constructor signatures as the reference
(``LowerThan(*, value)``, ``ProductFilter(factor=1)`` ...).  Each is one
``hd_elementwise`` launch; operands may be scalars, ndarrays or device rasters.
The result dtype follows NumPy's promotion rules, evaluated on empty host
arrays so that the reference's dtype flow is reproduced.
"""
import numpy as np

from . import DeviceFilter
from .. import _lib, device as dev


def _operand(value, like):
    """-> (DeviceRaster | None, scalar, numpy dtype-or-python-scalar for promotion)."""
    if isinstance(value, dev.DeviceRaster):
        return value, 0.0, np.zeros(0, dtype=value.ref_dtype)
    if isinstance(value, np.ndarray):
        if value.ndim == 0:
            return None, value.item(), value
        if value.shape != like.shape:
            raise ValueError(f"operand shape {value.shape} does not match raster shape {like.shape}")
        return dev.upload(value), 0.0, np.zeros(0, dtype=value.dtype)
    return None, value, value


def binary_op(op, raster, operand, np_func):
    """out = np_func(operand, raster) on the device, in NumPy's result dtype."""
    b, scalar, proto = _operand(operand, raster)
    ref = np_func(proto, np.zeros(0, dtype=raster.ref_dtype)).dtype
    out = dev.empty(raster.ny, raster.nx, dev.hd_dtype_of(ref), ref)
    return dev.elementwise(op, raster, b, scalar, out)


class LowerThan(DeviceFilter):
    """``image < value`` -> bool (simple_filters.py:7-50)."""

    def __init__(self, *, value):
        self.value = value

    def run_device(self, raster):
        out = dev.empty(raster.ny, raster.nx, _lib.U8, np.bool_)
        b, scalar, _ = _operand(self.value, raster)
        return dev.elementwise(_lib.OP_LT, raster, b, scalar, out)


class GreaterThan(DeviceFilter):
    """``image > value`` -> bool (simple_filters.py:53-96)."""

    def __init__(self, *, value):
        self.value = value

    def run_device(self, raster):
        out = dev.empty(raster.ny, raster.nx, _lib.U8, np.bool_)
        b, scalar, _ = _operand(self.value, raster)
        return dev.elementwise(_lib.OP_GT, raster, b, scalar, out)


class BooleanToInteger(DeviceFilter):
    """``image * 1`` (simple_filters.py:99-131): bool -> int64, other dtypes unchanged."""

    def run_device(self, raster):
        return binary_op(_lib.OP_MUL, raster, 1, np.multiply)


class ProductFilter(DeviceFilter):
    """``factor * image`` (simple_filters.py:134-180)."""

    def __init__(self, factor=1):
        self.factor = factor

    def run_device(self, raster):
        return binary_op(_lib.OP_MUL, raster, self.factor, np.multiply)


class AdditionFilter(DeviceFilter):
    """``addend + image`` (simple_filters.py:183-229)."""

    def __init__(self, addend=0):
        self.addend = addend

    def run_device(self, raster):
        return binary_op(_lib.OP_ADD, raster, self.addend, np.add)


class SubtractionFilter(DeviceFilter):
    """``minuend - image`` (simple_filters.py:232-275)."""

    def __init__(self, *, minuend=0.0):
        self.minuend = minuend

    def run_device(self, raster):
        return binary_op(_lib.OP_RSUB, raster, self.minuend, np.subtract)

    def apply(self, subtracting):  # pylint: disable=arguments-renamed
        # the reference skips the ndarray check here (simple_filters.py:261-275)
        if not isinstance(subtracting, np.ndarray):
            return self.minuend - subtracting
        return dev.download(self.run_device(dev.upload(subtracting)))
