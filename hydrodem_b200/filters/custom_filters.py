"""Windowed and composed filters of the conditioning chain, CUDA backed.

Drop-in for ``cguerrero/hydrodem/filters/custom_filters.py``: same class
names, keyword-only constructors, result dtypes, border conventions and
aliasing (the in-place filters still return the caller's own array).  The
per-cell Python loops over ``SlidingWindow`` are gone: each class is one or a
few kernels of libhydrodem_b200 working on device rasters.
"""
import numpy as np

from . import ComposedFilter, ComposedFilterResults, DeviceFilter, Filter, LazyResults, run_stage
from .. import _lib, device as dev
from ..exceptions import (NumpyArrayExpectedError, WindowSizeEvenError, WindowSizeHighError)
from .extension_filters import (AbsoluteValues, Around, BinaryClosing, BinaryErosion, BitwiseXOR, Convolve,  # noqa: F401
                                GreyDilation)
from .simple_filters import (AdditionFilter, BooleanToInteger, GreaterThan, LowerThan, ProductFilter,     # noqa: F401
                             SubtractionFilter)


def check_window(shape, window_size):
    """Guards of SlidingWindow.window_size (sliding_window.py:150-156): too large first, then even."""
    if any(window_size > n for n in shape):
        raise WindowSizeHighError(window_size, shape)
    if window_size % 2 != 1:
        raise WindowSizeEvenError(window_size)


def as_f32(raster):
    """The ``grid.astype('float32')`` every sliding window starts with (sliding_window.py:132)."""
    return dev.convert(raster, _lib.F32)


class WindowFilter(DeviceFilter):
    """A filter that builds SlidingWindow objects in the reference: the
    ndarray check comes from the grid setter (sliding_window.py:130-131)."""

    window_size = None

    def apply(self, image_to_filter):
        if not isinstance(image_to_filter, np.ndarray):
            raise NumpyArrayExpectedError(image_to_filter)
        check_window(image_to_filter.shape, self.window_size)
        return dev.download(self.run_device(dev.upload(image_to_filter)))


class MajorityFilter(WindowFilter):
    """Mode of the corner-less window if it fills more than 70 % of it
    (custom_filters.py:22-73).  float64 zeros elsewhere and on the border."""

    def __init__(self, *, window_size):
        self.window_size = window_size

    def run_device(self, raster):
        check_window(raster.shape, self.window_size)
        ws = int(self.window_size)
        min_count = int(np.floor((ws ** 2 - 1) * 0.7)) + 1        # count > (ws**2 - 1) * 0.7  (:71)
        src = as_f32(raster)
        out = dev.empty(raster.ny, raster.nx, _lib.F32, np.float64)
        _lib.check(_lib.load().hd_majority(src.ptr, src.pitch, out.ptr, out.dtype, out.pitch, src.ny, src.nx, ws,
                                           min_count, dev.stream_ptr()), window_size=ws, shape=raster.shape)
        return out


class ExpandFilter(WindowFilter):
    """1 where any cell of the corner-less window is > 0 (custom_filters.py:76-125)."""

    def __init__(self, *, window_size):
        self.window_size = window_size

    def run_device(self, raster):
        check_window(raster.shape, self.window_size)
        src = raster if raster.dtype in (_lib.U8, _lib.F32) else as_f32(raster)
        out = dev.empty(raster.ny, raster.nx, _lib.U8, np.float64)
        _lib.check(_lib.load().hd_expand(src.ptr, src.dtype, src.pitch, out.ptr, out.dtype, out.pitch, src.ny, src.nx,
                                         int(self.window_size), dev.stream_ptr()),
                   window_size=self.window_size, shape=raster.shape)
        return out


class _InPlace3(WindowFilter):
    """Shared plumbing of the two in-place 3x3 filters: Jacobi on the device
    (the reference reads a float32 snapshot, sliding_window.py:132), result
    copied back into the caller's array, which is returned."""

    _entry = None

    def run_device(self, raster):
        check_window(raster.shape, self.window_size)
        if int(self.window_size) != 3:
            raise dev.DeviceError(f"{type(self).__name__}: only window_size=3 is implemented on the device")
        if raster.dtype in (_lib.F32, _lib.F64):
            src = raster
        else:
            src = dev.convert(raster, _lib.F64)
        out = dev.empty(raster.ny, raster.nx, src.dtype, raster.ref_dtype)
        fn = getattr(_lib.load(), self._entry)
        _lib.check(fn(src.ptr, src.pitch, out.ptr, out.pitch, src.dtype, src.ny, src.nx, dev.stream_ptr()))
        return out

    def apply(self, image_to_filter):
        if not isinstance(image_to_filter, np.ndarray):
            raise NumpyArrayExpectedError(image_to_filter)
        check_window(image_to_filter.shape, self.window_size)
        return dev.download(self.run_device(dev.upload(image_to_filter)), out=image_to_filter)


class CorrectNANValues(_InPlace3):
    """Voids (< 0) replaced by the float32 mean of their valid neighbours,
    in place (custom_filters.py:260-317)."""

    _entry = "hd_nanfix"

    def __init__(self, *, window_size=3):
        self.window_size = window_size


class IsolatedPoints(_InPlace3):
    """Mask cells without any positive neighbour cleared, in place
    (custom_filters.py:320-366)."""

    _entry = "hd_isolated"

    def __init__(self, *, window_size):
        self.window_size = window_size


class MaskNegatives(ComposedFilter):
    """(image < 0) * 1 (custom_filters.py:465-486)."""

    def __init__(self):
        super().__init__()
        self.filters = [LowerThan(value=0.0), BooleanToInteger()]


class MaskPositives(ComposedFilter):
    """(image > 0) * 1 (custom_filters.py:489-510)."""

    def __init__(self):
        super().__init__()
        self.filters = [GreaterThan(value=0.0), BooleanToInteger()]


class MaskTallGroves(ComposedFilter):
    """(image > 1.5) * 1 (custom_filters.py:513-534)."""

    def __init__(self):
        super().__init__()
        self.filters = [GreaterThan(value=1.5), BooleanToInteger()]


class TidyingLagoons(ComposedFilter):
    """BinaryErosion(2) -> ExpandFilter(7) -> x majority image -> GreyDilation(7x7)
    (custom_filters.py:564-610)."""

    def __init__(self):
        super().__init__()
        self.filters = [BinaryErosion(iterations=2), ExpandFilter(window_size=7), ProductFilter(),
                        GreyDilation(size=(7, 7))]

    def run_device(self, raster):
        self.filters[2].factor = raster                     # the majority image is the factor (:607)
        return super().run_device(raster)

    def apply(self, image_to_filter):
        Filter.apply(self, image_to_filter)
        return dev.download(self.run_device(dev.upload(image_to_filter)))


class LagoonsDetection(ComposedFilterResults):
    """CorrectNANValues -> MajorityFilter(11) -> TidyingLagoons -> MaskPositives
    (custom_filters.py:613-661)."""

    def __init__(self):
        super().__init__()
        self.filters = [CorrectNANValues(), MajorityFilter(window_size=11), TidyingLagoons(), MaskPositives()]
        self.hsheds_nan_fixed = None
        self.mask_lagoons = None
        self.lagoons_values = None

    def apply(self, image_to_filter):
        Filter.apply(self, image_to_filter)
        check_window(image_to_filter.shape, 3)
        result = self._run(dev.upload(image_to_filter))
        # CorrectNANValues works in place: the caller's array is updated and IS results["CorrectNANValues"] (:658)
        fixed = dev.download(self.results.device("CorrectNANValues"), out=image_to_filter)
        self.results["CorrectNANValues"] = fixed
        self.hsheds_nan_fixed = fixed
        self.mask_lagoons = self.results["MaskPositives"]
        self.lagoons_values = self.results["TidyingLagoons"]
        return dev.download(result) if isinstance(result, dev.DeviceRaster) else result


class PostProcessingFinal(ComposedFilter):
    """Convolve() (3x3 mean, reflect) then Around() (custom_filters.py:1104-1125), fused in one kernel."""

    def __init__(self):
        super().__init__()
        self.filters = [Convolve(), Around()]

    def run_device(self, raster):
        conv, rnd = self.filters
        if type(conv) is Convolve and type(rnd) is Around:
            return conv.run_device(raster, do_round=True)
        return super().run_device(raster)

    def apply(self, image_to_filter):
        Filter.apply(self, image_to_filter)
        return dev.download(self.run_device(dev.upload(image_to_filter)))
