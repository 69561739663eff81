"""Filter protocol of the conditioning path -- the drop-in boundary.

Mirrors ``cguerrero/hydrodem/filters/__init__.py:10-127`` of the reference:
``Filter.apply(ndarray) -> ndarray``, ``ComposedFilter`` (a list of filters run
left to right) and ``ComposedFilterResults`` (same, keeping every
intermediate under the filter's class name).

What is new is underneath: every filter of this package also implements
``run_device(DeviceRaster) -> DeviceRaster``.  ``apply`` uploads the array
once, composed filters hand the device raster from stage to stage without
touching the host, and only the final result (or a ``results[...]`` entry
somebody actually reads) is copied back, converted on the device to the
dtype the reference would have returned.
"""
from abc import ABC, abstractmethod

import numpy as np

from .. import device as dev
from ..exceptions import NumpyArrayExpectedError


class Filter(ABC):
    """Base class: anything with ``apply(ndarray) -> ndarray``."""

    @abstractmethod
    def apply(self, image_to_filter):
        # reference: filters/__init__.py:38-39
        if not isinstance(image_to_filter, np.ndarray):
            raise NumpyArrayExpectedError(image_to_filter)


class DeviceFilter(Filter):
    """A filter whose work is a chain of CUDA kernels on a device raster."""

    def run_device(self, raster):
        raise NotImplementedError

    def apply(self, image_to_filter):
        Filter.apply(self, image_to_filter)
        return dev.download(self.run_device(dev.upload(image_to_filter)))


def run_stage(filter_, content):
    """Run one stage of a composition on ``content`` (ndarray or DeviceRaster).

    Filters of this package stay on the device; a foreign host-only Filter is
    fed a downloaded array and its result is uploaded again."""
    if isinstance(filter_, DeviceFilter):
        if isinstance(content, np.ndarray):
            content = dev.upload(content)
        return filter_.run_device(content)
    if isinstance(content, dev.DeviceRaster):
        content = dev.download(content)
    return filter_.apply(content)


def to_host(content):
    return dev.download(content) if isinstance(content, dev.DeviceRaster) else content


class ComposedFilter(DeviceFilter):
    """``filters`` applied left to right (reference: filters/__init__.py:42-80)."""

    def __init__(self):
        self.filters = []

    def run_device(self, raster):
        content = raster
        for filter_ in self.filters:
            content = run_stage(filter_, content)
        return content if isinstance(content, dev.DeviceRaster) else dev.upload(content)

    def apply(self, image_to_filter):
        Filter.apply(self, image_to_filter)
        content = image_to_filter
        for filter_ in self.filters:
            content = run_stage(filter_, content)
        return to_host(content)


class LazyResults(dict):
    """``results`` of a ComposedFilterResults: device rasters are copied to the
    host (in the reference's dtype) the first time an entry is read."""

    def __getitem__(self, key):
        value = dict.__getitem__(self, key)
        if isinstance(value, dev.DeviceRaster):
            value = dev.download(value)
            dict.__setitem__(self, key, value)
        return value

    def device(self, key):
        """The raw stored value (DeviceRaster if it has not been read yet)."""
        return dict.__getitem__(self, key)

    def get(self, key, default=None):
        return self[key] if key in self else default

    def items(self):
        return [(k, self[k]) for k in self.keys()]

    def values(self):
        return [self[k] for k in self.keys()]


class ComposedFilterResults(DeviceFilter):
    """Like ComposedFilter, keeping ``results[ClassName]`` per stage
    (reference: filters/__init__.py:83-127)."""

    def __init__(self):
        self.filters = []
        self.results = LazyResults()

    def _run(self, content):
        for filter_ in self.filters:
            content = run_stage(filter_, content)
            self.results[filter_.__class__.__name__] = content
        return content

    def run_device(self, raster):
        content = self._run(raster)
        return content if isinstance(content, dev.DeviceRaster) else dev.upload(content)

    def apply(self, image_to_filter):
        Filter.apply(self, image_to_filter)
        return to_host(self._run(image_to_filter))
