"""Flood-extent scores of simulated water-depth rasters against an NDWI water mask: the reference's ``Stats`` class
(stats.py:6-93) with its six counts taken by one CUDA pass per raster
(``hd_confusion_counts``) instead of a dozen full-size NumPy temporaries.

The reference binds the class to ``Config`` keys and GDAL file names; here the mask and the simulated rasters are given as
arrays, device rasters or GeoTIFF paths.  Everything else is kept: the method names, the order of ``stats_functions``,
the keys of ``get_stats`` (``'day'`` and ``<function>_<sufix>``), ``f1_score`` with FP counted twice (stats.py:50-52).
One deliberate difference: the counts are Python ints, which is what ``np.count_nonzero`` returned when the reference was
written.  Under NumPy >= 2 it returns ``numpy.int64`` and the reference's MCC denominator (a product of four counts,
stats.py:57-60) wraps silently for rasters with more than ~55 000 cells per class; here it is exact, and an empty
denominator raises ``ZeroDivisionError`` instead of returning inf / nan with a warning.  On rasters small enough not to
overflow (tests/golden/run_stats.npz) all eight scores are identical to the reference's.
"""
import ctypes
import math

import numpy as np

from . import _lib, device as dev
from .exceptions import NumpyArrayExpectedError


def _to_device(x):
    if isinstance(x, dev.DeviceRaster):
        return x
    if isinstance(x, (str, bytes)) or hasattr(x, "__fspath__"):
        from . import geotiff
        return geotiff.read_to_device(x)
    if not isinstance(x, np.ndarray):
        raise NumpyArrayExpectedError(x)
    return dev.upload(x)


class Stats:

    THRESHOLD = 0.0001                                        # stats.py:68

    def __init__(self, ndwi, files=(), sufix=""):
        """``ndwi``: the water mask (uint8 / int16 / float32 array, device raster or GeoTIFF path); ``files``: the
        simulated rasters of day 1, 2, ... (float32 / float64, same forms)."""
        import torch
        self.sufix = '_' + sufix
        self.files = list(files)
        self.range_files = len(self.files) + 1                 # the reference iterates range(1, range_files)
        self.ndwi = _to_device(ndwi)
        if self.ndwi.dtype not in (_lib.U8, _lib.I16, _lib.F32):
            self.ndwi = dev.convert(self.ndwi, _lib.F32)
        self._counts = torch.zeros(6, dtype=torch.int64, device=dev.device())
        self.total_positives = self.total_negatives = None     # filled by the first counting pass (same kernel)
        self.values_file = dict()
        self.stats_functions = [self.accuracy, self.sensitivity, self.BACC, self.f1_score, self.MCC, self.precision,
                                self.specificity, self.fall_out]

    @property
    def total_values(self):
        return self.total_positives + self.total_negatives

    def accuracy(self):
        return (self.values_file['TN'] + self.values_file['TP']) / self.total_values

    def sensitivity(self):
        return self.values_file['TP'] / self.values_file['P']

    def precision(self):
        return self.values_file['TP'] / (self.values_file['TP'] + self.values_file['FP'])

    def specificity(self):
        return self.values_file['TN'] / self.values_file['N']

    def fall_out(self):
        return self.values_file['FP'] / (self.values_file['FP'] + self.values_file['TN'])

    def BACC(self):
        return (self.sensitivity() + self.specificity()) / 2.

    def f1_score(self):
        return (2. * self.values_file['TP']) / (2. * self.values_file['TP'] + self.values_file['FP'] +
                                                self.values_file['FP'])

    def MCC(self):
        v = self.values_file
        numerator_mcc = (v['TP'] * v['TN']) - (v['FP'] * v['FN'])
        denominator_mcc = (v['TP'] + v['FP']) * (v['TP'] + v['FN']) * (v['TN'] + v['FP']) * (v['TN'] + v['FN'])
        return numerator_mcc / math.sqrt(denominator_mcc)

    def _set_values(self, file):
        """Stats._set_values (stats.py:63-86) for one simulated raster (array, device raster or path)."""
        sim = _to_device(file)
        if sim.dtype not in (_lib.F32, _lib.F64):
            sim = dev.convert(sim, _lib.F32)
        if sim.shape != self.ndwi.shape:
            raise ValueError(f"simulated raster {sim.shape} and NDWI mask {self.ndwi.shape} differ in shape")
        t = self.ndwi
        _lib.check(_lib.load().hd_confusion_counts(sim.ptr, sim.dtype, sim.pitch, t.ptr, t.dtype, t.pitch, t.ny, t.nx,
                                                   ctypes.c_double(self.THRESHOLD),
                                                   ctypes.c_void_p(self._counts.data_ptr()), dev.stream_ptr()))
        tp, fn, fp, tn, pos, neg = (int(c) for c in self._counts.cpu().tolist())
        self.total_positives, self.total_negatives = pos, neg
        self.values_file.update(TP=tp, FN=fn, P=tp + fn, FP=fp, TN=tn, N=fp + tn)

    def get_stats(self):
        list_stats = []
        for i in range(1, self.range_files):
            stats_file = {'day': i}
            self._set_values(self.files[i - 1])
            for f in self.stats_functions:
                stats_file[f.__name__ + self.sufix] = f()
            list_stats.append(stats_file)
        return list_stats
