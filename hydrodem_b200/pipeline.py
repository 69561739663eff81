"""The full conditioning chain, device resident.

Host-side mirror of ``HydroDEMProcess.start`` (hydro_dem_process.py:122-153)
with the GDAL file handling removed: the four rasters that ``image_srtm.py``
and ``image_hsheds.py`` read with ``gdal.Open(...).ReadAsArray()`` come in as
NumPy arrays, everything in between runs as CUDA kernels on device rasters,
and the results come back as NumPy arrays in the reference's dtypes.

    SRTM branch    DetectApplyFourier                       image_srtm.py:125-126
                   BinaryClosing(ones(3,3)) on the groves   image_srtm.py:177-178
                   GrovesCorrectionsIter(iterations=3)      image_srtm.py:199
    HSHEDS branch  LagoonsDetection                         image_hsheds.py:133-136
                   (river routing is out of scope: ``rivers`` is a given 0/1 raster, default none)
    combine        _prepare_final_terms + sum               hydro_dem_process.py:60-91, :148
                   PostProcessingFinal                      hydro_dem_process.py:149
    NEW            SinkFill + D8FlowDirection on the final DEM

Inside the chain rasters are stored as float32 / uint8 (the reference's
float64 intermediates hold float32-representable or tolerance-class values);
conversion to the reference dtype happens once, on download.
"""
import ctypes

import numpy as np

from . import _lib, device as dev
from .exceptions import NumpyArrayExpectedError
from .filters import custom_filters as cf
from .filters import extension_filters as ef
from .filters import new_filters as nf


class ChainResult:
    """Device rasters of one run; ``.host(name)`` / attribute access copies to NumPy lazily."""

    def __init__(self, rasters, info):
        self.rasters = rasters
        self.info = info
        self._host = {}

    def host(self, name):
        if name not in self._host:
            self._host[name] = dev.download(self.rasters[name])
        return self._host[name]

    def __getattr__(self, name):
        rasters = self.__dict__.get("rasters", {})
        if name in rasters:
            return self.host(name)
        raise AttributeError(name)


class CapturedChain:
    """A CUDA graph of one pass of the chain (see ConditioningChain.capture)."""

    def __init__(self, graph, result, launches):
        self.graph, self.result, self.launches = graph, result, launches

    def replay(self):
        self.graph.replay()
        self.result._host.clear()
        return self.result


class ConditioningChain:
    """Run the whole chain on the GPU.

    >>> out = ConditioningChain().apply(srtm, groves_class, hsheds)
    >>> out.final, out.filled, out.d8
    """

    def __init__(self, *, groves_iterations=3, with_hydrology=True, keep_intermediates=False, fill_stats=False):
        self.fill_stats = fill_stats          # True: SinkFill synchronises once to report its tile-visit count
        self.groves_iterations = groves_iterations
        self.with_hydrology = with_hydrology
        self.keep_intermediates = keep_intermediates

    # ---- device path --------------------------------------------------------------------------------
    def run_device(self, srtm, groves_class, hsheds, rivers=None, ready=None, on_ready=None):
        """``ready``: optional {name: torch.cuda.Event} -- the compute stream waits for an input's upload only where
        that input is first used.  ``on_ready(name, raster)`` is called as soon as an output raster is enqueued."""
        import torch
        lib = _lib.load()
        ny, nx = srtm.shape
        out = {}
        info = {}
        cur = torch.cuda.current_stream()

        def need(name):
            if ready and ready.get(name) is not None:
                cur.wait_event(ready[name])

        def emit(name, raster):
            out[name] = raster
            if on_ready:
                on_ready(name, raster)

        # SRTM branch
        need("srtm")
        daf = cf.DetectApplyFourier()
        corrected = daf.run_device(srtm)                                   # F32 storage, ref float64
        need("groves")
        groves = ef.BinaryClosing(structure=np.ones((3, 3))).run_device(groves_class)   # U8 0/1
        dem = corrected
        gc = cf.GrovesCorrection(groves)
        for _ in range(self.groves_iterations):
            dem = gc.run_device(dem, out_dtype=_lib.F32)
        # HSHEDS branch: LagoonsDetection (custom_filters.py:633-661) with float32 / uint8 intermediates
        need("hsheds")
        fixed = cf.CorrectNANValues().run_device(hsheds)
        majority = cf.MajorityFilter(window_size=11).run_device(fixed)
        eroded = ef.BinaryErosion(iterations=2).run_device(majority)
        prod = dev.empty(ny, nx, _lib.F32, np.float64)                      # majority * expand(7): one fused kernel
        _lib.check(lib.hd_expand_select(eroded.ptr, eroded.dtype, eroded.pitch, majority.ptr, majority.pitch, prod.ptr,
                                        prod.pitch, ny, nx, 7, dev.stream_ptr()))
        tidy = ef.GreyDilation(size=(7, 7)).run_device(prod)                # lagoons_values
        # combine + post-processing, float64 like the reference
        fixed32 = dev.convert(fixed, _lib.F32)
        if rivers is not None:
            need("rivers")
        riv = dev.convert(rivers, _lib.F32) if rivers is not None else None
        complete = dev.empty(ny, nx, _lib.F64, np.float64)
        _lib.check(lib.hd_final_terms(dem.ptr, dem.dtype, dem.pitch, tidy.ptr, tidy.pitch, fixed32.ptr, fixed32.pitch,
                                      riv.ptr if riv is not None else None, riv.pitch if riv is not None else 0,
                                      complete.ptr, complete.dtype, complete.pitch, ny, nx, dev.stream_ptr()))
        final32 = dev.empty(ny, nx, _lib.F32, np.float32) if self.with_hydrology else None   # integer metres: exact
        final = cf.PostProcessingFinal().run_device(complete, copy32=final32)
        emit("final", final)
        if self.with_hydrology:
            fill = nf.SinkFill(want_stats=self.fill_stats)
            emit("filled", fill.run_device(final32))
            info["fill_sweeps"] = fill.sweeps
            emit("d8", nf.D8FlowDirection().run_device(out["filled"]))
        if self.keep_intermediates:
            out.update(fourier=corrected, fourier_mask=daf._mask_dev, fabs=daf._fabs_dev, groves=groves, srtm=dem,
                       hsheds_nan_fixed=fixed, majority=majority, lagoons_values=tidy, dem_complete=complete)
        return ChainResult(out, info)

    def capture(self, srtm, groves_class, hsheds, rivers=None):
        """Capture the whole device-resident chain for these (static) input rasters in ONE CUDA graph.

        The chain is ~40 short kernel launches; issued one by one from Python the host cannot keep a B200 busy
        (~4 ms of interpreter + ctypes time per pass).  A captured graph replays them with a single launch.  Returns a
        CapturedChain: ``replay()`` re-runs the chain on whatever the input rasters hold at that moment and returns
        the same ChainResult (its rasters are the graph's static output buffers)."""
        import torch
        self.run_device(srtm, groves_class, hsheds, rivers)          # warm-up: FFT plans, function attributes
        torch.cuda.synchronize()
        lib = _lib.load()
        n0 = lib.hd_launch_count()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            result = self.run_device(srtm, groves_class, hsheds, rivers)
        return CapturedChain(graph, result, int(lib.hd_launch_count() - n0))

    # ---- host API ---------------------------------------------------------------------------------
    def upload_inputs(self, srtm_raw, groves_class_raw, hsheds, rivers=None):
        for a in (srtm_raw, groves_class_raw, hsheds) + ((rivers,) if rivers is not None else ()):
            if not isinstance(a, np.ndarray):
                raise NumpyArrayExpectedError(a)
        if not (srtm_raw.shape == groves_class_raw.shape == hsheds.shape):
            raise ValueError("srtm, groves_class and hsheds must have the same shape")
        return (dev.upload(srtm_raw), dev.upload(groves_class_raw), dev.upload(hsheds),
                dev.upload(rivers) if rivers is not None else None)

    def apply(self, srtm_raw, groves_class_raw, hsheds, rivers=None):
        """ndarrays in -> ChainResult (``final`` float64, ``filled`` float32, ``d8`` uint8); results are copied to
        the host lazily, on first access."""
        return self.run_device(*self.upload_inputs(srtm_raw, groves_class_raw, hsheds, rivers))

    def apply_to_host(self, srtm_raw, groves_class_raw, hsheds, rivers=None):
        """ndarrays in -> dict of ndarrays (``final``, ``filled``, ``d8``) with the PCIe traffic overlapped with the
        kernels: inputs go up on a copy stream (the Fourier stage starts as soon as the SRTM raster has landed,
        HydroSHEDS and groves follow underneath it), and each result starts its way down on a second copy stream
        the moment its last kernel is enqueued (the final DEM travels while the sink-fill runs).  Pageable inputs
        are staged through pinned buffers first."""
        import torch
        arrays = dict(srtm=srtm_raw, groves=groves_class_raw, hsheds=hsheds)
        if rivers is not None:
            arrays["rivers"] = rivers
        for a in arrays.values():
            if not isinstance(a, np.ndarray):
                raise NumpyArrayExpectedError(a)
        if not (srtm_raw.shape == groves_class_raw.shape == hsheds.shape):
            raise ValueError("srtm, groves_class and hsheds must have the same shape")
        dev.require_cuda()
        cur = torch.cuda.current_stream()
        if not hasattr(self, "_streams"):
            self._streams = (torch.cuda.Stream(), torch.cuda.Stream())
        up, down = self._streams
        up.wait_stream(cur)
        down.wait_stream(cur)
        rasters, ready, keep = {}, {}, []
        for name in ("srtm", "hsheds", "groves", "rivers"):
            if name not in arrays:
                continue
            host = np.ascontiguousarray(arrays[name])
            if not dev._is_pinned(host):
                pin = dev.pinned_empty(host.shape, host.dtype)
                pin[...] = host
                host = pin
            keep.append(host)
            rasters[name], ready[name] = dev.upload_async(host, up)
            rasters[name].buf.record_stream(cur)
        pending = {}

        def on_ready(name, raster):
            ev = torch.cuda.Event()
            ev.record(cur)
            down.wait_event(ev)
            pending[name] = dev.download_async(raster, down)

        self.run_device(rasters["srtm"], rasters["groves"], rasters["hsheds"], rasters.get("rivers"), ready=ready,
                        on_ready=on_ready)
        result = {}
        for name, (host, ev) in pending.items():
            ev.synchronize()
            result[name] = host
        cur.wait_stream(down)
        del keep
        return result
