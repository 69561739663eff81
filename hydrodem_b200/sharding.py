"""Row-band sharding of one large mosaic across the GPUs of a box.

A mosaic of ``ny`` rows is cut into ``world`` contiguous row bands (one per
rank / GPU).  A windowed filter with half-window ``h`` needs ``h`` rows of
its neighbours' bands: before each stage the ranks exchange halo rows with
``torch.distributed`` P2P (NCCL send/recv over NVLink on the GPU box, gloo in
the CPU tests), run the unchanged single-GPU kernel on ``[halo | band | halo]``
and keep the band rows.  At the true top / bottom of the mosaic there is no
halo, so the filter's own border convention applies exactly as on one GPU --
results of the exact-class stages are bit-identical to the unsharded run
(tolerance-class stages agree to float32 rounding: their per-tile arithmetic
depends on where tile boundaries fall).

Sink-fill is iterative: every rank relaxes its band (plus one halo row) to a
local fixed point, the ranks exchange their edge rows of W, and the loop ends
when an all-reduce says no halo row was lowered.  The fixed point is unique,
so the banded result equals the single-GPU one bit for bit.

The Fourier transforms of a banded mosaic are distributed too (``Band.fft2`` /
``Band.ifft2``: local row transforms, ONE all-to-all, local column
transforms).  The peak detector and the point-mirrored mask assembly between
them still run on one GPU, so the fully banded path
covers the stencil stages, the transforms and sink-fill / D8.

The communicator is abstract: ``DistComm`` (torch.distributed) for real runs,
``ThreadComm`` to emulate the ranks as threads of one process on one device
(used by the GPU tests; the host logic is additionally tested with gloo).
"""
import ctypes
import threading

import numpy as np
import torch

from . import _lib, device as dev


def band_bounds(ny, world):
    """Row range [r0, r1) of every rank: bands differ by at most one row."""
    base, rem = divmod(int(ny), int(world))
    out, r = [], 0
    for k in range(world):
        n = base + (1 if k < rem else 0)
        out.append((r, r + n))
        r += n
    return out


# ---- communicators -------------------------------------------------------------------------------
class DistComm:
    """Halo exchange over torch.distributed (NCCL on GPUs, gloo on CPU)."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)

    def exchange(self, send_up, send_down, recv_up_like, recv_down_like):
        """Send ``send_up`` to rank-1 and ``send_down`` to rank+1; return (from rank-1, from rank+1).
        Arguments / results are None at the ends of the mosaic."""
        dist = self.dist
        ops, recv_up, recv_down = [], None, None
        if self.rank > 0:
            recv_up = torch.empty_like(recv_up_like)
            ops.append(dist.P2POp(dist.isend, send_up.contiguous(), self.rank - 1, self.group))
            ops.append(dist.P2POp(dist.irecv, recv_up, self.rank - 1, self.group))
        if self.rank < self.world - 1:
            recv_down = torch.empty_like(recv_down_like)
            ops.append(dist.P2POp(dist.isend, send_down.contiguous(), self.rank + 1, self.group))
            ops.append(dist.P2POp(dist.irecv, recv_down, self.rank + 1, self.group))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        return recv_up, recv_down

    def all_to_all_shaped(self, chunks, recv_shapes):
        """chunks[j] goes to rank j; returns the chunks received (index = source rank).  The shapes that arrive are
        known from the global band table (``recv_shapes``), so no size exchange is needed.  Grouped isend / irecv
        pairs: NCCL fuses them into one all-to-all over NVLink, gloo runs them as P2P."""
        dist = self.dist
        out = [None] * self.world
        ops = []
        for j in range(self.world):
            if j == self.rank:
                out[j] = chunks[j].contiguous()
                continue
            out[j] = torch.empty(recv_shapes[j], dtype=chunks[j].dtype, device=chunks[j].device)
            ops.append(dist.P2POp(dist.isend, chunks[j].contiguous(), j, self.group))
            ops.append(dist.P2POp(dist.irecv, out[j], j, self.group))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        return out

    def any(self, flag):
        t = torch.tensor([1 if flag else 0], dtype=torch.int32,
                         device="cuda" if self.dist.get_backend(self.group) == "nccl" else "cpu")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX, group=self.group)
        return bool(t.item())


class ThreadComm:
    """Emulates ``world`` ranks as threads of one process (one GPU): used to test the banded algorithms
    without several devices.  ``ThreadComm.run(world, fn)`` starts fn(comm) per rank."""

    class _Shared:
        def __init__(self, world):
            self.world = world
            self.barrier = threading.Barrier(world)
            self.box = {}
            self.flags = [0] * world

    def __init__(self, shared, rank):
        self.shared, self.rank, self.world = shared, rank, shared.world

    def exchange(self, send_up, send_down, recv_up_like, recv_down_like):
        sh = self.shared
        if torch.cuda.is_available():
            torch.cuda.synchronize()
        if self.rank > 0:
            sh.box[(self.rank, "up")] = send_up.clone()
        if self.rank < self.world - 1:
            sh.box[(self.rank, "down")] = send_down.clone()
        sh.barrier.wait()
        recv_up = sh.box[(self.rank - 1, "down")] if self.rank > 0 else None
        recv_down = sh.box[(self.rank + 1, "up")] if self.rank < self.world - 1 else None
        sh.barrier.wait()
        return recv_up, recv_down

    def all_to_all_shaped(self, chunks, recv_shapes):
        sh = self.shared
        if torch.cuda.is_available():
            torch.cuda.synchronize()
        for j in range(self.world):
            sh.box[("a2a", self.rank, j)] = chunks[j].clone()
        sh.barrier.wait()
        out = [sh.box[("a2a", i, self.rank)] for i in range(self.world)]
        sh.barrier.wait()
        for i, t in enumerate(out):
            assert tuple(t.shape) == tuple(recv_shapes[i]), (t.shape, recv_shapes[i])
        return out

    def any(self, flag):
        sh = self.shared
        sh.flags[self.rank] = 1 if flag else 0
        sh.barrier.wait()
        out = any(sh.flags)
        sh.barrier.wait()
        return out

    @staticmethod
    def run(world, fn):
        shared = ThreadComm._Shared(world)
        results, errors = [None] * world, []

        def work(rank):
            try:
                results[rank] = fn(ThreadComm(shared, rank))
            except BaseException as exc:           # noqa: BLE001
                errors.append(exc)
                shared.barrier.abort()

        threads = [threading.Thread(target=work, args=(r,)) for r in range(world)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        if errors:
            raise errors[0]
        return results


# ---- banded stages ---------------------------------------------------------------------------------------
class Band:
    """One rank's view of the mosaic."""

    def __init__(self, comm, ny, nx):
        self.comm, self.ny, self.nx = comm, int(ny), int(nx)
        self.r0, self.r1 = band_bounds(ny, comm.world)[comm.rank]

    @property
    def rows(self):
        return self.r1 - self.r0

    def take(self, mosaic):
        """This rank's rows of a host mosaic (tests / loaders)."""
        return np.ascontiguousarray(mosaic[self.r0:self.r1])

    # -- halo plumbing
    def extend(self, raster, h):
        """[halo_up | band | halo_down] as a new device raster, plus the number of halo rows on top."""
        t = raster.tensor()
        h_up = min(h, t.shape[0])
        send_up, send_down = t[:h_up], t[t.shape[0] - h_up:]
        recv_up, recv_down = self.comm.exchange(send_up, send_down, send_up, send_down)
        n_up = recv_up.shape[0] if recv_up is not None else 0
        n_down = recv_down.shape[0] if recv_down is not None else 0
        ext = dev.empty(raster.ny + n_up + n_down, raster.nx, raster.dtype, raster.ref_dtype)
        te = ext.tensor()
        if n_up:
            te[:n_up].copy_(recv_up)
        te[n_up:n_up + raster.ny].copy_(t)
        if n_down:
            te[n_up + raster.ny:].copy_(recv_down)
        return ext, n_up

    def apply(self, filt, raster, h):
        """Run a single-GPU filter (anything with ``run_device``) on the band with an h-row halo."""
        ext, n_up = self.extend(raster, h)
        out = filt.run_device(ext)
        return out.sub(n_up, n_up + raster.ny, 0, raster.nx)

    # -- distributed 2-D Fourier transform (FourierTransform / FourierITransform, extension_filters.py:363-480)
    def _rows_fft(self, src, n, inverse, transpose_out):
        """hd_fft_rows on a local (rows x n) raster -> C64 raster, (n x rows) when transposed."""
        lib = _lib.load()
        rows = src.ny
        plan = dev.fft_plan(n, n)
        nbytes = lib.hd_fft2_workspace_bytes(rows, n)
        work = dev.scratch(nbytes)
        out = dev.empty(n, rows, _lib.C64, np.complex64) if transpose_out else dev.empty(rows, n, _lib.C64, np.complex64)
        _lib.check(lib.hd_fft_rows(plan, src.ptr, src.dtype, src.pitch, out.ptr, out.pitch, rows, int(inverse),
                                   int(transpose_out), ctypes.c_void_p(work.data_ptr()), nbytes, dev.stream_ptr()))
        return out

    def _exchange_transposed(self, t_raster, my_bounds, other_bounds):
        """t_raster: (n_other x my_len) -- the transposed local block.  Sends rows [a, b) of it to the rank that owns
        [a, b) of the other axis; returns the (my_other_len x total_len) raster assembled from what arrives."""
        comm = self.comm
        tt = t_raster.tensor()
        chunks = [tt[a:b] for (a, b) in other_bounds]
        mine = other_bounds[comm.rank][1] - other_bounds[comm.rank][0]
        shapes = [(mine, b - a) for (a, b) in my_bounds]
        got = comm.all_to_all_shaped([torch.view_as_real(c.contiguous()) for c in chunks],
                                     [sh + (2,) for sh in shapes])
        total = sum(b - a for (a, b) in my_bounds)
        out = dev.empty(mine, total, _lib.C64, np.complex64)
        to = out.tensor()
        for (a, b), g in zip(my_bounds, got):
            to[:, a:b].copy_(torch.view_as_complex(g))
        return out

    def fft2(self, band):
        """Forward 2-D DFT of the mosaic whose row band this rank holds (F32 or C64 device raster, rows r0:r1).
        Returns the spectrum in TRANSPOSED band layout: a C64 raster (c1 - c0, ny) with out[c - c0, k] = F[k, c] for
        this rank's column band [c0, c1) = band_bounds(nx, world)[rank].  One all-to-all, as in SURVEY.md section 8(e):
        local row transforms -> exchange -> local column transforms."""
        rows_b, cols_b = band_bounds(self.ny, self.comm.world), band_bounds(self.nx, self.comm.world)
        t = self._rows_fft(band, self.nx, False, True)                    # (nx, rows) transposed row spectra
        cols = self._exchange_transposed(t, rows_b, cols_b)               # (my cols, ny)
        return self._rows_fft(cols, self.ny, False, False)                # transform along ny, stays (my cols, ny)

    def ifft2(self, spec_t):
        """Inverse of fft2: transposed-layout spectrum band (c1 - c0, ny) -> C64 row band (r1 - r0, nx)."""
        rows_b, cols_b = band_bounds(self.ny, self.comm.world), band_bounds(self.nx, self.comm.world)
        t = self._rows_fft(spec_t, self.ny, True, True)                   # (ny, my cols)
        rows = self._exchange_transposed(t, cols_b, rows_b)               # (my rows, nx)
        return self._rows_fft(rows, self.nx, True, False)

    # rows of context a cell of MaskFourier depends on: two hollow-mean passes (27 each), IsolatedPoints (1), Expand 13 (6)
    _MASK_HALO = 2 * 27 + 1 + 6

    def _quarter_mask(self, quarter):
        """MaskFourier (custom_filters.py:537-561) of one spectrum quarter with the COMPUTE split over the ranks: every
        rank holds the whole |F| (all-gathered), runs the detector on its rows of the quarter plus 61 rows of context
        cut from its own copy -- windows are clipped only at the true edges of the quarter, rows near a cut are
        discarded -- and the U8 mask slabs are all-gathered."""
        from .filters import custom_filters as cf
        comm = self.comm
        qh, qw = quarter.shape
        bounds = band_bounds(qh, comm.world)
        if min(b - a for a, b in bounds) < 16:                     # tiny quarters: not worth cutting
            return cf.MaskFourier().run_device(quarter)
        a, b = bounds[comm.rank]
        lo, hi = max(0, a - self._MASK_HALO), min(qh, b + self._MASK_HALO)
        m = cf.MaskFourier().run_device(quarter.sub(lo, hi, 0, qw))
        mine = dev.convert(m, _lib.U8).tensor()[a - lo:b - lo].contiguous()
        parts = comm.all_to_all_shaped([mine] * comm.world, [(bb - aa, qw) for (aa, bb) in bounds])
        full = dev.empty(qh, qw, _lib.U8, np.float64)
        full.tensor().copy_(torch.cat(parts, dim=0))
        return full

    def detect_apply_fourier(self, band):
        """DetectApplyFourier (custom_filters.py:1053-1101) on a banded mosaic: distributed forward transform, the
        |F| bands all-gathered (every rank holds the whole magnitude spectrum), the window-55 peak detector computed in
        row slabs of the quarters split over the ranks and its masks all-gathered, point-mirrored mask assembly, mask
        applied to the local spectrum band, distributed inverse, abs.
        Returns this rank's rows of the stripe-free DEM (F32 storage, float64 reference dtype)."""
        from .filters import custom_filters as cf, extension_filters as ef
        comm = self.comm
        cols_b = band_bounds(self.nx, comm.world)
        c0, c1 = cols_b[comm.rank]
        spec_t = self.fft2(band)                                           # (c1 - c0, ny): F[k, c] at [c - c0, k]
        fabs_loc = dev.empty(spec_t.ny, spec_t.nx, _lib.F32, np.float32)
        dev.elementwise(_lib.OP_ABS, spec_t, None, 0.0, fabs_loc)          # AbsoluteValues, extension_filters.py:78-95
        part = fabs_loc.tensor().contiguous()
        parts = comm.all_to_all_shaped([part] * comm.world, [(b - a, self.ny) for (a, b) in cols_b])   # all-gather
        fabs = dev.empty(self.ny, self.nx, _lib.F32, np.float32)
        fabs.tensor().copy_(torch.cat(parts, dim=0).t())                   # |F| in natural (ny, nx) layout
        fabs_shift = ef.FourierShift().run_device(fabs)                    # FourierInitial, custom_filters.py:859-877
        keep = cf.FourierProcessQuarters(fabs_shift).run_device(fabs_shift, invert=True, out_dtype=_lib.F32,
                                                                mask_fn=self._quarter_mask)            # 1 - mask
        keep = ef.FourierIShift().run_device(keep)                         # back to the unshifted layout of F
        keep_t = dev.empty(spec_t.ny, spec_t.nx, _lib.F32, np.float32)
        keep_t.tensor().copy_(keep.tensor()[:, c0:c1].t())                 # this rank's columns, transposed
        masked = dev.empty(spec_t.ny, spec_t.nx, _lib.C64, np.complex64)
        dev.elementwise(_lib.OP_MUL, spec_t, keep_t, 0.0, masked)          # ProductFilter(factor=F), (1 - mask) * F
        back = self.ifft2(masked)                                          # (r1 - r0, nx) complex
        out = dev.empty(back.ny, back.nx, _lib.F32, np.float64)
        dev.elementwise(_lib.OP_ABS, back, None, 0.0, out)
        return out

    # -- the whole chain on row bands (BASELINE.json configs[4]: one mosaic over the GPUs of a box)
    def conditioning_chain(self, srtm, groves_class, hsheds, groves_iterations=3, with_hydrology=True):
        """HydroDEMProcess.start (hydro_dem_process.py:122-153) on this rank's rows of a mosaic: every windowed stage
        runs the single-GPU kernel on [halo | band | halo] after a halo exchange with the two neighbours (h rows:
        closing 2, quadratic 7, nanfix 1, majority 5, erosion x2 2, expand 3, max 7x7 3, mean3 1), the Fourier stage
        uses the distributed transforms, the sink-fill iterates to the global fixed point.  Inputs: device rasters of
        the band (F32, U8 0/1, F32).  Returns {"final": F64-ref raster, "filled", "d8"} for the band's rows."""
        from .filters import custom_filters as cf, extension_filters as ef
        lib = _lib.load()
        ny_b, nx = srtm.ny, srtm.nx
        dem = self.detect_apply_fourier(srtm)                                        # image_srtm.py:125-126
        groves = self.apply(ef.BinaryClosing(structure=np.ones((3, 3))), groves_class, 2)    # image_srtm.py:177-178
        g_ext, g_up = self.extend(dev.convert(groves, _lib.U8), 7)
        gc = cf.GrovesCorrection(g_ext)
        for _ in range(groves_iterations):                                           # image_srtm.py:199
            ext, n_up = self.extend(dem, 7)
            assert n_up == g_up
            dem = gc.run_device(ext, out_dtype=_lib.F32).sub(n_up, n_up + ny_b, 0, nx)
        fixed = self.apply(cf.CorrectNANValues(), hsheds, 1)                         # LagoonsDetection, :633-661
        majority = self.apply(cf.MajorityFilter(window_size=11), fixed, 5)
        eroded = self.apply(ef.BinaryErosion(iterations=2), majority, 2)
        expanded = self.apply(cf.ExpandFilter(window_size=7), eroded, 3)
        prod = dev.empty(ny_b, nx, _lib.F32, np.float64)
        dev.elementwise(_lib.OP_MUL, expanded, majority, 0.0, prod)                  # ProductFilter(factor=majority), :607
        tidy = self.apply(ef.GreyDilation(size=(7, 7)), prod, 3)
        fixed32 = dev.convert(fixed, _lib.F32)
        complete = dev.empty(ny_b, nx, _lib.F64, np.float64)
        _lib.check(lib.hd_final_terms(dem.ptr, dem.dtype, dem.pitch, tidy.ptr, tidy.pitch, fixed32.ptr, fixed32.pitch,
                                      None, 0, complete.ptr, complete.dtype, complete.pitch, ny_b, nx, dev.stream_ptr()))
        final = self.apply(cf.PostProcessingFinal(), complete, 1)                    # hydro_dem_process.py:149
        out = {"final": final, "dem_complete": complete}
        if with_hydrology:
            out["filled"], out["d8"] = self.sinkfill(dev.convert(final, _lib.F32, np.float32))
        return out

    # -- sink-fill + D8
    def sinkfill(self, z, max_rounds=10000):
        """Banded Planchon-Darboux fixed point (bit-identical to the single-GPU result)."""
        lib = _lib.load()
        zext, n_up = self.extend(dev.convert(z, _lib.F32), 1)
        n_down = zext.ny - n_up - z.ny
        w = dev.empty(zext.ny, zext.nx, _lib.F32, np.float32)
        nbytes = lib.hd_pdfill_workspace_bytes(zext.ny, zext.nx)
        work = dev.scratch(nbytes)
        flags = (2 if n_up else 0) | (4 if n_down else 0)          # halo rows are not raster frame
        visits = ctypes.c_int(0)
        rounds = 0
        while True:
            _lib.check(lib.hd_pdfill_band(zext.ptr, zext.pitch, w.ptr, w.pitch, zext.ny, zext.nx,
                                          ctypes.c_void_p(work.data_ptr()), nbytes, flags | (1 if rounds else 0),
                                          ctypes.byref(visits), dev.stream_ptr()))
            rounds += 1
            tw = w.tensor()
            # my first / last OWNED rows go to the neighbours' halo rows
            send_up, send_down = tw[n_up:n_up + 1], tw[n_up + z.ny - 1:n_up + z.ny]
            recv_up, recv_down = self.comm.exchange(send_up, send_down, send_up, send_down)
            lowered = False
            if recv_up is not None:
                lowered |= bool((recv_up < tw[0:1]).any().item())
                tw[0:1].copy_(torch.minimum(tw[0:1], recv_up))
            if recv_down is not None:
                lowered |= bool((recv_down < tw[-1:]).any().item())
                tw[-1:].copy_(torch.minimum(tw[-1:], recv_down))
            if not self.comm.any(lowered):
                break
            if rounds >= max_rounds:
                raise dev.DeviceError("banded sink-fill did not converge")
        _lib.check(lib.hd_pdfill_finish(zext.ptr, zext.pitch, w.ptr, w.pitch, zext.ny, zext.nx, dev.stream_ptr()))
        self.fill_rounds = rounds
        d8 = dev.empty(zext.ny, zext.nx, _lib.U8, np.uint8)
        # D8 reads a one-row halo of the converged surface: refresh it once more
        tw = w.tensor()
        recv_up, recv_down = self.comm.exchange(tw[n_up:n_up + 1], tw[n_up + z.ny - 1:n_up + z.ny], tw[0:1], tw[0:1])
        if recv_up is not None:
            tw[0:1].copy_(recv_up)
        if recv_down is not None:
            tw[-1:].copy_(recv_down)
        _lib.check(lib.hd_d8(w.ptr, w.pitch, d8.ptr, d8.pitch, zext.ny, zext.nx, dev.stream_ptr()))
        return w.sub(n_up, n_up + z.ny, 0, z.nx), d8.sub(n_up, n_up + z.ny, 0, z.nx)
