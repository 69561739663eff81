"""Row-band sharding of one large mosaic across the GPUs of a box (BASELINE.json configs[3] / [4]).

A mosaic of ``ny`` rows is cut into ``world`` contiguous row bands (one per rank / GPU; band starts are even so that
transforms which pack two rows keep the same pairs as on one GPU).  Everything here is built so that the sharded
result equals the single-GPU result BIT FOR BIT -- exact-class and tolerance-class stages alike:

* **Stencil stages** run the unchanged single-GPU stage code (``ConditioningChain._stage_*``) on the band EXTENDED by
  ``HALO`` rows of its neighbours.  The halo is exchanged ONCE per input (NCCL send / recv straight into the margin rows
  of a buffer that was allocated with them -- no per-stage copies); a stage with half-window h leaves h more rows at
  the cut edges invalid, and HALO = 24 covers the deepest dependency chain (closing 2 + 3 x quadratic 7 + 1 for the
  final 3x3 mean).  Every kernel's arithmetic for a cell is a fixed sequence of operations on that cell's window
  (csrc/quadratic.cu accumulates in double for exactly this reason), so cut positions cannot change a bit.
* **Fourier stage**: the row / column transforms of ``hd_fft2_forward_shift_abs`` / ``hd_fft2_masked_inverse_abs`` run
  on each rank's local rows through ``hd_fft_band_pass`` (the same row passes, same pairing), the four transposes
  become all-to-all exchanges (grouped NCCL send / recv), the conjugate half of the spectrum is completed locally
  because each rank owns a set of spectrum rows closed under ky -> -ky ("K layout"), and the window-55 peak detector
  works on row slabs of the spectrum quarters with 61 rows of context (|F| stays banded: no all-gather).
* **Sink-fill**: every rank relaxes [halo row | band | halo row] to its local fixed point, edge rows are exchanged, a
  small kernel lowers the halo rows and raises a DEVICE flag, the flags are all-reduced and the host reads one word
  per round.  The fixed point is unique, so the banded surface equals the single-GPU one; NaN restoration and D8 are
  one fused pass (``hd_pdfill_finish_d8``).

The communicator is abstract: ``DistComm`` (torch.distributed: NCCL on the GPU box, gloo in the CPU tests) or
``ThreadComm`` (ranks emulated as threads of one process on one device: the single-GPU parity tests).
"""
import ctypes
import threading

import numpy as np
import torch

from . import _lib, device as dev

HALO = 24               # rows of context every extended band carries (see module docstring)
MASK_HALO = 2 * 27 + 1 + 6     # MaskFourier: two hollow-mean passes (27 each), IsolatedPoints (1), Expand 13 (6)
LOAD_REAL, LOAD_C64, LOAD_MASKED, LOAD_HPAIR = 0, 1, 2, 3


def band_bounds(n, world, align=1):
    """Row range [r0, r1) of every rank; starts are multiples of ``align``; sizes differ by at most ``align``."""
    n, world, align = int(n), int(world), int(align)
    units = -(-n // align)
    base, rem = divmod(units, world)
    out, u = [], 0
    for k in range(world):
        c = base + (1 if k < rem else 0)
        out.append((min(n, u * align), min(n, (u + c) * align)))
        u += c
    return out


# ---- communicators -------------------------------------------------------------------------------------------------
class DistComm:
    """Point-to-point exchange over torch.distributed (NCCL on GPUs, gloo on CPU)."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)

    def p2p(self, sends, recvs):
        """sends: [(contiguous tensor, dst rank)], recvs: [(contiguous tensor, src rank)].  Messages between one pair of
        ranks are matched in posting order.  One grouped batch: NCCL fuses it into a single kernel over NVLink."""
        dist = self.dist
        local = [t for t, d in sends if d == self.rank]
        ops = []
        for t, s in recvs:
            if s == self.rank:
                t.copy_(local.pop(0))
            else:
                ops.append(dist.P2POp(dist.irecv, t, s, self.group))
        for t, d in sends:
            if d != self.rank:
                ops.append(dist.P2POp(dist.isend, t, d, self.group))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()

    def allreduce_max_(self, t):
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX, group=self.group)
        return t


class ThreadComm:
    """Emulates ``world`` ranks as threads of one process (one GPU): the banded algorithms are tested against the
    single-GPU path without several devices.  ``ThreadComm.run(world, fn)`` starts fn(comm) per rank."""

    class _Shared:
        def __init__(self, world):
            self.world = world
            self.barrier = threading.Barrier(world)
            self.box = {}
            self.vals = [None] * world

    def __init__(self, shared, rank):
        self.shared, self.rank, self.world = shared, rank, shared.world

    def p2p(self, sends, recvs):
        sh = self.shared
        if torch.cuda.is_available():
            torch.cuda.synchronize()
        for t, d in sends:
            sh.box.setdefault((self.rank, d), []).append(t.clone())
        sh.barrier.wait()
        for t, s in recvs:
            src = sh.box[(s, self.rank)].pop(0)
            assert tuple(src.shape) == tuple(t.shape), (tuple(src.shape), tuple(t.shape))
            t.copy_(src)
        if torch.cuda.is_available():
            torch.cuda.synchronize()
        sh.barrier.wait()

    def allreduce_max_(self, t):
        sh = self.shared
        if torch.cuda.is_available():
            torch.cuda.synchronize()
        sh.vals[self.rank] = t.clone()
        sh.barrier.wait()
        out = torch.stack(sh.vals).max(dim=0).values
        sh.barrier.wait()
        t.copy_(out)
        return t

    @staticmethod
    def run(world, fn):
        shared = ThreadComm._Shared(world)
        results, errors = [None] * world, []
        device = torch.cuda.current_device() if torch.cuda.is_available() else None

        def work(rank):
            try:
                if device is not None:
                    torch.cuda.set_device(device)
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


# ---- row redistribution between two layouts ----------------------------------------------------------------------------
def _runs(src_ids, dst_ids):
    """Maximal runs that are contiguous in BOTH id lists: [(src_lo, dst_lo, length)], ordered by dst position."""
    src_ids, dst_ids = np.asarray(src_ids, dtype=np.int64), np.asarray(dst_ids, dtype=np.int64)
    if len(src_ids) == 0 or len(dst_ids) == 0:
        return []
    where = np.full(int(max(src_ids.max(), dst_ids.max())) + 2, -1, dtype=np.int64)
    where[src_ids] = np.arange(len(src_ids))
    p = where[dst_ids]                                   # source position of every wanted row, -1 = held elsewhere
    k = np.nonzero(p >= 0)[0]
    if len(k) == 0:
        return []
    brk = np.nonzero((np.diff(k) != 1) | (np.diff(p[k]) != 1))[0] + 1
    starts = np.concatenate([[0], brk])
    ends = np.concatenate([brk, [len(k)]])
    return [(int(p[k[a]]), int(k[a]), int(b - a)) for a, b in zip(starts, ends)]


_PLAN_CACHE = {}


def redistribute_rows(comm, src, src_ids, dst_ids, dst, key=None):
    """Move rows between two layouts of the same global row set.  src / dst: 2-D tensors (local rows x cols, a row is
    contiguous); src_ids[r] / dst_ids[r]: the global row ids rank r holds / wants, in local order (the same lists on
    every rank).  A global row may be wanted by several ranks (halos) but is held by exactly one.  ``key``: caches the
    message plan (the layouts of a mosaic shape do not change from step to step)."""
    me = comm.rank
    plan = _PLAN_CACHE.get((key, me, comm.world)) if key is not None else None
    if plan is None:
        plan = ([(s0, ln, j) for j in range(comm.world) for s0, _, ln in _runs(src_ids[me], dst_ids[j])],
                [(d0, ln, i) for i in range(comm.world) for _, d0, ln in _runs(src_ids[i], dst_ids[me])])
        if key is not None:
            _PLAN_CACHE[(key, me, comm.world)] = plan
    comm.p2p([(src[s0:s0 + ln], j) for s0, ln, j in plan[0]], [(dst[d0:d0 + ln], i) for d0, ln, i in plan[1]])
    return dst


def _cview(t):
    """complex64 tensor -> float32 view with a trailing dimension of 2 (what the communicators move)."""
    return torch.view_as_real(t) if t.is_complex() else t


# ---- one rank's view of the mosaic --------------------------------------------------------------------------------------
class ExtRaster:
    """A band with its halo margins: ``raster`` has up + rows + down rows; rows [up, up + rows) are owned."""

    def __init__(self, raster, up, rows, down):
        self.raster, self.up, self.rows, self.down = raster, up, rows, down

    def owned(self):
        r = self.raster
        return r.sub(self.up, self.up + self.rows, 0, r.nx)

    def with_halo(self, h):
        """View of the owned rows plus h halo rows (fewer at the ends of the mosaic)."""
        r = self.raster
        u, d = min(h, self.up), min(h, self.down)
        return r.sub(self.up - u, self.up + self.rows + d, 0, r.nx), u, d


class Band:
    """One rank's view of the mosaic."""

    def __init__(self, comm, ny, nx, halo=HALO):
        self.comm, self.ny, self.nx, self.halo = comm, int(ny), int(nx), int(halo)
        self.bounds = band_bounds(ny, comm.world, 2)
        self.r0, self.r1 = self.bounds[comm.rank]
        if comm.world > 1 and min(b - a for a, b in self.bounds) < self.halo:
            # the same test on every rank (it only depends on ny and world): nobody enters a collective alone
            raise ValueError(f"{ny} rows over {comm.world} ranks leave bands thinner than the {self.halo}-row halo")
        self.up = self.halo if comm.rank > 0 else 0
        self.down = self.halo if comm.rank < comm.world - 1 else 0
        self.fill_rounds = None

    @property
    def rows(self):
        return self.r1 - self.r0

    def take(self, mosaic):
        """This rank's rows of a host mosaic (tests / loaders)."""
        return np.ascontiguousarray(mosaic[self.r0:self.r1])

    # -- halo plumbing ------------------------------------------------------------------------------------------------
    def alloc_ext(self, dtype, ref_dtype=None):
        """Uninitialised extended raster [halo | band | halo] (no halo at the ends of the mosaic)."""
        r = dev.empty(self.up + self.rows + self.down, self.nx, dtype, ref_dtype)
        return ExtRaster(r, self.up, self.rows, self.down)

    def exchange_halo(self, ext, h=None):
        """Fill the margin rows of ``ext`` with the neighbours' edge rows: sends / receives whole pitched rows straight
        from / into the buffer (a row range of a pitched raster is contiguous)."""
        h = self.halo if h is None else int(h)
        buf, up, rows = ext.raster.buf, ext.up, ext.rows
        y0 = ext.raster._y0
        sends, recvs = [], []
        if ext.up:
            sends.append((buf[y0 + up:y0 + up + h], self.comm.rank - 1))
            recvs.append((buf[y0 + up - h:y0 + up], self.comm.rank - 1))
        if ext.down:
            sends.append((buf[y0 + up + rows - h:y0 + up + rows], self.comm.rank + 1))
            recvs.append((buf[y0 + up + rows:y0 + up + rows + h], self.comm.rank + 1))
        self.comm.p2p([(_cview(t), d) for t, d in sends], [(_cview(t), s) for t, s in recvs])
        return ext

    def extended(self, raster):
        """Extended copy of a band raster (inputs: done once per input, not per stage) with the halo exchanged."""
        ext = self.alloc_ext(raster.dtype, raster.ref_dtype)
        ext.owned().tensor().copy_(raster.tensor())
        return self.exchange_halo(ext)

    def apply(self, filt, raster, h):
        """Run a single-GPU filter (anything with ``run_device``) on the band with an h-row halo; returns the band's
        rows of the result."""
        ext = self.extended(raster)
        view, u, _ = ext.with_halo(h)
        out = filt.run_device(view)
        return out.sub(u, u + self.rows, 0, raster.nx)

    # -- the Fourier stage ------------------------------------------------------------------------------------------------
    def _dense(self, rows, cols, dtype):
        tdt = {_lib.C64: torch.complex64, _lib.F32: torch.float32, _lib.U8: torch.uint8}[dtype]
        return dev.DeviceRaster(torch.empty((max(int(rows), 1), int(cols)), dtype=tdt, device=dev.device()), rows, cols, dtype)

    def _pass(self, plan, axis, load, src, nrows, out_t, keep_cols=0, mask=None, shift_cols=0, inverse=0, real_out=0):
        lib = _lib.load()
        n = self.nx if axis == 0 else self.ny
        nbytes = int(nrows) * n * 8
        work = dev.scratch(nbytes)
        mp, mpitch = (mask.ptr, mask.pitch) if mask is not None else (None, 0)
        _lib.check(lib.hd_fft_band_pass(plan, axis, load, src.ptr, src.pitch, int(nrows), mp, mpitch, int(shift_cols),
                                        int(inverse), int(real_out), out_t.ptr, out_t.pitch, int(keep_cols),
                                        ctypes.c_void_p(work.data_ptr()), nbytes, dev.stream_ptr()))

    def _transpose_exchange(self, t, send_bounds, recv_cols, out):
        """t: (N x my_len) local transposed block.  Rows [a, b) of it go to the rank that owns [a, b) (send_bounds);
        what arrives from rank i is an (mine x len_i) block that lands in columns recv_cols[i] of ``out``
        (recv_cols[i] = list of (col0, col1) ranges, in the order of rank i's local rows)."""
        comm = self.comm
        tt = t.tensor()
        mine = out.ny
        sends = [(_cview(tt[a:b]), j) for j, (a, b) in enumerate(send_bounds) if b > a and tt.shape[1] > 0]
        bufs, recvs = [], []
        for i in range(comm.world):
            width = sum(c1 - c0 for c0, c1 in recv_cols[i])
            if width == 0 or mine == 0:
                bufs.append(None)
                continue
            b = torch.empty((mine, width), dtype=tt.dtype, device=tt.device)
            bufs.append(b)
            recvs.append((_cview(b), i))
        comm.p2p(sends, recvs)
        to = out.tensor()
        for i, b in enumerate(bufs):
            if b is None:
                continue
            k = 0
            for c0, c1 in recv_cols[i]:
                to[:, c0:c1].copy_(b[:, k:k + (c1 - c0)])
                k += c1 - c0
        return out

    def _k_layouts(self):
        """Per rank: (a, b, rows, [ky of every local row]) -- spectrum rows ky in [a, b) plus their mirrors."""
        lib = _lib.load()
        nyh = self.ny // 2 + 1
        out = []
        for a, b in band_bounds(nyh, self.comm.world, 1):
            n = int(lib.hd_klayout_rows(a, b, self.ny))
            nlo = b - a
            nhi = n - nlo
            hlo = int(lib.hd_klayout_ky(a, b, self.ny, nlo)) if nhi else 0
            out.append(dict(a=a, b=b, rows=n, nlo=nlo, nhi=nhi, hlo=hlo,
                            ky=np.concatenate([np.arange(a, b), np.arange(hlo, hlo + nhi)]).astype(np.int64)))
        return out

    def detect_apply_fourier(self, srtm, out_ext=None):
        """DetectApplyFourier (custom_filters.py:1053-1101) on a banded mosaic.  ``srtm``: F32 device raster of this
        rank's rows.  Returns an ExtRaster (F32 storage, float64 reference dtype) whose owned rows hold the stripe-free
        DEM -- bit-identical to ``DetectApplyFourier().run_device`` on the whole mosaic; its halo rows are NOT filled."""
        from .filters import custom_filters as cf
        lib = _lib.load()
        comm, W, me = self.comm, self.comm.world, self.comm.rank
        ny, nx = self.ny, self.nx
        nh = nx // 2 + 1
        odd = bool((ny & 1) and (nx & 1))
        RB, CB, XB = self.bounds, band_bounds(nh, W, 1), band_bounds(nx, W, 2)
        KL = self._k_layouts()
        kl = KL[me]
        plan = dev.fft_plan(ny, nx)
        rows = self.rows
        src = dev.convert(srtm, _lib.F32)

        # forward, along x: two real rows per transform, half spectrum kept; transposed block (nh x rows)
        t1 = self._dense(nh, rows, _lib.C64)
        self._pass(plan, 0, LOAD_REAL, src, rows, t1, keep_cols=nh)
        c0, c1 = CB[me]
        at = self._dense(c1 - c0, ny, _lib.C64)                     # my columns of the half spectrum, all y
        self._transpose_exchange(t1, CB, [[RB[i]] for i in range(W)], at)
        del t1
        # forward, along y; transposed block (ny x my columns), natural ky order
        t2 = self._dense(ny, c1 - c0, _lib.C64)
        self._pass(plan, 1, LOAD_C64, at, c1 - c0, t2)
        del at
        # to the K layout: rank i gets rows ky in [a_i, b_i) and their mirrors
        half = self._dense(kl["rows"], nh, _lib.C64)
        tt, ht = t2.tensor(), half.tensor()
        sends, recvs, bufs = [], [], []
        for i, k in enumerate(KL):
            if c1 > c0:
                sends.append((_cview(tt[k["a"]:k["b"]].contiguous()), i))
                if k["nhi"]:
                    sends.append((_cview(tt[k["hlo"]:k["hlo"] + k["nhi"]].contiguous()), i))
        for j, (d0, d1) in enumerate(CB):
            if d1 > d0:
                lo = torch.empty((kl["nlo"], d1 - d0), dtype=torch.complex64, device=ht.device)
                recvs.append((_cview(lo), j))
                hi = None
                if kl["nhi"]:
                    hi = torch.empty((kl["nhi"], d1 - d0), dtype=torch.complex64, device=ht.device)
                    recvs.append((_cview(hi), j))
                bufs.append((d0, d1, lo, hi))
        comm.p2p(sends, recvs)
        for d0, d1, lo, hi in bufs:
            ht[:kl["nlo"], d0:d1].copy_(lo)
            if hi is not None:
                ht[kl["nlo"]:, d0:d1].copy_(hi)
        del t2, bufs
        # conjugate half, column shift, |F|  (FourierInitial, custom_filters.py:859-877)
        nk = kl["rows"]
        fshift = self._dense(nk, nx, _lib.C64)
        fabs = self._dense(nk, nx, _lib.F32)
        _lib.check(lib.hd_hermitian_complete(half.ptr, half.pitch, fshift.ptr, fshift.pitch, fabs.ptr, fabs.pitch, kl["a"],
                                             kl["b"], ny, nx, dev.stream_ptr()))
        del half

        # ---- peak detector on row slabs of the two upper quarters (FourierProcessQuarters, :880-1050) -------------
        my, y_odd, mx, x_odd = ny // 2, ny & 1, nx // 2, nx & 1
        m = cf.FOURIER_MARGIN
        qh, qw = my - m, mx - m
        cf.check_window((qh, qw), cf.BLANKS_WINDOW)
        shifted = [(k["ky"] + ny // 2) % ny for k in KL]            # shifted row id of every local K row, per rank
        SL = band_bounds(qh, W, 1)
        if min(b - a for a, b in SL) < 16:                           # tiny quarters: one rank runs the detector
            SL = [(0, qh)] + [(qh, qh)] * (W - 1)
        ctx = [(max(0, a - MASK_HALO), min(qh, b + MASK_HALO)) if b > a else (0, 0) for a, b in SL]
        slab_ids = [np.arange(lo, hi) for lo, hi in ctx]
        s0, s1 = SL[me]
        lo, hi = ctx[me]
        masks = []
        if hi > lo:
            slab = self._dense(hi - lo, nx, _lib.F32)
        else:
            slab = self._dense(0, nx, _lib.F32)
        redistribute_rows(comm, fabs.tensor(), shifted, slab_ids, slab.tensor()[:hi - lo], key=("fabs", ny, nx))
        x0 = mx + m + x_odd
        for (xa, xb) in ((0, qw), (x0, nx)):
            own = self._dense(s1 - s0, qw, _lib.U8)
            if hi > lo:
                q = dev.empty(hi - lo, qw, _lib.F32)
                dev.elementwise(_lib.OP_COPY, slab.sub(0, hi - lo, xa, xb), None, 0.0, q)
                mk = dev.convert(cf.MaskFourier().run_device(q), _lib.U8)
                own.tensor().copy_(mk.tensor()[s0 - lo:s1 - lo])
            masks.append(own)
        del slab
        # the quarter-mask rows each rank's K rows look at (point mirror for the lower half, :1002-1027)
        need = []
        for sh in shifted:
            top = sh < my
            by = np.where(top, sh, sh - my - y_odd)
            qy = np.where(top, by, my - 1 - by)
            ok = (by >= 0) & (qy >= 0) & (qy < qh)
            need.append((int(qy[ok].min()), int(qy[ok].max()) + 1) if ok.any() else (0, 0))
        own_ids = [np.arange(a, b) for a, b in SL]
        need_ids = [np.arange(a, b) for a, b in need]
        q0, q1 = need[me]
        loc = []
        for own in masks:
            mloc = self._dense(q1 - q0, qw, _lib.U8)
            redistribute_rows(comm, own.tensor()[:s1 - s0], own_ids, need_ids, mloc.tensor()[:q1 - q0], key=("qmask", ny, nx))
            loc.append(mloc)
        mask = self._dense(nk, nx, _lib.U8)
        _lib.check(lib.hd_fourier_mask_assemble_rows(loc[0].ptr, loc[0].pitch, loc[1].ptr, loc[1].pitch, q0, q1 - q0,
                                                     mask.ptr, mask.pitch, nk, kl["a"], kl["b"], ny, nx, m,
                                                     dev.stream_ptr()))
        self._last_mask = (mask, kl)

        # ---- inverse, along x: (1 - mask) * F on this rank's spectrum rows; the Hermitian (odd x odd) path needs only
        # ky <= ny/2.  Transposed block (nx x rows) ---------------------------------------------------------------
        n_i1 = kl["nlo"] if odd else nk
        t3 = self._dense(nx, n_i1, _lib.C64)
        self._pass(plan, 0, LOAD_MASKED, fshift, n_i1, t3, mask=mask, shift_cols=nx // 2, inverse=1)
        del fshift
        xa, xb = XB[me]
        bt = self._dense(xb - xa, ny, _lib.C64)
        cols = [[(k["a"], k["b"])] + ([(k["hlo"], k["hlo"] + k["nhi"])] if (not odd and k["nhi"]) else []) for k in KL]
        self._transpose_exchange(t3, XB, cols, bt)
        del t3
        if odd:
            _lib.check(lib.hd_conj_mirror_fill(bt.ptr, bt.pitch, xb - xa, ny, dev.stream_ptr()))
        # inverse, along y: real output (two columns per transform on the Hermitian path); transposed block (ny x my x)
        t4 = self._dense(ny, xb - xa, _lib.F32)
        self._pass(plan, 1, LOAD_HPAIR if odd else LOAD_C64, bt, xb - xa, t4, inverse=1, real_out=1)
        del bt
        out_ext = out_ext or self.alloc_ext(_lib.F32, np.float64)
        self._transpose_exchange(t4, RB, [[XB[i]] for i in range(W)], out_ext.owned())
        return out_ext

    # -- the whole chain on row bands (BASELINE.json configs[4]: one mosaic over the GPUs of a box) --------------------
    def conditioning_chain(self, srtm, groves_class, hsheds, groves_iterations=3, with_hydrology=True):
        """HydroDEMProcess.start (hydro_dem_process.py:122-153) on this rank's rows of a mosaic.  Inputs: device rasters
        of the band's rows (F32, U8 0/1, F32), or ExtRaster objects whose owned rows are already in place (groves,
        hsheds: saves the input copy).  Returns {"final", "dem_complete", "filled", "d8"} for the band's rows -- every
        one of them bit-identical to the single-GPU ConditioningChain on the whole mosaic."""
        from .pipeline import ConditioningChain
        chain = ConditioningChain(groves_iterations=groves_iterations, with_hydrology=False)
        g_ext = groves_class if isinstance(groves_class, ExtRaster) else self.extended(dev.convert(groves_class, _lib.U8))
        h_ext = hsheds if isinstance(hsheds, ExtRaster) else self.extended(hsheds)
        if isinstance(g_ext, ExtRaster) and groves_class is g_ext:
            self.exchange_halo(g_ext)
        if isinstance(h_ext, ExtRaster) and hsheds is h_ext:
            self.exchange_halo(h_ext)
        dem = self.detect_apply_fourier(srtm)                                        # image_srtm.py:125-126
        self.exchange_halo(dem)
        st = {"fourier": dem.raster}
        chain._stage_groves(st, g_ext.raster)                                        # image_srtm.py:177-199
        chain._stage_combine(st, h_ext.raster, None)                                 # LagoonsDetection ... :149
        own = lambda r: r.sub(self.up, self.up + self.rows, 0, self.nx)              # noqa: E731
        out = {"final": own(st["final"]), "dem_complete": own(st["dem_complete"])}
        if with_hydrology:
            out["filled"], out["d8"] = self.sinkfill(ExtRaster(st["final32"], self.up, self.rows, self.down))
        return out

    # -- sink-fill + D8 -----------------------------------------------------------------------------------------------------
    def sinkfill(self, z, max_rounds=10000):
        """Banded Planchon-Darboux fixed point + D8 (bit-identical to the single-GPU result).  ``z``: F32 device raster
        of the band's rows, or an ExtRaster whose owned rows hold them (its margin rows are overwritten)."""
        lib = _lib.load()
        comm = self.comm
        if not isinstance(z, ExtRaster):
            ext = self.alloc_ext(_lib.F32, np.float32)
            ext.owned().tensor().copy_(dev.convert(z, _lib.F32).tensor())
            z = ext
        self.exchange_halo(z, 1)
        zv, n_up, n_down = z.with_halo(1)
        rows, nx = self.rows, self.nx
        w = dev.empty(zv.ny, nx, _lib.F32, np.float32)
        d8 = dev.empty(zv.ny, nx, _lib.U8, np.uint8)
        nbytes = lib.hd_pdfill_workspace_bytes(zv.ny, nx)
        work = dev.scratch(nbytes)
        wp = ctypes.c_void_p(work.data_ptr())
        flags = (2 if n_up else 0) | (4 if n_down else 0)          # halo rows are not raster frame
        lowered = torch.zeros(1, dtype=torch.int32, device=dev.device())
        tw = w.tensor()
        recv_up = torch.empty(nx, dtype=torch.float32, device=dev.device()) if n_up else None
        recv_down = torch.empty(nx, dtype=torch.float32, device=dev.device()) if n_down else None
        rounds = 0
        while True:
            _lib.check(lib.hd_pdfill_band(zv.ptr, zv.pitch, w.ptr, w.pitch, zv.ny, nx, wp, nbytes,
                                          flags | (1 if rounds else 0), None, dev.stream_ptr()))
            rounds += 1
            if comm.world == 1:
                break
            lowered.zero_()
            sends, recvs = [], []
            if n_up:                                               # my first / last OWNED rows go to the neighbours' halo rows
                sends.append((tw[n_up, :nx], comm.rank - 1))
                recvs.append((recv_up, comm.rank - 1))
            if n_down:
                sends.append((tw[n_up + rows - 1, :nx], comm.rank + 1))
                recvs.append((recv_down, comm.rank + 1))
            comm.p2p(sends, recvs)
            lp = ctypes.c_void_p(lowered.data_ptr())
            if n_up:
                _lib.check(lib.hd_halo_min_flag(w.ptr, ctypes.c_void_p(recv_up.data_ptr()), nx, lp, dev.stream_ptr()))
            if n_down:
                last = w.sub(zv.ny - 1, zv.ny, 0, nx)
                _lib.check(lib.hd_halo_min_flag(last.ptr, ctypes.c_void_p(recv_down.data_ptr()), nx, lp, dev.stream_ptr()))
            comm.allreduce_max_(lowered)
            if int(lowered.item()) == 0:                           # the one host read of the round
                break
            if rounds >= max_rounds:
                raise dev.DeviceError("banded sink-fill did not converge")
        self.fill_rounds = rounds
        # at the fixed point a halo row equals the neighbour's owned row, so D8 needs no further exchange; NaN
        # restoration and the flow directions are one fused pass
        _lib.check(lib.hd_pdfill_finish_d8(w.ptr, w.pitch, d8.ptr, d8.pitch, zv.ny, nx, wp, dev.stream_ptr()))
        self._fill_work = work
        return w.sub(n_up, n_up + rows, 0, nx), d8.sub(n_up, n_up + rows, 0, nx)

    def fill_status(self):
        """Sticky status word of the last sinkfill on this rank (0 = fixed point reached)."""
        st = ctypes.c_int(-1)
        _lib.check(_lib.load().hd_pdfill_status(ctypes.c_void_p(self._fill_work.data_ptr()), ctypes.byref(st),
                                                dev.stream_ptr()))
        return st.value

    # -- host API: one rank's rows in, one rank's rows out --------------------------------------------------------------
    def apply_to_host(self, srtm_rows, groves_rows, hsheds_rows):
        """The banded chain through host buffers: this rank's rows of the three input mosaics (ndarrays: float32,
        uint8 / bool, float32; pinned memory is copied asynchronously) -> {"final" float64, "filled" float32, "d8" uint8}
        ndarrays of the same rows.  What bench.py's e2e measures at N > 1."""
        import torch
        for a in (srtm_rows, groves_rows, hsheds_rows):
            if not isinstance(a, np.ndarray):
                from .exceptions import NumpyArrayExpectedError
                raise NumpyArrayExpectedError(a)
            if a.shape != (self.rows, self.nx):
                raise ValueError(f"expected this rank's rows {(self.rows, self.nx)}, got {a.shape}")
        cur = torch.cuda.current_stream()
        srtm = dev.upload(np.ascontiguousarray(srtm_rows, dtype=np.float32))
        g_ext, h_ext = self.alloc_ext(_lib.U8, np.uint8), self.alloc_ext(_lib.F32, np.float32)
        g = np.ascontiguousarray(groves_rows)
        dev.upload_into(g_ext.owned(), g.view(np.uint8) if g.dtype == np.bool_ else g.astype(np.uint8, copy=False), cur)
        dev.upload_into(h_ext.owned(), np.ascontiguousarray(hsheds_rows, dtype=np.float32), cur)
        out = self.conditioning_chain(srtm, g_ext, h_ext)
        return {k: dev.download(out[k]) for k in ("final", "filled", "d8")}
