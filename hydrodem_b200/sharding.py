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


class _Trace:
    """HD_BAND_TRACE=1: per-phase host and device times of the banded chain (tools/band_phases.py prints them)."""

    def __init__(self):
        import os
        self.on = bool(os.environ.get("HD_BAND_TRACE"))
        self.marks = []

    def mark(self, name):
        if self.on:
            import time
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            self.marks.append((name, time.perf_counter(), ev))

    def report(self):
        torch.cuda.synchronize()
        out = []
        for (n0, t0, e0), (n1, t1, e1) in zip(self.marks, self.marks[1:]):
            out.append((n1, (t1 - t0) * 1e3, e0.elapsed_time(e1)))
        self.marks = []
        return out


TRACE = _Trace()


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

    class _Handle:
        def __init__(self, works):
            self.works = works

        def wait(self):
            for w in self.works:
                w.wait()                        # NCCL: the current stream waits for the transfer, the host does not

    def p2p_async(self, sends, recvs):
        """sends: [(contiguous tensor, dst rank)], recvs: [(contiguous tensor, src rank)].  Messages between one pair of
        ranks are matched in posting order.  One grouped batch: NCCL fuses it into a single kernel over NVLink, on its
        own stream -- kernels launched before ``wait()`` overlap the transfer."""
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
        return self._Handle(dist.batch_isend_irecv(ops) if ops else [])

    def p2p(self, sends, recvs):
        self.p2p_async(sends, recvs).wait()

    def allreduce_max_(self, t):
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX, group=self.group)
        return t

    same_process = False

    def gather_objects(self, obj):
        out = [None] * self.world
        self.dist.all_gather_object(out, obj, group=self.group)
        return out

    def device_barrier(self):
        """Stream-ordered barrier: kernels launched after it start only when every rank's earlier kernels are done."""
        if not hasattr(self, "_bar"):
            self._bar = torch.zeros(1, dtype=torch.int32,
                                    device="cuda" if self.dist.get_backend(self.group) == "nccl" else "cpu")
        self.dist.all_reduce(self._bar, group=self.group)


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

    class _Done:
        def wait(self):
            pass

    same_process = True

    def gather_objects(self, obj):
        sh = self.shared
        sh.vals[self.rank] = obj
        sh.barrier.wait()
        out = list(sh.vals)
        sh.barrier.wait()
        return out

    def device_barrier(self):
        if torch.cuda.is_available():
            torch.cuda.synchronize()
        self.shared.barrier.wait()

    def p2p_async(self, sends, recvs):
        self.p2p(sends, recvs)
        return self._Done()

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


# ---- peer memory: persistent exchange buffers every rank can store into ---------------------------------------------------
class PeerBuffers:
    """Named device buffers of this rank, living in ONE allocation that the other ranks have mapped (CUDA IPC; ranks that
    are threads of one process simply share the pointers).  ``ptr(rank, name)`` is valid in this process's kernels."""

    def __init__(self, comm, sizes):
        lib = _lib.load()
        self.comm = comm
        self.offsets, total = {}, 0
        for name, nbytes in sizes.items():
            self.offsets[name] = total
            total += (int(nbytes) + 255) // 256 * 256
        self.arena = torch.empty(max(total, 256), dtype=torch.uint8, device=dev.device())
        base = self.arena.data_ptr()
        self.ok = True
        if comm.same_process:
            infos = comm.gather_objects((base, self.offsets))
            self.bases = [b for b, _ in infos]
        else:
            # every rank runs the same sequence of collectives whatever fails locally; `ok` is agreed on at the end
            handle = ctypes.create_string_buffer(64)
            off = ctypes.c_int64(0)
            mine = None
            if lib.hd_ipc_export(ctypes.c_void_p(base), handle, ctypes.byref(off)) == _lib.HD_OK:
                mine = (handle.raw, off.value, self.offsets)
            infos = comm.gather_objects(mine if mine is not None else (None, 0, self.offsets))
            self.bases, self._opened = [], []
            good = mine is not None and all(i[0] is not None for i in infos)
            for j, (raw, off_j, _) in enumerate(infos):
                if j == comm.rank or not good:
                    self.bases.append(base)
                    continue
                p = ctypes.c_void_p()
                if lib.hd_ipc_import(ctypes.create_string_buffer(raw, 64), ctypes.byref(p)) != _lib.HD_OK:
                    good = False
                    self.bases.append(base)
                    continue
                self._opened.append(p)
                self.bases.append(p.value + off_j)
            self.ok = all(comm.gather_objects(bool(good)))
        self.peer_offsets = [info[-1] for info in infos]

    def ptr(self, rank, name):
        return self.bases[rank] + self.peer_offsets[rank][name]

    def view(self, name, rows, cols, dtype):
        """This rank's buffer ``name`` as a dense DeviceRaster."""
        tdt = {_lib.C64: torch.complex64, _lib.F32: torch.float32, _lib.U8: torch.uint8}[dtype]
        n = int(rows) * int(cols) * torch.empty(0, dtype=tdt).element_size()
        o = self.offsets[name]
        t = self.arena[o:o + max(n, 0)].view(tdt).view(int(rows), int(cols)) if rows and cols else \
            torch.empty((1, int(cols)), dtype=tdt, device=self.arena.device)
        return dev.DeviceRaster(t, rows, cols, dtype)

    def close(self):
        lib = _lib.load()
        for p in getattr(self, "_opened", []):
            lib.hd_ipc_close(p)
        self._opened = []


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
        # band starts are multiples of 8: even (two rows share one transform on the device) and aligned with the 8x8 blocks
        # of the sink-fill's coarse level (a block never straddles a cut)
        self.bounds = band_bounds(ny, comm.world, 8)
        self.r0, self.r1 = self.bounds[comm.rank]
        if comm.world > 1 and min(b - a for a, b in self.bounds) < self.halo:
            # the same test on every rank (it only depends on ny and world): nobody enters a collective alone
            raise ValueError(f"{ny} rows over {comm.world} ranks leave bands thinner than the {self.halo}-row halo")
        import os
        self.fill_overlap = max(8, int(os.environ.get("HD_FILL_OVERLAP", "64")) // 8 * 8)
        self.narrow_final = os.environ.get("HD_NARROW_FINAL", "1") != "0"   # apply_to_host: final DEM down as int16
        if comm.world > 1 and min(b - a for a, b in self.bounds) < 2 * self.fill_overlap:
            self.fill_overlap = max(8, min(b - a for a, b in self.bounds) // 16 * 8)
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

    def _peer_setup(self, KL, CB, XB):
        """Persistent exchange buffers of the Fourier stage in peer-mapped memory (once per Band).  None: the ranks could
        not map each other's memory (or HD_BAND_DIRECT=0) -- the exchanges then go through NCCL send / recv."""
        if not hasattr(self, "_peers"):
            import os
            self._peers = None
            if self.comm.world > 1 and os.environ.get("HD_BAND_DIRECT", "1") != "0":
                lib = _lib.load()
                me = self.comm.rank
                kl = KL[me]
                pitch = int(lib.hd_pitch_elems(self.nx, _lib.F32))
                sizes = {"at": (CB[me][1] - CB[me][0]) * self.ny * 8, "half": kl["rows"] * (self.nx // 2 + 1) * 8,
                         "bt": (XB[me][1] - XB[me][0]) * self.ny * 8, "dem": (self.up + self.rows + self.down) * pitch * 4}
                peers = PeerBuffers(self.comm, sizes)
                if peers.ok:
                    peers.dem_pitch = self.comm.gather_objects((pitch, self.up))
                    self._peers = peers
        return self._peers

    def _pass_scatter(self, plan, axis, load, src, nrows, segs, colsegs, keep_cols=0, mask=None, shift_cols=0, inverse=0,
                      real_out=0):
        """Row pass + transpose whose stores go straight into the owning ranks' buffers (hd_fft_band_pass_scatter).
        segs: [(row0, row1, device pointer, pitch in elements, dst_row0)]; colsegs: [(local row0, column0, length)]."""
        lib = _lib.load()
        n = self.nx if axis == 0 else self.ny
        sc = _lib.Scatter()
        segs = [g for g in segs if g[1] > g[0]]
        sc.nseg, sc.ncolseg = len(segs), len(colsegs)
        for k, (a, b, ptr, pitch, d0) in enumerate(segs):
            sc.seg[k].row0, sc.seg[k].row1, sc.seg[k].base, sc.seg[k].pitch, sc.seg[k].dst_row0 = a, b, ptr, pitch, d0
        for k, (l0, c0, ln) in enumerate(colsegs):
            sc.col_local0[k], sc.col_dst0[k], sc.col_len[k] = l0, c0, ln
        if nrows < 1:
            return
        nbytes = int(nrows) * n * 8
        work = dev.scratch(nbytes)
        mp, mpitch = (mask.ptr, mask.pitch) if mask is not None else (None, 0)
        _lib.check(lib.hd_fft_band_pass_scatter(plan, axis, load, src.ptr, src.pitch, int(nrows), mp, mpitch, int(shift_cols),
                                                int(inverse), int(real_out), ctypes.byref(sc), int(keep_cols),
                                                ctypes.c_void_p(work.data_ptr()), nbytes, dev.stream_ptr()))

    def _pass_exchange(self, plan, axis, load, src, counts, n_out, ranges, segs, out, dtype, mask=None, **kw):
        """(Fallback when the ranks cannot map each other's memory.)  A row pass of the sharded Fourier stage followed by its all-to-all, software pipelined: the local rows are
        transformed in a few chunks, the transposed block of chunk c (n_out x chunk rows) is on its way (NCCL's stream)
        while chunk c + 1 is being transformed, and the blocks that arrive are put in place one chunk behind.
          counts[i]  rows rank i transforms in this pass          ranges[j]  row ranges of the transposed block rank j gets
          segs[i]    [(local row0, output column0, length)]: where rank i's local rows land along the columns of ``out``
        ``out`` rows = the concatenation of ranges[me]."""
        comm, W, me = self.comm, self.comm.world, self.comm.rank
        nch = 1          # (chunked pipelining did not overlap: NCCL's send / recv kernels wait behind the persistent grids)
        chunks = [band_bounds(c, nch, 2) for c in counts]
        tdt = {_lib.C64: torch.complex64, _lib.F32: torch.float32}[dtype]
        pending = []
        to = out.tensor()

        def place(item):
            handle, places, _keep = item
            handle.wait()
            for buf, off, i, ia, ib in places:
                for l0, c0, ln in segs[i]:
                    lo, hi = max(ia, l0), min(ib, l0 + ln)
                    if hi > lo:
                        to[off:off + buf.shape[0], c0 + (lo - l0):c0 + (hi - l0)].copy_(buf[:, lo - ia:hi - ia])

        for c in range(nch):
            ra, rb = chunks[me][c]
            t, sends = None, []
            if rb > ra:
                t = self._dense(n_out, rb - ra, dtype)
                self._pass(plan, axis, load, src.sub(ra, rb, 0, src.nx), rb - ra, t,
                           mask=mask.sub(ra, rb, 0, mask.nx) if mask is not None else None, **kw)
                tt = t.tensor()
                for j in range(W):
                    sends += [(_cview(tt[a:b]), j) for a, b in ranges[j] if b > a]
            recvs, places = [], []
            for i in range(W):
                ia, ib = chunks[i][c]
                off = 0
                for a, b in ranges[me]:
                    if b > a and ib > ia:
                        buf = torch.empty((b - a, ib - ia), dtype=tdt, device=to.device)
                        recvs.append((_cview(buf), i))
                        places.append((buf, off, i, ia, ib))
                    off += b - a
            pending.append((comm.p2p_async(sends, recvs), places, t))
            if len(pending) > 1:
                place(pending.pop(0))
        while pending:
            place(pending.pop(0))
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

        TRACE.mark("daf:start")
        counts_rows = [b - a for a, b in RB]
        counts_cb = [b - a for a, b in CB]
        counts_xb = [b - a for a, b in XB]
        c0, c1 = CB[me]
        # forward along x (two real rows per transform, half spectrum kept) + all-to-all: my columns of the half
        # spectrum, all y
        peers = self._peer_setup(KL, CB, XB)
        if peers is not None:
            # the transposes store straight into the owning rank's buffers over NVLink (no NCCL, no unpack copies);
            # device-side barriers say "destinations free" / "all stores landed"
            comm.device_barrier()
            at = peers.view("at", c1 - c0, ny, _lib.C64)
            self._pass_scatter(plan, 0, LOAD_REAL, src, rows, [(CB[j][0], CB[j][1], peers.ptr(j, "at"), ny, 0) for j in range(W)],
                               [(0, RB[me][0], rows)], keep_cols=nh)
            comm.device_barrier()
            TRACE.mark("F1 rows + T1")
            half = peers.view("half", kl["rows"], nh, _lib.C64)
            segs2 = []
            for j, k in enumerate(KL):
                segs2.append((k["a"], k["b"], peers.ptr(j, "half"), nh, 0))
                segs2.append((k["hlo"], k["hlo"] + k["nhi"], peers.ptr(j, "half"), nh, k["nlo"]))
            self._pass_scatter(plan, 1, LOAD_C64, at, c1 - c0, segs2, [(0, c0, c1 - c0)])
            comm.device_barrier()
        else:
            at = self._dense(c1 - c0, ny, _lib.C64)
            self._pass_exchange(plan, 0, LOAD_REAL, src, counts_rows, nh, [[cb] for cb in CB],
                                [[(0, RB[i][0], counts_rows[i])] for i in range(W)], at, _lib.C64, keep_cols=nh)
            TRACE.mark("F1 rows + T1")
            # forward along y + exchange into the K layout: rank i gets rows ky in [a_i, b_i) and their mirrors
            half = self._dense(kl["rows"], nh, _lib.C64)
            self._pass_exchange(plan, 1, LOAD_C64, at, counts_cb, ny,
                                [[(k["a"], k["b"]), (k["hlo"], k["hlo"] + k["nhi"])] for k in KL],
                                [[(0, CB[i][0], counts_cb[i])] for i in range(W)], half, _lib.C64)
        del at
        TRACE.mark("F2 cols + T2")
        # conjugate half, column shift, |F|  (FourierInitial, custom_filters.py:859-877)
        nk = kl["rows"]
        fshift = self._dense(nk, nx, _lib.C64)
        fabs = self._dense(nk, nx, _lib.F32)
        _lib.check(lib.hd_hermitian_complete(half.ptr, half.pitch, fshift.ptr, fshift.pitch, fabs.ptr, fabs.pitch, kl["a"],
                                             kl["b"], ny, nx, dev.stream_ptr()))
        del half
        TRACE.mark("hermitian complete")

        # ---- peak detector on row slabs of the two upper quarters (FourierProcessQuarters, :880-1050) -------------
        my, y_odd, mx, x_odd = ny // 2, ny & 1, nx // 2, nx & 1
        m = cf.FOURIER_MARGIN
        qh, qw = my - m, mx - m
        cf.check_window((qh, qw), cf.BLANKS_WINDOW)
        shifted = [(k["ky"] + ny // 2) % ny for k in KL]            # shifted row id of every local K row, per rank
        # a rank's slab = the quarter rows it already owns (its mirror rows are a contiguous range of the upper half):
        # only the 61 rows of context on either side have to travel
        SL = []
        for sh in shifted:
            q = np.sort(sh[sh < qh])
            SL.append((int(q[0]), int(q[-1]) + 1) if len(q) and q[-1] - q[0] + 1 == len(q) else None)
        if any(x is None for x in SL) or sorted(SL) != sorted(set(SL)) or sum(b - a for a, b in SL) != qh \
                or min(b - a for a, b in SL) < 16:
            SL = band_bounds(qh, W, 1)
            if min(b - a for a, b in SL) < 16:                       # tiny quarters: one rank runs the detector
                SL = [(0, qh)] + [(qh, qh)] * (W - 1)
        ctx = [(max(0, a - MASK_HALO), min(qh, b + MASK_HALO)) if b > a else (0, 0) for a, b in SL]
        slab_ids = [np.arange(lo, hi) for lo, hi in ctx]
        s0, s1 = SL[me]
        lo, hi = ctx[me]
        masks = []
        slab = self._dense(hi - lo, nx, _lib.F32)
        redistribute_rows(comm, fabs.tensor(), shifted, slab_ids, slab.tensor()[:hi - lo], key=("fabs", ny, nx))
        TRACE.mark("fabs to slabs")
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
        TRACE.mark("peak detector")
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
        TRACE.mark("mask back + assemble")

        # ---- inverse along x: (1 - mask) * F on this rank's spectrum rows (the Hermitian, odd x odd, path needs only
        # ky <= ny/2) + all-to-all: my x columns, all ky -------------------------------------------------------------
        counts_i1 = [(k["nlo"] if odd else k["rows"]) for k in KL]
        xa, xb = XB[me]
        segs = [[(0, k["a"], k["nlo"])] + ([(k["nlo"], k["hlo"], k["nhi"])] if (not odd and k["nhi"]) else []) for k in KL]
        if peers is not None:
            bt = peers.view("bt", xb - xa, ny, _lib.C64)
            self._pass_scatter(plan, 0, LOAD_MASKED, fshift, counts_i1[me],
                               [(XB[j][0], XB[j][1], peers.ptr(j, "bt"), ny, 0) for j in range(W)], segs[me], mask=mask,
                               shift_cols=nx // 2, inverse=1)
            comm.device_barrier()
        else:
            bt = self._dense(xb - xa, ny, _lib.C64)
            self._pass_exchange(plan, 0, LOAD_MASKED, fshift, counts_i1, nx, [[x] for x in XB], segs, bt, _lib.C64, mask=mask,
                                shift_cols=nx // 2, inverse=1)
        del fshift
        TRACE.mark("I1 rows + T3")
        if odd:
            _lib.check(lib.hd_conj_mirror_fill(bt.ptr, bt.pitch, xb - xa, ny, dev.stream_ptr()))
        # inverse along y: real output (two columns per transform on the Hermitian path) + all-to-all: my rows, all x
        if peers is not None:
            pitch, _ = peers.dem_pitch[me]
            dbuf = peers.arena[peers.offsets["dem"]:peers.offsets["dem"] + (self.up + rows + self.down) * pitch * 4]
            out_ext = ExtRaster(dev.DeviceRaster(dbuf.view(torch.float32).view(self.up + rows + self.down, pitch),
                                                 self.up + rows + self.down, nx, _lib.F32, np.float64), self.up, rows, self.down)
            self._pass_scatter(plan, 1, LOAD_HPAIR if odd else LOAD_C64, bt, xb - xa,
                               [(RB[j][0], RB[j][1], peers.ptr(j, "dem") + peers.dem_pitch[j][1] * peers.dem_pitch[j][0] * 4,
                                 peers.dem_pitch[j][0], 0) for j in range(W)], [(0, xa, xb - xa)], inverse=1, real_out=1)
            comm.device_barrier()
        else:
            out_ext = out_ext or self.alloc_ext(_lib.F32, np.float64)
            self._pass_exchange(plan, 1, LOAD_HPAIR if odd else LOAD_C64, bt, counts_xb, ny, [[rb] for rb in RB],
                                [[(0, XB[i][0], counts_xb[i])] for i in range(W)], out_ext.owned(), _lib.F32, inverse=1, real_out=1)
        del bt
        TRACE.mark("I2 cols + T4")
        return out_ext

    # -- the whole chain on row bands (BASELINE.json configs[4]: one mosaic over the GPUs of a box) --------------------
    def conditioning_chain(self, srtm, groves_class, hsheds, groves_iterations=3, with_hydrology=True, keep_complete=False,
                           ready=None, on_ready=None):
        """HydroDEMProcess.start (hydro_dem_process.py:122-153) on this rank's rows of a mosaic.  Inputs: device rasters
        of the band's rows (F32, U8 0/1, F32), or ExtRaster objects whose owned rows are already in place (groves,
        hsheds: saves the input copy).  Returns {"final", "dem_complete", "filled", "d8"} for the band's rows -- every
        one of them bit-identical to the single-GPU ConditioningChain on the whole mosaic."""
        from .pipeline import ConditioningChain
        chain = ConditioningChain(groves_iterations=groves_iterations, with_hydrology=False, keep_complete=keep_complete)
        TRACE.mark("chain:start")
        g_ext = groves_class if isinstance(groves_class, ExtRaster) else self.extended(dev.convert(groves_class, _lib.U8))
        h_ext = hsheds if isinstance(hsheds, ExtRaster) else self.extended(hsheds)
        cur = torch.cuda.current_stream()

        def input_halo(name, ext, given):
            if ready and ready.get(name) is not None:
                cur.wait_event(ready[name])               # its upload ran underneath the stages before
            if given is ext:
                self.exchange_halo(ext)

        if ready is None:
            input_halo("groves", g_ext, groves_class)
            input_halo("hsheds", h_ext, hsheds)
            TRACE.mark("input halos")
        elif ready.get("srtm") is not None:
            cur.wait_event(ready["srtm"])
        dem = self.detect_apply_fourier(srtm)                                        # image_srtm.py:125-126
        st = {"fourier": dem.raster}
        if ready is not None:
            input_halo("hsheds", h_ext, hsheds)
        chain._stage_lagoons(st, h_ext.raster)                                       # LagoonsDetection: HydroSHEDS only
        TRACE.mark("lagoons")
        if ready is not None:
            input_halo("groves", g_ext, groves_class)
        self.exchange_halo(dem)
        TRACE.mark("dem halo")
        chain._stage_groves(st, g_ext.raster)                                        # image_srtm.py:177-199
        TRACE.mark("groves")
        chain._stage_final(st, None)                                                 # final terms ... :149
        TRACE.mark("combine")
        own = lambda r: r.sub(self.up, self.up + self.rows, 0, self.nx)              # noqa: E731
        out = {"final": own(st["final"])}
        if on_ready:
            on_ready("final", out["final"])
        if keep_complete:
            out["dem_complete"] = own(st["dem_complete"])
        if with_hydrology:
            out["filled"], out["d8"] = self.sinkfill(ExtRaster(st["final32"], self.up, self.rows, self.down))
            TRACE.mark("sink-fill + D8")
            if on_ready:
                on_ready("filled", out["filled"])
                on_ready("d8", out["d8"])
        return out

    # -- sink-fill + D8 -----------------------------------------------------------------------------------------------------
    def sinkfill(self, z, max_rounds=10000):
        """Banded Planchon-Darboux fixed point + D8 (bit-identical to the single-GPU result).  ``z``: F32 device raster
        of the band's rows, or an ExtRaster whose owned rows hold them (its margin rows are overwritten)."""
        lib = _lib.load()
        comm = self.comm
        # The bands OVERLAP by FILL_OVERLAP rows on either side of a cut (an alternating Schwarz iteration converges in far
        # fewer rounds with overlap: a drainage path has to wander that many rows across a cut before it costs another
        # round).  The outermost row of the extended band is the boundary row received from the neighbour; the rows in
        # between are relaxed by both ranks and agree at the fixed point.
        ov = self.fill_overlap if comm.world > 1 else 1
        src = z.owned() if isinstance(z, ExtRaster) else dev.convert(z, _lib.F32)
        up, down = (ov if comm.rank > 0 else 0), (ov if comm.rank < comm.world - 1 else 0)
        zx = ExtRaster(dev.empty(up + self.rows + down, self.nx, _lib.F32, np.float32), up, self.rows, down)
        zx.owned().tensor().copy_(src.tensor())
        z = zx
        self.exchange_halo(z, ov)
        zv, n_up, n_down = z.with_halo(ov)
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
        wc_g = None
        TRACE.mark("fill: overlap exchange")
        if comm.world > 1:
            # global multigrid start: pool my band, all-gather the coarse DEM, fill the whole coarse mosaic on every rank
            nyc, nxc = -(-self.ny // 8), -(-nx // 8)
            cp = (nxc + 31) // 32 * 32
            zc_g = torch.empty((nyc, cp), dtype=torch.float32, device=dev.device())
            wc_g = torch.empty((nyc, cp), dtype=torch.float32, device=dev.device())
            cb = [(a // 8, -(-b // 8)) for a, b in self.bounds]
            c0, c1 = cb[comm.rank]
            scratch = torch.zeros(((c1 - c0) // 64 + 2) * (nxc // 64 + 2), dtype=torch.int32, device=dev.device())
            own = z.owned()
            _lib.check(lib.hd_fill_pool_band(own.ptr, own.pitch, rows, nx, ctypes.c_void_p(zc_g[c0].data_ptr()),
                                             ctypes.c_void_p(wc_g[c0].data_ptr()), cp, flags,
                                             ctypes.c_void_p(scratch.data_ptr()), dev.stream_ptr()))
            for g in (zc_g, wc_g):
                comm.p2p([(g[c0:c1], j) for j in range(comm.world) if j != comm.rank],
                         [(g[a:b], i) for i, (a, b) in enumerate(cb) if i != comm.rank])
            cbytes = lib.hd_pdfill_workspace_bytes(nyc, nxc)
            cwork = dev.scratch(cbytes)
            _lib.check(lib.hd_pdfill_coarse(ctypes.c_void_p(zc_g.data_ptr()), ctypes.c_void_p(wc_g.data_ptr()), cp, nyc, nxc,
                                            ctypes.c_void_p(cwork.data_ptr()), cbytes, dev.stream_ptr()))
        TRACE.mark("fill: global coarse level")
        while True:
            if rounds == 0 and wc_g is not None:
                _lib.check(lib.hd_pdfill_band_start(zv.ptr, zv.pitch, w.ptr, w.pitch, zv.ny, nx, wp, nbytes, flags,
                                                    ctypes.c_void_p(wc_g.data_ptr()), cp, self.r0 - n_up, dev.stream_ptr()))
            else:
                _lib.check(lib.hd_pdfill_band(zv.ptr, zv.pitch, w.ptr, w.pitch, zv.ny, nx, wp, nbytes,
                                              flags | (1 if rounds else 0), None, dev.stream_ptr()))
            rounds += 1
            if comm.world == 1:
                break
            lowered.zero_()
            sends, recvs = [], []
            if n_up:                                               # my first / last OWNED rows go to the neighbours' halo rows
                sends.append((tw[2 * n_up - 1, :nx], comm.rank - 1))      # = the neighbour's last (boundary) row
                recvs.append((recv_up, comm.rank - 1))
            if n_down:
                sends.append((tw[n_up + rows - n_down, :nx], comm.rank + 1))   # = the neighbour's first (boundary) row
                recvs.append((recv_down, comm.rank + 1))
            comm.p2p(sends, recvs)
            lp = ctypes.c_void_p(lowered.data_ptr())
            if n_up:
                _lib.check(lib.hd_halo_min_flag(w.ptr, ctypes.c_void_p(recv_up.data_ptr()), nx, lp, dev.stream_ptr()))
            if n_down:
                last = w.sub(zv.ny - 1, zv.ny, 0, nx)
                _lib.check(lib.hd_halo_min_flag(last.ptr, ctypes.c_void_p(recv_down.data_ptr()), nx, lp, dev.stream_ptr()))
            comm.allreduce_max_(lowered)
            TRACE.mark(f"fill: round {rounds}")
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
        # three streams: the SRTM rows go up first, groves + HydroSHEDS follow underneath the Fourier stage, and the
        # final DEM travels down while the sink-fill rounds run (what ConditioningChain.apply_to_host does on one GPU)
        cur = torch.cuda.current_stream()
        if not hasattr(self, "_copy_streams"):
            self._copy_streams = (torch.cuda.Stream(), torch.cuda.Stream())
        up, down = self._copy_streams
        g_ext, h_ext = self.alloc_ext(_lib.U8, np.uint8), self.alloc_ext(_lib.F32, np.float32)
        up.wait_stream(cur)
        ready, keep = {}, []

        def send_up(name, host, into=None):
            if into is None:
                r, ready[name] = dev.upload_async(host, up)
            else:
                r, ready[name] = into, dev.upload_into(into, host, up)
            if not dev._is_pinned(host):
                up.synchronize()                              # a pageable source may be released by the caller
            keep.append(host)
            return r

        srtm = send_up("srtm", np.ascontiguousarray(srtm_rows, dtype=np.float32))
        send_up("hsheds", np.ascontiguousarray(hsheds_rows, dtype=np.float32), h_ext.owned())
        g = np.ascontiguousarray(groves_rows)
        send_up("groves", g.view(np.uint8) if g.dtype == np.bool_ else g.astype(np.uint8, copy=False), g_ext.owned())
        pending = {}

        narrow = {}

        def send_down(name, raster):
            if self.narrow_final and raster.dtype == _lib.F32 and raster.ref_dtype == np.float64:
                from .pipeline import NarrowDownload
                narrow[name] = NarrowDownload(raster, cur, down)        # int16 over PCIe, widened on host threads
                return
            conv = dev.convert(raster, dev.hd_dtype_of(raster.ref_dtype))
            ev = torch.cuda.Event()
            ev.record(cur)
            down.wait_event(ev)
            pending[name] = dev.download_async(conv, down) + (conv,)

        self.conditioning_chain(srtm, g_ext, h_ext, ready=ready, on_ready=send_down)
        out = {}
        for name, (host, ev, _conv) in pending.items():
            ev.synchronize()
            out[name] = host
        for name, nd in narrow.items():
            out[name] = nd.result()
        self.last_transfer_bytes = (sum(h.nbytes for h in keep),
                                    sum(h.nbytes for h, _e, _c in pending.values())
                                    + sum(nd.nbytes + nd.extra_bytes for nd in narrow.values()))
        return {k: out[k] for k in ("final", "filled", "d8")}
