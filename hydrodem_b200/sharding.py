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

The Fourier stage needs a global transform (all-to-all transpose); that is
not built yet, so the banded path covers the stencil
stages and sink-fill / D8 (BASELINE.json configs[3]).

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
