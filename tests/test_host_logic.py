"""CPU: host-side logic that needs no GPU -- reference-compatible surface of the filter package, band
partitioning, and the halo-exchange communicator over gloo (world_size 2)."""
import inspect
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

import hydrodem_b200
from hydrodem_b200 import sharding
from hydrodem_b200.exceptions import (CenterCloseBorderError, HydroDEMException, NumpyArrayExpectedError,
                                      WindowSizeEvenError, WindowSizeHighError)
from hydrodem_b200.filters import ComposedFilter, ComposedFilterResults, Filter, custom_filters as cf
from hydrodem_b200.filters import extension_filters as ef, simple_filters as sf


def test_reference_class_surface():
    """Every filter class of the reference's three modules exists with the same constructor keywords."""
    want = {
        cf: dict(MajorityFilter=["window_size"], ExpandFilter=["window_size"], QuadraticFilter=["window_size"],
                 CorrectNANValues=["window_size"], IsolatedPoints=["window_size"], BlanksFourier=["window_size"],
                 DetectBlanksFourier=[], MaskNegatives=[], MaskPositives=[], MaskTallGroves=[], MaskFourier=[],
                 TidyingLagoons=[], LagoonsDetection=[], GrovesCorrection=["groves_class"],
                 GrovesCorrectionsIter=["groves_class", "iterations"], FourierInitial=[],
                 FourierProcessQuarters=["fft_transform_abs"], DetectApplyFourier=[], PostProcessingFinal=[]),
        ef: dict(BitwiseXOR=["operand"], AbsoluteValues=[], Around=[], Convolve=["weights"], BinaryErosion=["iterations"],
                 BinaryClosing=["structure"], GreyDilation=["size"], FourierTransform=[], FourierITransform=[],
                 FourierShift=[], FourierIShift=[]),
        sf: dict(LowerThan=["value"], GreaterThan=["value"], BooleanToInteger=[], ProductFilter=["factor"],
                 AdditionFilter=["addend"], SubtractionFilter=["minuend"]),
    }
    for mod, classes in want.items():
        for name, params in classes.items():
            cls = getattr(mod, name)
            assert issubclass(cls, Filter), name
            got = [] if cls.__init__ is object.__init__ else \
                [p for p in inspect.signature(cls.__init__).parameters if p != "self"]
            assert got == params, (name, got)
    # keyword-only constructors as in the reference (custom_filters.py:39, simple_filters.py:24, ...)
    with pytest.raises(TypeError):
        cf.MajorityFilter(11)
    with pytest.raises(TypeError):
        sf.LowerThan(0.0)
    assert sf.ProductFilter(3).factor == 3 and sf.AdditionFilter(2).addend == 2


def test_exceptions_match_reference_messages():
    assert str(WindowSizeHighError(7, (5, 5))) == "Window size: 7 cannot be higher than grid dimensions: (5, 5)"
    assert str(WindowSizeEvenError(4)) == "Window size: 4 cannot be an even number"
    assert str(CenterCloseBorderError((0, 1), 3)) == "Center of window: (0, 1) too close of border. Window size: 3"
    assert "Expected numpy ndarray type" in str(NumpyArrayExpectedError([1]))
    assert issubclass(WindowSizeEvenError, HydroDEMException)


def test_boundary_checks_run_before_any_device_work():
    """The reference's argument errors do not need a GPU (sliding_window.py:130-156 order)."""
    with pytest.raises(NumpyArrayExpectedError):
        cf.MajorityFilter(window_size=11).apply([[1.0]])
    with pytest.raises(WindowSizeHighError):
        cf.ExpandFilter(window_size=7).apply(np.zeros((5, 9)))
    with pytest.raises(WindowSizeEvenError):
        cf.QuadraticFilter(window_size=4).apply(np.zeros((9, 9)))
    with pytest.raises(WindowSizeHighError):
        cf.ExpandFilter(window_size=6).apply(np.zeros((5, 9)))      # too large wins over even
    assert isinstance(cf.LagoonsDetection(), ComposedFilterResults) and isinstance(cf.MaskFourier(), ComposedFilter)


def test_install_as_reference_filters():
    saved = {k: sys.modules.get(k) for k in ("filters", "filters.custom_filters", "exceptions")}
    try:
        hydrodem_b200.install_as_reference_filters()
        from filters.custom_filters import MajorityFilter          # the reference's flat import style
        from filters import Filter as F2
        from exceptions import WindowSizeEvenError as E2
        assert MajorityFilter is cf.MajorityFilter and F2 is Filter and E2 is WindowSizeEvenError
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def test_band_bounds():
    assert sharding.band_bounds(10, 3) == [(0, 4), (4, 7), (7, 10)]
    b = sharding.band_bounds(36000, 8)
    assert b[0] == (0, 4500) and b[-1] == (31500, 36000)
    for ny, w in ((3601, 8), (5, 5), (18000, 4)):
        bb = sharding.band_bounds(ny, w)
        assert bb[0][0] == 0 and bb[-1][1] == ny and all(a[1] == c[0] for a, c in zip(bb, bb[1:]))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _halo_worker(rank, world, port, ny, nx, h, out):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        comm = sharding.DistComm()
        mosaic = torch.arange(ny * nx, dtype=torch.float32).reshape(ny, nx)
        r0, r1 = sharding.band_bounds(ny, world)[rank]
        band = mosaic[r0:r1].clone()
        up, down = comm.exchange(band[:h], band[-h:], band[:h], band[-h:])
        ok = True
        if rank > 0:
            ok &= bool(torch.equal(up, mosaic[r0 - h:r0]))
        else:
            ok &= up is None
        if rank < world - 1:
            ok &= bool(torch.equal(down, mosaic[r1:r1 + h]))
        else:
            ok &= down is None
        ok &= comm.any(rank == 1) is True and comm.any(False) is False
        out[rank] = int(ok)
    finally:
        dist.destroy_process_group()


def test_halo_exchange_gloo_world2():
    world = 2
    out = mp.Array("i", [0] * world)
    port = _free_port()
    procs = [mp.Process(target=_halo_worker, args=(r, world, port, 23, 17, 3, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert list(out) == [1, 1]


def test_thread_comm_emulation():
    def fn(comm):
        x = torch.full((2, 4), float(comm.rank))
        up, down = comm.exchange(x[:1], x[-1:], x[:1], x[-1:])
        return (None if up is None else float(up[0, 0]), None if down is None else float(down[0, 0]), comm.any(comm.rank == 2))
    res = sharding.ThreadComm.run(3, fn)
    assert res == [(None, 1.0, True), (0.0, 2.0, True), (1.0, None, True)]


def _a2a_worker(rank, world, port, ny, nx, out):
    """The all-to-all of the distributed fft2 over gloo: block (rows of rank i, columns of rank j) of a matrix must
    land on rank j, i.e. the exchange of transposed blocks reassembles full columns."""
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        comm = sharding.DistComm()
        mat = torch.arange(ny * nx, dtype=torch.float32).reshape(ny, nx)
        rows_b, cols_b = sharding.band_bounds(ny, world), sharding.band_bounds(nx, world)
        r0, r1 = rows_b[rank]
        t = mat[r0:r1].t().contiguous()                              # (nx, my rows): the transposed local band
        chunks = [t[a:b] for (a, b) in cols_b]
        mine = cols_b[rank][1] - cols_b[rank][0]
        got = comm.all_to_all_shaped(chunks, [(mine, b - a) for (a, b) in rows_b])
        cols = torch.cat(got, dim=1)                                 # (my cols, ny)
        c0, c1 = cols_b[rank]
        out[rank] = int(torch.equal(cols, mat[:, c0:c1].t()))
    finally:
        dist.destroy_process_group()


def test_all_to_all_gloo_world2():
    world = 2
    out = mp.Array("i", [0] * world)
    port = _free_port()
    procs = [mp.Process(target=_a2a_worker, args=(r, world, port, 11, 7, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert list(out) == [1, 1]


def test_thread_comm_all_to_all():
    def fn(comm):
        chunks = [torch.full((1, 2), float(10 * comm.rank + j)) for j in range(comm.world)]
        got = comm.all_to_all_shaped(chunks, [(1, 2)] * comm.world)
        return [float(g[0, 0]) for g in got]
    res = sharding.ThreadComm.run(3, fn)
    assert res == [[0.0, 10.0, 20.0], [1.0, 11.0, 21.0], [2.0, 12.0, 22.0]]
