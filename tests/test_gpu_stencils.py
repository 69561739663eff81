"""GPU parity: the CUDA path (through the filter classes and the C ABI) against the reference goldens
and the CPU oracle.  Bit-exact unless a tolerance is written in the test."""
import numpy as np
import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from hydrodem_b200.filters import custom_filters as cf
    from hydrodem_b200.filters import extension_filters as ef
    from hydrodem_b200.filters import simple_filters as sf
    from hydrodem_b200.exceptions import (NumpyArrayExpectedError, WindowSizeEvenError, WindowSizeHighError)
    from hydrodem_b200.synth import SynthScene
    from oracle import stencils, morphology, chain, clib


def eq(a, b):
    assert a.dtype == b.dtype, (a.dtype, b.dtype)
    assert a.shape == b.shape
    np.testing.assert_array_equal(a, b)


# ---- reference goldens through the drop-in classes ------------------------------------------------
def test_g1_majority_golden():
    g = load_golden("ref_lagoons")
    got = cf.MajorityFilter(window_size=11).apply(g["hsheds_nan_values_expected"])
    assert got.dtype == np.float64
    np.testing.assert_array_equal(got, g["hsheds_majority_11_expected"])


def test_g2_lagoons_golden():
    g = load_golden("ref_lagoons")
    inp = g["hsheds_nan_values_expected"].copy()
    lag = cf.LagoonsDetection()
    ret = lag.apply(inp)
    np.testing.assert_array_equal(lag.results["TidyingLagoons"], g["lagoons_expected"])
    assert lag.results["CorrectNANValues"] is inp and lag.hsheds_nan_fixed is inp
    assert ret.dtype == np.int64 and lag.results["MajorityFilter"].dtype == np.float64
    np.testing.assert_array_equal(lag.mask_lagoons, (g["lagoons_expected"] > 0) * 1)


def test_g3_expand_golden():
    g = load_golden("ref_mask_fourier")
    got = cf.ExpandFilter(window_size=13).apply(g["isolated_filter_expected"])
    eq(got, g["mask_fourier_expected"].astype(np.float64))


# ---- outputs of the reference classes on seeded inputs ---------------------------------------------
def test_run_stencils_fixtures():
    g = load_golden("run_stencils")
    eq(cf.MajorityFilter(window_size=11).apply(g["maj_in"]), g["maj11_out"])
    eq(cf.MajorityFilter(window_size=5).apply(g["maj_in"]), g["maj5_out"])
    assert cf.MajorityFilter(window_size=3).apply(g["maj_in"]).sum() == 0
    for ws in (3, 7, 13):
        eq(cf.ExpandFilter(window_size=ws).apply(g["expand_in"]), g[f"expand{ws}_out"])
    for tag in ("nanfix", "nanfix_f"):
        inp = g[tag + "_in"].copy()
        got = cf.CorrectNANValues().apply(inp)
        assert got is inp
        assert got.tobytes() == g[tag + "_out"].tobytes() or np.array_equal(got, g[tag + "_out"], equal_nan=True)
    inp = g["iso_in"].copy()
    got = cf.IsolatedPoints(window_size=3).apply(inp)
    assert got is inp
    eq(got, g["iso_out"])


def test_run_lagoons_fixtures():
    g = load_golden("run_lagoons")
    lag = cf.LagoonsDetection()
    ret = lag.apply(g["hsheds"].copy())
    eq(lag.results["CorrectNANValues"], g["nanfixed"])
    eq(lag.results["MajorityFilter"], g["majority"])
    eq(lag.results["TidyingLagoons"], g["tidying"])
    eq(ret, g["mask"])
    eq(ef.BinaryErosion(iterations=2).apply(g["majority"]), g["erosion2"])
    eq(ef.BinaryClosing().apply(g["groves_raw"]), g["closing_cross"])
    eq(ef.BinaryClosing(structure=np.ones((3, 3))).apply(g["groves_raw"]), g["closing_full"])
    eq(ef.GreyDilation(size=(7, 7)).apply(g["majority"]), g["greydil7"])
    conv = ef.Convolve().apply(g["dem64"])
    assert conv.tobytes() == g["conv3"].tobytes()              # double accumulation in scipy's order: bit-exact
    eq(cf.PostProcessingFinal().apply(g["dem64"]), g["post"])


def test_simple_filters_fixtures():
    g = load_golden("run_simple")
    a, b = g["a"], g["b"]
    eq(sf.LowerThan(value=0.0).apply(a), g["lower"])
    eq(sf.GreaterThan(value=1.5).apply(a), g["greater"])
    eq(sf.BooleanToInteger().apply(a > 0), g["b2i"])
    eq(sf.ProductFilter(factor=b).apply(a), g["prod"])
    eq(sf.ProductFilter(factor=3).apply(a), g["prod_s"])
    eq(sf.AdditionFilter(addend=b).apply(a), g["add"])
    eq(sf.SubtractionFilter(minuend=1).apply(a), g["sub"])
    eq(sf.SubtractionFilter(minuend=b).apply(a), g["sub_a"])
    eq(ef.AbsoluteValues().apply(a), g["absv"])
    eq(ef.Around().apply(a * np.float32(1.25)), g["around"])
    eq(ef.BitwiseXOR(operand=(a > 0)).apply(b > 0), g["xor"])


# ---- oracle side by side, ragged sizes (tiles are 32 x 128) ------------------------------------------
@pytest.mark.parametrize("shape", [(33, 129), (64, 128), (200, 517), (31, 40), (15, 300)])
def test_ragged_sizes_vs_oracle(shape):
    ny, nx = shape
    sc = SynthScene(ny, nx, 21)
    hs = sc.hsheds()
    rng = np.random.default_rng(ny * nx)
    hs[rng.random(shape) < 0.01] = -32768.0
    if min(shape) >= 11:
        eq(cf.MajorityFilter(window_size=11).apply(hs), stencils.majority(hs, 11))
    mask = (rng.random(shape) < 0.01).astype(np.float32)
    for ws in (3, 7, 13):
        if min(shape) >= ws:
            eq(cf.ExpandFilter(window_size=ws).apply(mask), stencils.expand(mask, ws))
    with np.errstate(all="ignore"):
        eq(cf.CorrectNANValues().apply(hs.copy()), stencils.correct_nan(hs.copy()))
    eq(cf.IsolatedPoints(window_size=3).apply(mask.copy()), stencils.isolated_points(mask.copy()))
    m = rng.random(shape) < 0.7
    eq(ef.BinaryErosion(iterations=2).apply(m), morphology.binary_erosion(m, iterations=2))
    eq(ef.BinaryErosion(iterations=1).apply(m.astype(np.float32)), morphology.binary_erosion(m, iterations=1))
    eq(ef.BinaryClosing().apply(~m), morphology.binary_closing(~m))
    eq(ef.BinaryClosing(structure=np.ones((3, 3))).apply(~m), morphology.binary_closing(~m, np.ones((3, 3))))
    a32 = sc.srtm()
    if min(shape) >= 7:
        eq(ef.GreyDilation(size=(7, 7)).apply(a32), morphology.grey_dilation_square(a32, 7))
        eq(ef.GreyDilation(size=(7, 7)).apply(a32.astype(np.float64) * 1.1),
           morphology.grey_dilation_square(a32.astype(np.float64) * 1.1, 7))
    eq(cf.PostProcessingFinal().apply(a32.astype(np.float64) * 1.1), stencils.mean3_round(a32.astype(np.float64) * 1.1))
    eq(cf.PostProcessingFinal().apply(a32), stencils.mean3_round(a32).astype(np.float32))


def test_majority_edge_cases():
    """NaN keys never win, -0.0 / 0.0 share a key and the first inserted one is returned."""
    a = np.full((40, 150), 5.0, dtype=np.float32)
    a[10:30, 20:60] = np.nan
    a[5:25, 100:140] = -0.0
    a[6, 101] = 0.0
    got = cf.MajorityFilter(window_size=11).apply(a)
    want = stencils.majority(a, 11)
    eq(got, want)
    assert np.isnan(got).sum() == 0


def test_full_size_tile_vs_c_oracle():
    """C2 size (3601 x 3601): majority / expand against the C oracle."""
    sc = SynthScene(3601, 3601, 1002)
    hs = sc.hsheds()
    got = cf.MajorityFilter(window_size=11).apply(hs)
    want = clib.majority(hs, 11, 85)
    eq(got, want)
    assert (want > 0).sum() > 10000
    eq(cf.ExpandFilter(window_size=7).apply(want), clib.expand(want, 7))


# ---- error behaviour of the boundary (tests/test_sliding_window.py:78-133 of the reference) ------------
def test_boundary_errors():
    with pytest.raises(NumpyArrayExpectedError):
        cf.MajorityFilter(window_size=11).apply([[1, 2], [3, 4]])
    with pytest.raises(NumpyArrayExpectedError):
        sf.LowerThan(value=0).apply("x")
    with pytest.raises(WindowSizeEvenError):
        cf.ExpandFilter(window_size=4).apply(np.zeros((9, 9)))
    with pytest.raises(WindowSizeHighError):
        cf.ExpandFilter(window_size=11).apply(np.zeros((9, 20)))
    with pytest.raises(WindowSizeHighError):
        cf.MajorityFilter(window_size=10).apply(np.zeros((9, 20)))     # too large is reported before even


def test_int16_production_dtype_lagoons():
    """gdal ReadAsArray hands the reference INT16 rasters (image_hsheds.py:133-135).  CorrectNANValues then stores its
    float32 means into the caller's int16 array -- truncation on assignment -- and MajorityFilter counts those integers.
    Fixture: the unmodified reference on an int16 HydroSHEDS raster with a run of voids (non-integer means)."""
    from hydrodem_b200 import device as dev
    g = load_golden("run_int16")
    hs = g["hsheds_i16"].copy()
    assert hs.dtype == np.int16
    rt = dev.download(dev.upload(hs))
    assert rt.dtype == np.int16
    np.testing.assert_array_equal(rt, hs)
    lag = cf.LagoonsDetection()
    ret = lag.apply(hs)
    assert lag.results["CorrectNANValues"] is hs and hs.dtype == np.int16          # in place, caller's array
    np.testing.assert_array_equal(hs, g["lag_CorrectNANValues"])
    assert (g["lag_CorrectNANValues"] != g["hsheds_i16"]).sum() >= 5                # the voids were filled
    for k in ("MajorityFilter", "TidyingLagoons", "MaskPositives"):
        got = lag.results[k]
        assert got.dtype == g["lag_" + k].dtype, k
        np.testing.assert_array_equal(got, g["lag_" + k])
    np.testing.assert_array_equal(ret, g["lag_return"])


@pytest.mark.parametrize("shape", [(97, 131), (64, 128), (33, 260), (300, 7), (7, 300), (161, 129)])
def test_tidy_lagoons_fused_kernel(shape):
    """hd_tidy_lagoons (custom_filters.py:587-610 in one kernel) against the four filter classes and the oracle: blobs on
    the frame, across tile seams (x = 128, y = 32) and thin ones that the erosion removes."""
    from hydrodem_b200 import _lib, device as dev
    rng = np.random.default_rng(shape[0] * 1000 + shape[1])
    ny, nx = shape
    maj = np.zeros(shape, dtype=np.float32)
    for _ in range(12):
        y, x = rng.integers(0, ny), rng.integers(0, nx)
        hh, ww = rng.integers(1, 14), rng.integers(1, 14)
        maj[max(0, y - hh):y + hh, max(0, x - ww):x + ww] = rng.integers(1, 4)
    maj[rng.random(shape) < 0.01] = 0
    lib = _lib.load()
    src = dev.upload(maj)
    out = dev.empty(ny, nx, _lib.F32, np.float64)
    _lib.check(lib.hd_tidy_lagoons(src.ptr, src.pitch, out.ptr, out.pitch, ny, nx, dev.stream_ptr()))
    got = dev.download(out)
    eroded = ef.BinaryErosion(iterations=2).apply(maj)
    expanded = cf.ExpandFilter(window_size=7).apply(eroded)
    want = ef.GreyDilation(size=(7, 7)).apply(sf.ProductFilter(factor=maj).apply(expanded))
    np.testing.assert_array_equal(got, want)
    np.testing.assert_array_equal(got, cf.TidyingLagoons().apply(maj))
    np.testing.assert_array_equal(got, chain.tidying_lagoons(maj))
