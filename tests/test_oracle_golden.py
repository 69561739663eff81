"""CPU: the oracle against (1) the reference's stored goldens G1-G4 and (2)
outputs of the reference classes themselves (tests/golden/run_*.npz, made by
tests/golden/gen_golden.py in the build container)."""
import numpy as np
import pytest
from scipy import ndimage

from oracle import windows, stencils, morphology, fourier, hydrology, chain, clib

from conftest import load_golden


def eq(a, b):
    """Reference test idiom: numpy.testing.assert_array_equal (NaN == NaN, -0 == 0), plus dtype."""
    assert a.dtype == b.dtype, (a.dtype, b.dtype)
    np.testing.assert_array_equal(a, b)


# ---- G1-G4: the reference's own stored expectations ---------------------------
def test_g1_majority_11():
    g = load_golden("ref_lagoons")
    got = stencils.majority(g["hsheds_nan_values_expected"], 11)
    assert got.dtype == np.float64
    np.testing.assert_array_equal(got, g["hsheds_majority_11_expected"])
    np.testing.assert_array_equal(clib.majority(g["hsheds_nan_values_expected"], 11, 85), got)


def test_g2_lagoons_chain():
    g = load_golden("ref_lagoons")
    res = chain.lagoons_detection(g["hsheds_nan_values_expected"].copy())
    np.testing.assert_array_equal(res["TidyingLagoons"], g["lagoons_expected"])


def test_g3_expand_13():
    g = load_golden("ref_mask_fourier")
    got = stencils.expand(g["isolated_filter_expected"], 13)
    np.testing.assert_array_equal(got, g["mask_fourier_expected"])
    np.testing.assert_array_equal(clib.expand(g["isolated_filter_expected"], 13), got)


def test_g4_blanks_clean_spectrum():
    g = load_golden("ref_mask_fourier")
    mask, _ = stencils.blanks_fourier(g["filtered_blank_expected_2"], 55)
    assert mask.sum() == 0


# ---- A0 window iterators ---------------------------------------------------------------------
@pytest.mark.parametrize("tag,grid,ws,variant", [
    ("plain3", "grid5", 3, {}),
    ("circ5", "grid7", 5, dict(circular=True)),
    ("nocenter3", "grid5", 3, dict(no_center=True)),
    ("inner5_3", "grid7", 5, dict(inner_size=3)),
    ("ignore3", "grid5", 3, dict(ignore_border=True)),
    ("combo5_3", "grid7", 5, dict(ignore_border=True, inner_size=3, no_center=True)),
])
def test_window_kats(tag, grid, ws, variant):
    g = load_golden("run_windows")
    w, jj, ii = windows.all_windows(g[grid], ws, **variant)
    got = w.reshape(-1, ws, ws)
    eq(got, g[tag + "_w"])
    cen = np.array([(j, i) for j in jj for i in ii])
    np.testing.assert_array_equal(cen, g[tag + "_c"])
    j, i = g[tag + "_c"][len(cen) // 2]
    np.testing.assert_array_equal(windows.window_at(g[grid], ws, j, i, **variant), g[tag + "_w"][len(cen) // 2])


def test_window_guards():
    a = np.zeros((5, 7))
    with pytest.raises(windows.OracleWindowError, match="even"):
        windows.check_window_size(a.shape, 4)
    with pytest.raises(windows.OracleWindowError, match="high"):
        windows.check_window_size(a.shape, 7)
    with pytest.raises(windows.OracleWindowError, match="high"):
        windows.check_window_size(a.shape, 6)          # too large is tested before even
    with pytest.raises(windows.OracleWindowError, match="border"):
        windows.window_at(a, 3, 0, 3)


# ---- stencils vs reference runs -----------------------------------------------------------------
def test_correct_nan_matches_reference():
    g = load_golden("run_stencils")
    for tag in ("nanfix", "nanfix_f"):
        inp = g[tag + "_in"].copy()
        got = stencils.correct_nan(inp)
        assert got is inp
        eq(got, g[tag + "_out"])
    assert np.isnan(g["nanfix_out"]).sum() > 1            # the empty-neighbour case is covered


def test_np_pairwise_emulation():
    rng = np.random.default_rng(0)
    for n in range(0, 9):
        for _ in range(200):
            v = (rng.normal(0, 100, n) * 10.0 ** rng.integers(-3, 4)).astype(np.float32)
            assert stencils.np_pairwise_sum_f32(v).tobytes() == np.sum(v, dtype=np.float32).tobytes() or n == 0


def test_majority_matches_reference():
    g = load_golden("run_stencils")
    eq(stencils.majority(g["maj_in"], 11), g["maj11_out"])
    eq(stencils.majority(g["maj_in"], 5), g["maj5_out"])
    eq(clib.majority(g["maj_in"], 11, stencils.majority_min_count(11)), g["maj11_out"])
    eq(clib.majority(g["maj_in"], 5, stencils.majority_min_count(5)), g["maj5_out"])
    assert stencils.majority_min_count(11) == 85 and stencils.majority_min_count(5) == 17
    assert stencils.majority(g["maj_in"], 3).sum() == 0  # ws=3 can never fire


@pytest.mark.parametrize("ws", [3, 7, 13])
def test_expand_matches_reference(ws):
    g = load_golden("run_stencils")
    eq(stencils.expand(g["expand_in"], ws), g[f"expand{ws}_out"])
    eq(clib.expand(g["expand_in"], ws), g[f"expand{ws}_out"])


def test_isolated_matches_reference():
    g = load_golden("run_stencils")
    inp = g["iso_in"].copy()
    got = stencils.isolated_points(inp)
    assert got is inp
    eq(got, g["iso_out"])


def test_quadratic_matches_reference():
    g = load_golden("run_stencils")
    q32 = stencils.quadratic(g["srtm"], 15)
    assert q32.dtype == np.float32
    np.testing.assert_allclose(q32, g["quad32"], rtol=1e-6)
    q64 = stencils.quadratic(g["srtm64"], 15)
    assert q64.dtype == np.float64
    np.testing.assert_allclose(q64, g["quad64"], rtol=1e-6)
    # border ws//2 untouched
    np.testing.assert_array_equal(q64[:7], g["srtm64"][:7])
    # the equivalent fixed kernel (what the CUDA kernel applies)
    k = stencils.quadratic_kernel(15)
    assert abs(k.sum() - 1) < 1e-12
    corr = ndimage.correlate(g["srtm"].astype(np.float64), k, mode="constant")
    np.testing.assert_allclose(corr[7:-7, 7:-7], g["quad32"][7:-7, 7:-7], rtol=2e-6)


def test_groves_matches_reference():
    g = load_golden("run_stencils")
    eq(morphology.binary_closing(g["groves_raw"], np.ones((3, 3))), g["groves_closed"])
    np.testing.assert_allclose(stencils.groves_correction(g["srtm64"], g["groves_closed"]), g["groves1"], rtol=1e-6)
    np.testing.assert_allclose(stencils.groves_corrections_iter(g["srtm64"], g["groves_closed"], 3), g["groves3"],
                               rtol=1e-6)


# ---- lagoons / morphology / final --------------------------------------------------------------
def test_lagoons_matches_reference():
    g = load_golden("run_lagoons")
    res = chain.lagoons_detection(g["hsheds"].copy())
    eq(res["CorrectNANValues"], g["nanfixed"])
    eq(res["MajorityFilter"], g["majority"])
    eq(res["TidyingLagoons"], g["tidying"])
    eq(res["MaskPositives"], g["mask"])
    assert g["mask"].sum() > 100                          # lagoons are actually detected


def test_morphology_matches_reference_and_scipy():
    g = load_golden("run_lagoons")
    eq(morphology.binary_erosion(g["majority"], iterations=2), g["erosion2"])
    eq(morphology.binary_closing(g["groves_raw"]), g["closing_cross"])
    eq(morphology.binary_closing(g["groves_raw"], np.ones((3, 3))), g["closing_full"])
    eq(morphology.grey_dilation_square(g["majority"], 7), g["greydil7"])
    rng = np.random.default_rng(5)
    m = rng.random((40, 53)) < 0.6
    eq(morphology.binary_erosion(m, iterations=2), ndimage.binary_erosion(m, iterations=2))
    eq(morphology.binary_closing(m), ndimage.binary_closing(m))
    eq(morphology.binary_closing(m, np.ones((3, 3))), ndimage.binary_closing(m, structure=np.ones((3, 3))))
    a = rng.normal(0, 1, (40, 53))
    eq(morphology.grey_dilation_square(a, 7), ndimage.grey_dilation(a, size=(7, 7)))


def test_postprocessing_matches_reference():
    g = load_golden("run_lagoons")
    conv = stencils.convolve_reflect(g["dem64"], np.ones((3, 3))) / 9
    assert conv.tobytes() == g["conv3"].tobytes()         # bit-exact double accumulation order
    eq(stencils.mean3_round(g["dem64"]), g["post"])


# ---- Fourier -----------------------------------------------------------------------------------
def test_fourier_stage_matches_reference():
    g = load_golden("run_fourier")
    fabs, fshift = fourier.fourier_initial(g["srtm"])
    eq(fabs, g["fabs"])
    eq(fshift, g["fshift"])
    q1, q2 = fourier.first_quarters(g["fabs"])
    eq(np.ascontiguousarray(q1), g["q1"])
    eq(np.ascontiguousarray(q2), g["q2"])
    bm, bmod = stencils.blanks_fourier(g["q1"], 55)
    eq(bm, g["blanks_mask"])
    np.testing.assert_array_equal(bmod, g["blanks_mod"])
    eq(stencils.detect_blanks_fourier(g["q1"]), g["detect"])
    eq(stencils.mask_fourier(g["q1"]), g["mask_q1"])
    eq(stencils.mask_fourier(g["q2"]), g["mask_q2"])
    eq(fourier.process_quarters(g["fabs"]), g["mask"])
    assert g["mask"].sum() > 0                            # the stripes are detected
    corrected, mask, _ = fourier.detect_apply_fourier(g["srtm"])
    eq(mask, g["mask"])
    eq(corrected, g["corrected"])


@pytest.mark.parametrize("shape", [(131, 140), (140, 131), (131, 133), (134, 136)])
def test_mask_assembly_geometry(shape):
    g = load_golden("run_fourier")
    ny, nx = shape
    fake = g[f"geo_{ny}_{nx}_fake"]
    qa, qb = fourier.first_quarters(fake)
    full = fourier.assemble_mask((qa > 0.9).astype(np.float64), (qb > 0.8).astype(np.float64), ny, nx)
    eq(full, g[f"geo_{ny}_{nx}_full"])


def test_fft_vs_definition():
    rng = np.random.default_rng(2)
    a = rng.normal(0, 1, (13, 10))
    np.testing.assert_allclose(fourier.dft2_direct(a), np.fft.fft2(a), atol=1e-10)
    f, fs = fourier.fourier_initial(a.astype(np.float32))
    np.testing.assert_allclose(np.fft.ifftshift(fs), fourier.dft2_direct(a), atol=1e-4)


# ---- new stages (parity unpinned): two independent algorithms must agree ------------------------
def test_median_two_ways():
    rng = np.random.default_rng(4)
    a = rng.normal(100, 5, (40, 57)).astype(np.float32)
    for ws in (3, 5):
        m = stencils.median(a, ws)
        h = ws // 2
        ref = ndimage.median_filter(a, size=ws)
        np.testing.assert_array_equal(m[h:-h, h:-h], ref[h:-h, h:-h])
        np.testing.assert_array_equal(m[:h], a[:h])
        eq(clib.median(a, ws, False), m)
    b = a.copy()
    b[rng.random(a.shape) < 0.2] = np.nan
    b[10:16, 10:16] = np.nan
    for ws, circ in ((3, False), (5, False), (5, True), (3, True)):
        eq(clib.median(b, ws, circ), stencils.median(b, ws, circ))
    # nanmedian definition on one cell
    w = windows.window_at(b, 5, 20, 20, circular=True)
    with np.errstate(all="ignore"):
        assert stencils.median(b, 5, True)[20, 20] == np.nanmedian(w)


def test_sinkfill_two_ways_and_d8():
    rng = np.random.default_rng(6)
    z = (rng.normal(0, 1, (60, 71)).cumsum(0).cumsum(1) / 10).astype(np.float32)
    z[20:24, 30:33] = np.nan
    z[40, 40] = -50.0                                      # a pit
    it, sweeps = hydrology.sinkfill_iterative(z)
    pf = hydrology.sinkfill(z)
    eq(pf, it)
    assert sweeps > 3 and (pf[~np.isnan(z)] >= z[~np.isnan(z)]).all() and pf[40, 40] > z[40, 40]
    # fixed point: idempotent
    eq(hydrology.sinkfill(pf), pf)
    d = hydrology.d8(pf)
    eq(clib.d8(pf), d)
    assert d[0].sum() == 0 and d[:, 0].sum() == 0 and set(np.unique(d)) <= {0, 1, 2, 4, 8, 16, 32, 64, 128}
    # every directed cell points to a strictly lower neighbour
    jj, ii = np.nonzero(d)
    k = np.log2(d[jj, ii]).astype(int)
    dy = np.array(hydrology.D8_DY)[k]; dx = np.array(hydrology.D8_DX)[k]
    assert (pf[jj + dy, ii + dx] < pf[jj, ii]).all()
