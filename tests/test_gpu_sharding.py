"""GPU: row-band sharding emulated with ThreadComm (ranks = threads on one device): banded results must be
bit-identical to the single-GPU run / the oracle."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from hydrodem_b200 import device as dev, sharding
    from hydrodem_b200.filters import custom_filters as cf, extension_filters as ef
    from hydrodem_b200.synth import SynthScene
    from oracle import hydrology, stencils


@pytest.mark.parametrize("world", [2, 3])
def test_banded_stencils_match_single_gpu(world):
    sc = SynthScene(301, 333, 51)
    hs, srtm = sc.hsheds(), sc.srtm()
    want_maj = cf.MajorityFilter(window_size=11).apply(hs)
    want_quad = cf.QuadraticFilter(window_size=15).apply(srtm)
    want_er = ef.BinaryErosion(iterations=2).apply(want_maj)

    def fn(comm):
        band = sharding.Band(comm, *hs.shape)
        maj = band.apply(cf.MajorityFilter(window_size=11), dev.upload(band.take(hs)), 5)
        quad = band.apply(cf.QuadraticFilter(window_size=15), dev.upload(band.take(srtm)), 7)
        er = band.apply(ef.BinaryErosion(iterations=2), maj, 2)
        return dev.download(maj), dev.download(quad), dev.download(er)

    res = sharding.ThreadComm.run(world, fn)
    np.testing.assert_array_equal(np.concatenate([r[0] for r in res]), want_maj)
    # tolerance-class stage: float32 partial sums are re-centred per tile, and band tiles are not aligned with the
    # single-GPU tiles -> agreement to rounding, not to the bit
    np.testing.assert_allclose(np.concatenate([r[1] for r in res]), want_quad, rtol=1e-6)
    np.testing.assert_array_equal(np.concatenate([r[2] for r in res]), want_er)


@pytest.mark.parametrize("world,ny", [(2, 400), (4, 400), (2, 512), (3, 386)])
def test_banded_sinkfill_matches_oracle(world, ny):
    # ny = 512 / world 2: the extended band has 257 rows, so the halo row sits alone in its tile row
    sc = SynthScene(ny, 390, 52)
    z = np.round(sc.srtm())
    z[100:103, 50:60] = np.nan
    z[200:260, 100:180] -= 9                                   # a pan that spans band boundaries
    want = hydrology.sinkfill(z)

    def fn(comm):
        band = sharding.Band(comm, *z.shape)
        w, d8 = band.sinkfill(dev.upload(band.take(z)))
        return dev.download(w), dev.download(d8), band.fill_rounds

    res = sharding.ThreadComm.run(world, fn)
    np.testing.assert_array_equal(np.concatenate([r[0] for r in res]), want)
    np.testing.assert_array_equal(np.concatenate([r[1] for r in res]), hydrology.d8(want))
    assert res[0][2] >= 2


@pytest.mark.parametrize("world,shape", [(2, (150, 201)), (3, (129, 256)), (4, (257, 131))])
def test_distributed_fft2_matches_single_gpu(world, shape):
    """Band.fft2 / Band.ifft2: local row transforms, ONE all-to-all, local column transforms.  The transposed band
    layout reassembles to the full spectrum; tolerance as for the single-GPU transform (1e-5 RMS + 4 ulp of the DC
    bin); the round trip returns the input."""
    ny, nx = shape
    x = SynthScene(ny, nx, 61).srtm()
    want = np.fft.fft2(x.astype(np.float64))

    def fn(comm):
        band = sharding.Band(comm, ny, nx)
        spec_t = band.fft2(dev.upload(band.take(x)))
        back = band.ifft2(spec_t)
        return dev.download(spec_t), dev.download(back)

    res = sharding.ThreadComm.run(world, fn)
    got = np.concatenate([r[0] for r in res], axis=0).T              # (nx, ny) transposed bands -> (ny, nx)
    assert got.shape == want.shape and got.dtype == np.complex64
    tol = 1e-5 * np.sqrt(np.mean(np.abs(want) ** 2)) + 4 * np.spacing(np.float32(np.abs(want).max()))
    assert np.abs(got - want).max() <= tol
    back = np.concatenate([r[1] for r in res], axis=0)
    assert back.shape == x.shape
    np.testing.assert_allclose(back.real, x, rtol=0, atol=2e-5 * np.abs(x).max())
    assert np.abs(back.imag).max() <= 2e-5 * np.abs(x).max()


@pytest.mark.parametrize("world", [2, 3])
def test_banded_detect_apply_fourier_matches_single_gpu(world):
    """The stripe-removal stage on row bands (distributed transforms, replicated peak detector) against the
    single-GPU stage: same mask by construction, DEM within the stage's 1e-5 relative tolerance."""
    sc = SynthScene(301, 333, 71)
    srtm = sc.srtm()
    want = cf.DetectApplyFourier().apply(srtm)

    def fn(comm):
        band = sharding.Band(comm, *srtm.shape)
        return dev.download(band.detect_apply_fourier(dev.upload(band.take(srtm))))

    res = sharding.ThreadComm.run(world, fn)
    got = np.concatenate(res)
    assert got.dtype == np.float64 and got.shape == want.shape
    np.testing.assert_allclose(got, want, rtol=1e-5)


@pytest.mark.parametrize("world", [2, 3])
def test_banded_chain_matches_single_gpu(world):
    """The whole conditioning chain on row bands against the single-GPU chain: exact stages identical, the DEM before
    rounding within 1e-5, rounded DEM equal except where the mean sits on a half-integer, hydrology consistent with
    the banded final DEM (the fill's fixed point is unique)."""
    from hydrodem_b200.pipeline import ConditioningChain
    sc = SynthScene(420, 333, 81)
    srtm, groves, hsheds = sc.srtm(), sc.groves(), sc.hsheds()
    ref = ConditioningChain(keep_intermediates=True).apply(srtm, groves, hsheds.copy())
    want_complete, want_final = ref.dem_complete, ref.final

    def fn(comm):
        band = sharding.Band(comm, *srtm.shape)
        out = band.conditioning_chain(dev.upload(band.take(srtm)), dev.upload(band.take(groves)),
                                      dev.upload(band.take(hsheds)))
        return {k: dev.download(v) for k, v in out.items()}

    res = sharding.ThreadComm.run(world, fn)
    got = {k: np.concatenate([r[k] for r in res]) for k in res[0]}
    np.testing.assert_allclose(got["dem_complete"], want_complete, rtol=1e-5)
    flips = got["final"] != want_final
    from scipy import ndimage
    mean = ndimage.convolve(want_complete, np.ones((3, 3)), mode="reflect") / 9.0
    frac = np.abs(mean - np.floor(mean) - 0.5)
    assert flips.mean() < 1e-3 and (frac[flips] < 5e-3).all()          # only where the mean sits on a half-integer
    np.testing.assert_array_equal(got["filled"], hydrology.sinkfill(got["final"]))
    np.testing.assert_array_equal(got["d8"], hydrology.d8(got["filled"]))
