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
