"""GPU: GeoTIFF files straight to device rasters and through the chain."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from hydrodem_b200 import device as dev, geotiff
    from hydrodem_b200.pipeline import ConditioningChain
    from hydrodem_b200.synth import SynthScene


@pytest.mark.parametrize("dtype,chunk", [(np.float32, 64 << 20), (np.float32, 5000), (np.int16, 3000), (np.uint8, 1000)])
def test_read_to_device_equals_read_array(tmp_path, dtype, chunk):
    """Chunked, double-buffered ingest (pinned staging + copy stream) delivers the file's pixels; one chunk, many chunks,
    a last chunk that is shorter."""
    rng = np.random.default_rng(7)
    a = (rng.standard_normal((211, 333)) * 300).astype(dtype)
    path = tmp_path / "a.tif"
    geotiff.write_geotiff(path, a, strip_bytes=2000)
    host = geotiff.read_array(path)                                  # pinned
    np.testing.assert_array_equal(host, a)
    r = geotiff.read_to_device(path, chunk_bytes=chunk)
    assert r.shape == a.shape
    got = dev.download(r)
    assert got.dtype == a.dtype
    np.testing.assert_array_equal(got, a)


@pytest.mark.parametrize("dtype", [np.float32, np.int16])
def test_process_geotiffs_matches_the_array_api(tmp_path, dtype):
    """Files in, file out: float32 or int16 SRTM / HydroSHEDS rasters and a uint8 groves raster, as GDAL delivers them; the
    written DEM is the float32 cast of the chain's float64 result and carries the SRTM raster's georeference."""
    sc = SynthScene(300, 340, 5)
    srtm = (np.round(sc.srtm()) if dtype == np.int16 else sc.srtm()).astype(dtype)
    groves = sc.groves().astype(np.uint8)
    hsheds = sc.hsheds().astype(dtype)
    scale = np.array([0.000833, 0.000833, 0.0])
    tie = np.array([0.0, 0.0, 0.0, -60.5, -31.25, 0.0])
    geo = {33550: (12, 3, scale.astype("<f8").tobytes()), 33922: (12, 6, tie.astype("<f8").tobytes())}
    paths = {k: str(tmp_path / f"{k}.tif") for k in ("srtm", "groves", "hsheds", "final")}
    geotiff.write_geotiff(paths["srtm"], srtm, geo_tags=geo, nodata=-32768)
    geotiff.write_geotiff(paths["groves"], groves)
    geotiff.write_geotiff(paths["hsheds"], hsheds, nodata=-32768)
    chain = ConditioningChain()
    out = geotiff.process_geotiffs(paths["srtm"], paths["groves"], paths["hsheds"], paths["final"], chain=chain)
    want = chain.apply(srtm, groves, hsheds.copy())
    np.testing.assert_array_equal(out["final"], want.final)
    np.testing.assert_array_equal(out["filled"], want.filled)
    np.testing.assert_array_equal(out["d8"], want.d8)
    info = geotiff.read_info(paths["final"])
    assert info.dtype == np.float32 and info.geo_tags() == geo
    np.testing.assert_allclose(info.geotransform(), (-60.5, 0.000833, 0.0, -31.25, 0.0, -0.000833))
    np.testing.assert_array_equal(geotiff.read_array(paths["final"]), want.final.astype(np.float32))
