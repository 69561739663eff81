"""RouteRivers / ProcessRivers / ClipLagoonsRivers (reference custom_filters.py:128-199, :770-831): the C oracle against
reference-run fixtures on CPU, the CUDA order-preserving wavefront against both on the GPU, and the drop-in import
lines of the reference's own callers."""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import load_golden

from oracle import clib


def test_oracle_route_rivers_matches_reference_fixtures():
    g = load_golden("run_rivers")
    for tag in ("a", "b"):
        for dem, want in ((g[f"{tag}_hsheds"], g[f"{tag}_routed_int"]), (g[f"{tag}_srtm"], g[f"{tag}_routed_float"])):
            got = clib.route_rivers(g[f"{tag}_mask"], dem)
            assert got.dtype == np.float64
            np.testing.assert_array_equal(got, want)
    assert g["a_routed_int"].sum() > g["a_routed_float"].sum() > 100      # plateaus: several cells tie for the minimum


def test_reference_import_lines_resolve_after_install():
    """image_hsheds.py:6-7, image_srtm.py:7-8 and hydro_dem_process.py:17-19 import these names from
    filters.custom_filters: after install_as_reference_filters() they resolve to this package."""
    import hydrodem_b200
    saved = {k: sys.modules.get(k) for k in list(sys.modules) if k == "filters" or k.startswith("filters.") or k == "exceptions"}
    try:
        hydrodem_b200.install_as_reference_filters()
        ns = {}
        exec("from filters.custom_filters import (LagoonsDetection, ClipLagoonsRivers,\n"
             "                                    ProcessRivers)", ns)                      # image_hsheds.py:6-7
        exec("from filters.custom_filters import (DetectApplyFourier, BinaryClosing,\n"
             "                                    GrovesCorrectionsIter)", ns)              # image_srtm.py:7-8
        exec("from filters.custom_filters import (SubtractionFilter, ProductFilter,\n"
             "                                    AdditionFilter, PostProcessingFinal)", ns)  # hydro_dem_process.py:17-19
        for name in ("LagoonsDetection", "ClipLagoonsRivers", "ProcessRivers", "DetectApplyFourier", "BinaryClosing",
                     "GrovesCorrectionsIter", "SubtractionFilter", "ProductFilter", "AdditionFilter", "PostProcessingFinal"):
            assert ns[name].__module__.startswith("hydrodem_b200.filters"), name
    finally:
        for k in [k for k in sys.modules if k == "filters" or k.startswith("filters.") or k == "exceptions"]:
            del sys.modules[k]
        for k, v in saved.items():
            if v is not None:
                sys.modules[k] = v


@pytest.mark.gpu
def test_route_rivers_fixtures_gpu():
    from hydrodem_b200.filters import custom_filters as cf
    g = load_golden("run_rivers")
    for tag in ("a", "b"):
        for dem, want in ((g[f"{tag}_hsheds"], g[f"{tag}_routed_int"]), (g[f"{tag}_srtm"], g[f"{tag}_routed_float"])):
            dem_before = dem.copy()
            got = cf.RouteRivers(window_size=3, dem=dem).apply(g[f"{tag}_mask"])
            assert got.dtype == np.float64
            np.testing.assert_array_equal(got, want)
            np.testing.assert_array_equal(dem, dem_before)                 # the ctor deep-copies the DEM (:163)
        routed = cf.ProcessRivers(g[f"{tag}_hsheds"]).apply(g[f"{tag}_rivers"])
        assert routed.dtype == np.bool_
        np.testing.assert_array_equal(routed, g[f"{tag}_process"])
        clip = cf.ClipLagoonsRivers(g[f"{tag}_mask_lagoons"], routed).apply(routed)
        assert clip.dtype == g[f"{tag}_clip"].dtype
        np.testing.assert_array_equal(clip, g[f"{tag}_clip"])


@pytest.mark.gpu
@pytest.mark.parametrize("shape,density,rows_per_launch", [((700, 901), 0.07, 0), ((1203, 640), 0.5, 0), ((400, 333), 1.0, 0),
                                                           ((2500, 300), 0.2, 512)])
def test_route_rivers_order_vs_oracle(shape, density, rows_per_launch, monkeypatch):
    """Dense masks make neighbouring visits interact everywhere: the wavefront must reproduce the raster order exactly
    (integer DEM = many ties; a NaN patch; rows_per_launch forces several stripes)."""
    from hydrodem_b200.filters import custom_filters as cf
    if rows_per_launch:
        monkeypatch.setenv("HD_RIVERS_ROWS_PER_LAUNCH", str(rows_per_launch))
    rng = np.random.default_rng(shape[0])
    dem = np.round(rng.normal(100, 3, shape)).astype(np.float32)
    dem[50:53, 60:70] = np.nan
    mask = (rng.random(shape) < density).astype(np.float32)
    mask[10, 10] = 1.7                                                     # int(1.7) == 1: visited
    mask[11, 40] = 2.0                                                     # not visited
    want = clib.route_rivers(mask, dem)
    got = cf.RouteRivers(window_size=3, dem=dem).apply(mask)
    np.testing.assert_array_equal(got, want)
    assert want.sum() > 0


@pytest.mark.gpu
def test_route_rivers_errors():
    from hydrodem_b200.exceptions import DeviceError, NumpyArrayExpectedError, WindowSizeEvenError, WindowSizeHighError
    from hydrodem_b200.filters import custom_filters as cf
    a = np.zeros((20, 20), dtype=np.float32)
    with pytest.raises(NumpyArrayExpectedError):
        cf.RouteRivers(window_size=3, dem=a).apply([1, 2])
    with pytest.raises(WindowSizeHighError):
        cf.RouteRivers(window_size=21, dem=a).apply(a)
    with pytest.raises(WindowSizeEvenError):
        cf.RouteRivers(window_size=4, dem=a).apply(a)
    with pytest.raises(DeviceError):
        cf.RouteRivers(window_size=5, dem=a).apply(a)                      # only 3 (ProcessRivers) on the device
