"""Generate the committed golden fixtures (run in the BUILD container only).

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/gen_golden.py

Two kinds of fixture are written next to this script:

ref_*.npz  -- the reference's own stored goldens, converted 1:1 from the
              GeoTIFFs in /root/reference/cguerrero/tests/resources/
              tests_expected.zip (read with PIL; GDAL is absent).  They pin
              chains G1-G4 of SURVEY.md section 4.3.
run_*.npz  -- outputs of the UNMODIFIED reference classes, imported from
              /root/reference/cguerrero/hydrodem, on small seeded synthetic
              inputs.  Inputs are stored with the outputs so the fixtures are
              self-contained on the GPU box, where /root/reference does not
              exist.

The reference tree is read-only: zips are extracted under /tmp and Python is
told not to write bytecode there.
"""
import os
import sys
import zipfile

import numpy as np

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference/cguerrero"
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REF, "hydrodem"))        # flat imports of the reference

from PIL import Image                                      # noqa: E402
from hydrodem_b200.synth import SynthScene                 # noqa: E402
import sliding_window as ref_sw                            # noqa: E402
from filters import custom_filters as ref_cf               # noqa: E402
from filters import extension_filters as ref_ef            # noqa: E402
from filters import simple_filters as ref_sf               # noqa: E402


def save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrays)
    print(f"{name}.npz  {os.path.getsize(path) / 1024:.0f} KiB  " +
          " ".join(f"{k}{v.shape}{v.dtype}" for k, v in arrays.items()))


def convert_reference_goldens():
    tmp = "/tmp/hydrodem_expected"
    zipfile.ZipFile(os.path.join(REF, "tests/resources/tests_expected.zip")).extractall(tmp)
    rd = lambda n: np.array(Image.open(os.path.join(tmp, "expected", n + ".tif")))
    save("ref_lagoons",                                     # G1 + G2
         hsheds_nan_values_expected=rd("hsheds_nan_values_expected"),
         hsheds_majority_11_expected=rd("hsheds_majority_11_expected"),
         lagoons_expected=rd("lagoons_expected"))
    save("ref_mask_fourier",                                # G3 + G4
         isolated_filter_expected=rd("isolated_filter_expected"),
         mask_fourier_expected=rd("mask_fourier_expected"),
         filtered_blank_expected_2=rd("filtered_blank_expected_2"))
    save("ref_tiles",                                       # bundled SRTM tiles (C1 inputs)
         srtm_corrected=rd("srtm_corrected"),
         srtm_uncompress_expected=rd("srtm_uncompress_expected"))


def window_kats():
    """Known-answer windows from the reference iterators (A0)."""
    out = {}
    grid5 = np.arange(25).reshape(5, 5)
    grid7 = np.arange(63).reshape(7, 9)

    def dump(tag, it):
        wins, cen = zip(*[(w, c) for w, c in it])
        out[tag + "_w"] = np.stack(wins)
        out[tag + "_c"] = np.array(cen)

    dump("plain3", ref_sw.SlidingWindow(grid5, 3))
    dump("circ5", ref_sw.CircularWindow(grid7, 5))
    dump("nocenter3", ref_sw.NoCenterWindow(grid5, 3))
    dump("inner5_3", ref_sw.InnerWindow(grid7, 5, 3))
    dump("ignore3", ref_sw.SlidingIgnoreBorder(grid5, 3))
    dump("combo5_3", ref_sw.IgnoreBorderInnerSliding(grid7, window_size=5, inner_size=3))
    ones = (np.arange(25).reshape(5, 5) % 3 == 1) * 1.5
    dump("gate3", ref_sw.SlidingWindow(ones, 3, iter_over_ones=True))
    out["grid5"], out["grid7"], out["gate_grid"] = grid5, grid7, ones
    save("run_windows", **out)


def run_stencils():
    sc = SynthScene(150, 170, 7)
    hs = sc.hsheds()
    hs[40:46, 60:64] = -32768.0                 # a void block: some voids have <8 / 0 valid neighbours
    hs[41, 61] = np.nan
    hs[100, 5] = -1.0
    srtm = sc.srtm()
    rng = np.random.default_rng(11)

    # A3 CorrectNANValues (in place)
    nanfix_in = hs.copy()
    with np.errstate(all="ignore"):
        nanfix_out = ref_cf.CorrectNANValues().apply(nanfix_in.copy())
    # float data (non-integer) version pins the float32 summation order
    nf2 = (srtm + rng.uniform(0, 1, srtm.shape).astype(np.float32)).astype(np.float32)
    nf2[rng.random(srtm.shape) < 0.02] = -5.0
    nf2[rng.random(srtm.shape) < 0.01] = np.nan
    with np.errstate(all="ignore"):
        nf2_out = ref_cf.CorrectNANValues().apply(nf2.copy())
    # A1 majority (11) on plateaus + signed zeros + NaN
    maj_in = nanfix_out.copy()
    maj_in[20:34, 20:40] = 0.0
    maj_in[22:30:2, 22:38:3] = -0.0
    maj_in[70, 70:90] = np.nan
    maj_out = ref_cf.MajorityFilter(window_size=11).apply(maj_in)
    maj5_out = ref_cf.MajorityFilter(window_size=5).apply(maj_in)
    # A2 expand 3/7/13 on a sparse mask with NaN and negatives
    ex_in = (rng.random((150, 170)) < 0.004).astype(np.float64)
    ex_in[0, :] = 1.0
    ex_in[75, 80] = np.nan
    ex_in[30, 30] = -2.0
    ex = {f"expand{w}_out": ref_cf.ExpandFilter(window_size=w).apply(ex_in) for w in (3, 7, 13)}
    # A4 isolated points (in place), values in [1,2) pass the int() gate
    iso_in = (rng.random((150, 170)) < 0.03).astype(np.float64)
    iso_in[10, 10] = 1.7
    iso_in[12, 12] = 2.0
    iso_in[14, 14] = 0.5
    iso_out = ref_cf.IsolatedPoints(window_size=3).apply(iso_in.copy())
    # A5 quadratic (15) f32 and f64 input, plus one groves iteration and the 3-iteration wrapper
    quad32 = ref_cf.QuadraticFilter(window_size=15).apply(srtm)
    srtm64 = srtm.astype(np.float64) + 1e-9
    quad64 = ref_cf.QuadraticFilter(window_size=15).apply(srtm64)
    groves = ref_ef.BinaryClosing(structure=np.ones((3, 3))).apply(sc.groves())
    g1 = ref_cf.GrovesCorrection(groves).apply(srtm64)
    g3 = ref_cf.GrovesCorrectionsIter(groves, iterations=3).apply(srtm64)
    save("run_stencils", nanfix_in=nanfix_in, nanfix_out=nanfix_out, nanfix_f_in=nf2, nanfix_f_out=nf2_out,
         maj_in=maj_in, maj11_out=maj_out, maj5_out=maj5_out, expand_in=ex_in, **ex,
         iso_in=iso_in, iso_out=iso_out, srtm=srtm, quad32=quad32, srtm64=srtm64, quad64=quad64,
         groves_raw=sc.groves(), groves_closed=groves, groves1=g1, groves3=g3)


def run_lagoons_and_final():
    sc = SynthScene(150, 170, 8)
    hs = sc.hsheds()
    lag = ref_cf.LagoonsDetection()
    with np.errstate(all="ignore"):
        ret = lag.apply(hs.copy())
    er = ref_ef.BinaryErosion(iterations=2).apply(lag.results["MajorityFilter"])
    cl_cross = ref_ef.BinaryClosing().apply(sc.groves())
    cl_full = ref_ef.BinaryClosing(structure=np.ones((3, 3))).apply(sc.groves())
    gd = ref_ef.GreyDilation(size=(7, 7)).apply(lag.results["MajorityFilter"])
    # PostProcessingFinal on a float64 raster
    dem = sc.srtm().astype(np.float64) * 1.0000001
    post = ref_cf.PostProcessingFinal().apply(dem)
    conv = ref_ef.Convolve().apply(dem)
    save("run_lagoons", hsheds=hs, nanfixed=lag.results["CorrectNANValues"], majority=lag.results["MajorityFilter"],
         tidying=lag.results["TidyingLagoons"], mask=ret, erosion2=er, groves_raw=sc.groves(),
         closing_cross=cl_cross, closing_full=cl_full, greydil7=gd, dem64=dem, conv3=conv, post=post)


def run_fourier():
    sc = SynthScene(150, 170, 9)                # quarters 65x75 >= 55
    srtm = sc.srtm()
    daf = ref_cf.DetectApplyFourier()
    corrected = daf.apply(srtm)
    fabs = daf.fft_transform_abs
    fshift = daf.initial.fourier_shift
    pq = ref_cf.FourierProcessQuarters(fabs)
    q1, q2 = pq._get_firsts_quarters()
    b_mask, b_mod = ref_cf.BlanksFourier(window_size=55).apply(q1)
    det = ref_cf.DetectBlanksFourier().apply(q1)
    m1 = ref_cf.MaskFourier().apply(q1)
    m2 = ref_cf.MaskFourier().apply(q2)
    mask = pq.apply(None)
    # odd x odd and even x even geometry of the mask assembly, with synthetic quarter masks
    geo = {}
    for (ny, nx) in ((131, 140), (140, 131), (131, 133), (134, 136)):
        fake = np.random.default_rng(ny * 1000 + nx).random((ny, nx)).astype(np.float32)
        p = ref_cf.FourierProcessQuarters(fake)
        qa, qb = p._get_firsts_quarters()
        ma = (qa > 0.9).astype(np.float64)
        mb = (qb > 0.8).astype(np.float64)
        full = p._fill_complete_mask(p._getting_reversed_masks(p._fill_complete_quarters((ma, mb))))
        geo[f"geo_{ny}_{nx}_fake"] = fake
        geo[f"geo_{ny}_{nx}_full"] = full
    save("run_fourier", srtm=srtm, fabs=fabs, fshift=fshift, q1=q1, q2=q2, blanks_mask=b_mask, blanks_mod=b_mod,
         detect=det, mask_q1=m1, mask_q2=m2, mask=mask, corrected=corrected, **geo)


def run_simple():
    rng = np.random.default_rng(3)
    a = rng.normal(0, 2, (20, 30)).astype(np.float32)
    b = rng.normal(0, 2, (20, 30))
    out = dict(a=a, b=b,
               lower=ref_sf.LowerThan(value=0.0).apply(a), greater=ref_sf.GreaterThan(value=1.5).apply(a),
               b2i=ref_sf.BooleanToInteger().apply(a > 0), prod=ref_sf.ProductFilter(factor=b).apply(a),
               prod_s=ref_sf.ProductFilter(factor=3).apply(a), add=ref_sf.AdditionFilter(addend=b).apply(a),
               sub=ref_sf.SubtractionFilter(minuend=1).apply(a), sub_a=ref_sf.SubtractionFilter(minuend=b).apply(a),
               absv=ref_ef.AbsoluteValues().apply(a), around=ref_ef.Around().apply(a * 1.25),
               xor=ref_ef.BitwiseXOR(operand=(a > 0)).apply(b > 0))
    save("run_simple", **out)


def synth_rivers(ny, nx, seed, n_rivers=6):
    """Rasterised river polylines: random walks drifting across the tile, value 1 on a float32 raster (what
    rasterize_rivers leaves, utils_dem.py); a few thick spots and a river that runs along the frame."""
    rng = np.random.default_rng(seed)
    r = np.zeros((ny, nx), dtype=np.float32)
    for k in range(n_rivers):
        y, x = int(rng.integers(0, ny)), 0
        while x < nx:
            r[min(max(y, 0), ny - 1), x] = 1
            step = rng.integers(-1, 2)
            if step and rng.random() < 0.5:
                y += int(step)
                r[min(max(y, 0), ny - 1), x] = 1
            x += 1
    r[ny // 3:ny // 3 + 4, nx // 2:nx // 2 + 5] = 1
    r[0, :nx // 3] = 1
    r[:, nx - 1] = 1
    return r


def run_rivers():
    """RouteRivers / ProcessRivers / ClipLagoonsRivers (custom_filters.py:128-199, :770-831) on seeded inputs: the
    integer-valued HydroSHEDS raster gives plateaus (several window cells tie for the minimum), the float one does not."""
    out = {}
    for tag, (ny, nx, seed) in {"a": (150, 190, 5), "b": (97, 260, 9)}.items():
        sc = SynthScene(ny, nx, 400 + seed)
        hs = sc.hsheds()
        hs[hs < 0] = 90.0                                             # (voids are fixed before rivers are routed)
        dem_f = sc.srtm()
        rivers = synth_rivers(ny, nx, seed)
        mask = ref_cf.ExpandFilter(window_size=3).apply(ref_cf.MaskPositives().apply(rivers))
        out[f"{tag}_hsheds"], out[f"{tag}_srtm"], out[f"{tag}_rivers"], out[f"{tag}_mask"] = hs, dem_f, rivers, mask
        out[f"{tag}_routed_int"] = ref_cf.RouteRivers(window_size=3, dem=hs).apply(mask)
        out[f"{tag}_routed_float"] = ref_cf.RouteRivers(window_size=3, dem=dem_f).apply(mask)
        routed = ref_cf.ProcessRivers(hs).apply(rivers)
        out[f"{tag}_process"] = routed
        lag = ref_cf.LagoonsDetection()
        lag.apply(sc.hsheds())
        out[f"{tag}_mask_lagoons"] = lag.mask_lagoons
        out[f"{tag}_clip"] = ref_cf.ClipLagoonsRivers(lag.mask_lagoons, routed).apply(routed)
    save("run_rivers", **out)


def run_int16():
    """The reference's PRODUCTION dtype: gdal ReadAsArray hands int16 rasters to LagoonsDetection (image_hsheds.py:133-135)
    and DetectApplyFourier (image_srtm.py:125-126).  CorrectNANValues then writes its float32 means into an int16 array
    (truncation on assignment) and scipy.fftpack promotes the integer raster to float64 / complex128."""
    sc = SynthScene(230, 260, 77)
    hs = sc.hsheds().astype(np.int16)
    hs[100, 100:104] = -32768                                         # a run of voids: means that are not integers
    hs[101, 101] = -32768
    out = {"hsheds_i16": hs.copy()}
    work = hs.copy()
    lag = ref_cf.LagoonsDetection()
    ret = lag.apply(work)
    out["lag_return"] = ret
    for k in ("CorrectNANValues", "MajorityFilter", "TidyingLagoons", "MaskPositives"):
        out["lag_" + k] = np.asarray(lag.results[k])
    assert lag.results["CorrectNANValues"].dtype == np.int16 and lag.results["CorrectNANValues"] is work
    srtm = np.round(sc.srtm()).astype(np.int16)
    out["srtm_i16"] = srtm
    daf = ref_cf.DetectApplyFourier()
    out["daf_i16"] = daf.apply(srtm)
    save("run_int16", **out)


def run_geotiff():
    """The reference's one GDAL-written raster (resources/images/final_dem.tif, an array2raster output): its georeference
    tags and a crop of its pixels.  The full file is checked here, at generation time: hydrodem_b200.geotiff reads
    exactly what libtiff (Pillow) reads."""
    from PIL import Image
    from hydrodem_b200 import geotiff
    path = os.path.join(REF, "resources", "images", "final_dem.tif")
    pil = Image.open(path)
    want = np.array(pil)
    info = geotiff.read_info(path)
    got = geotiff.read_array(path, pinned=False)
    assert got.dtype == np.float32 and got.shape == want.shape == (519, 508) and np.array_equal(got, want, equal_nan=True)
    assert info.compression == 1 and not info.tiled and geotiff._contiguous(info)
    tags = info.geo_tags()
    for t, (typ, cnt, raw) in tags.items():                           # libtiff decodes the same values
        v = pil.tag_v2[t]
        if typ == 12:
            assert np.array_equal(np.asarray(v, dtype=np.float64), np.frombuffer(raw, dtype="<f8"))
        elif typ == 3:
            assert np.array_equal(np.asarray(v), np.frombuffer(raw, dtype="<u2"))
    sx, sy = pil.tag_v2[33550][:2]
    tie = pil.tag_v2[33922]
    gt = (tie[3] - tie[0] * sx, sx, 0.0, tie[4] + tie[1] * sy, 0.0, -sy)   # what GetGeoTransform returns for these tags
    assert tuple(info.geotransform()) == gt
    ids = sorted(tags)
    save("run_geotiff", array=want[200:264, 100:180].copy(), full_shape=np.array(want.shape),
         full_sum=np.array(np.nansum(want.astype(np.float64))), tag_ids=np.array(ids),
         tag_types=np.array([tags[t][0] for t in ids]), tag_counts=np.array([tags[t][1] for t in ids]),
         tag_raw_len=np.array([len(tags[t][2]) for t in ids]),
         tag_raw=np.frombuffer(b"".join(tags[t][2] for t in ids), dtype=np.uint8), geotransform=np.array(gt))


def run_stats():
    """The reference's Stats class (stats.py) itself, driven through stub ``gdal`` / ``config_loader`` modules (GDAL is not
    installed; the class only uses gdal.Open(name).ReadAsArray() and Config.simulation(key)).  Water masks of the sample
    types GDAL delivers (uint8, int16, float32; one with values outside {0, 1} and a NaN) x three simulated rasters."""
    import types
    rasters = {}
    gdal = types.ModuleType("gdal")

    class _DS:
        def __init__(self, name):
            self.name = name

        def ReadAsArray(self):
            return rasters[self.name]

    gdal.Open = _DS
    cfg = types.ModuleType("config_loader")

    class Config:
        @staticmethod
        def simulation(key):
            return {"NDWI_IMAGE": "ndwi", "SIM": "sim_{}"}[key]

    cfg.Config = Config
    saved = {k: sys.modules.get(k) for k in ("gdal", "config_loader")}
    sys.modules["gdal"], sys.modules["config_loader"] = gdal, cfg
    try:
        import importlib
        ref_stats = importlib.import_module("stats")
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    rng = np.random.default_rng(77)
    shape = (97, 131)
    water = rng.random(shape) < 0.3
    out = {}
    sims = []
    for d in range(3):
        depth = np.where(water ^ (rng.random(shape) < 0.15), rng.random(shape) * 2, 0.0).astype(np.float32)
        depth[rng.random(shape) < 0.02] = np.float32(0.0001)            # exactly the threshold: not wet
        depth[rng.random(shape) < 0.02] = np.float32(0.00011)
        sims.append(depth)
        out[f"sim_{d + 1}"] = depth
    sims[2] = sims[2].astype(np.float64)
    out["sim_3"] = sims[2]
    masks = {"u8": water.astype(np.uint8), "i16": water.astype(np.int16), "f32": water.astype(np.float32)}
    odd = water.astype(np.float32)
    odd[5, 7] = np.nan
    odd[9, 3:9] = 2.0
    odd[11, 2:5] = -1.0
    masks["f32odd"] = odd
    odd8 = water.astype(np.uint8)
    odd8[9, 3:9] = 2
    masks["u8odd"] = odd8
    keys = None
    for name, ndwi in masks.items():
        rasters.clear()
        rasters["ndwi"] = ndwi
        for d in range(3):
            rasters[f"sim_{d + 1}"] = sims[d]
        with np.errstate(all="ignore"):
            st = ref_stats.Stats("SIM", 4, "x")
            got = st.get_stats()
            counts = []
            for d in range(3):
                st._set_values(d + 1)
                counts.append([st.values_file[k] for k in ("TP", "FN", "P", "FP", "TN", "N")])
        keys = [k for k in got[0] if k != "day"]
        out[f"ndwi_{name}"] = ndwi
        out[f"counts_{name}"] = np.array(counts, dtype=np.int64)
        out[f"totals_{name}"] = np.array([st.total_positives, st.total_negatives], dtype=np.int64)
        out[f"scores_{name}"] = np.array([[g[k] for k in keys] for g in got], dtype=np.float64)
    out["score_keys"] = np.array(keys)
    save("run_stats", **out)


if __name__ == "__main__":
    if "--only-stats" in sys.argv:
        run_stats()
        sys.exit(0)
    if "--only-geotiff" in sys.argv:
        run_geotiff()
        sys.exit(0)
    if "--only-int16" in sys.argv:
        run_int16()
        sys.exit(0)
    if "--only-rivers" in sys.argv:
        run_rivers()
        sys.exit(0)
    convert_reference_goldens()
    window_kats()
    run_stencils()
    run_lagoons_and_final()
    run_fourier()
    run_simple()
    run_rivers()
    run_int16()
    run_geotiff()
    run_stats()
