"""CPU: the GeoTIFF reader / writer behind the chain's callers against Pillow (libtiff) and the
georeference tags of the reference's own GDAL-written raster (tests/golden/run_geotiff.npz)."""
import struct
import zlib

import numpy as np
import pytest
from PIL import Image

from conftest import load_golden
from hydrodem_b200 import geotiff


@pytest.fixture(autouse=True, scope="module")
def _no_leftover_io_threads():
    yield
    geotiff.shutdown_pool()              # later test modules fork() worker processes (gloo ranks)


def _rng_raster(shape, dtype, seed=0):
    rng = np.random.default_rng(seed)
    if np.dtype(dtype).kind == "f":
        return (rng.standard_normal(shape) * 50 + 100).astype(dtype)
    info = np.iinfo(dtype)
    return rng.integers(info.min, info.max, shape, dtype=dtype, endpoint=True)


@pytest.mark.parametrize("mode,dtype,kw", [("F", np.float32, {}), ("F", np.float32, {"compression": "tiff_adobe_deflate"}),
                                           ("F", np.float32, {"big_tiff": True}), ("L", np.uint8, {}),
                                           ("I;16", np.uint16, {}), ("F", np.float32, {"compression": "tiff_deflate"}),
                                           ("F", np.float32, {"compression": "tiff_lzw"}),
                                           ("L", np.uint8, {"compression": "tiff_lzw"}),
                                           ("I;16", np.uint16, {"compression": "tiff_lzw"})])
def test_reads_what_libtiff_writes(tmp_path, mode, dtype, kw):
    a = _rng_raster((123, 77), dtype, 1)
    path = tmp_path / "a.tif"
    im = Image.fromarray(a)
    assert im.mode == mode
    im.save(path, format="TIFF", **kw)
    info = geotiff.read_info(path)
    assert info.shape == a.shape and info.big == bool(kw.get("big_tiff"))
    got = geotiff.read_array(path, pinned=False)
    assert got.dtype == a.dtype and got.flags.c_contiguous
    np.testing.assert_array_equal(got, a)


@pytest.mark.parametrize("dtype", [np.float32, np.int16, np.uint8, np.float64, np.int32])
@pytest.mark.parametrize("strip_bytes", [1 << 20, 600])
def test_writer_round_trip_and_libtiff_reads_it(tmp_path, dtype, strip_bytes):
    a = _rng_raster((67, 131), dtype, 2)
    path = tmp_path / "w.tif"
    geotiff.write_geotiff(path, a, strip_bytes=strip_bytes, nodata=-32768)
    info = geotiff.read_info(path)
    assert info.dtype == np.dtype(dtype) and info.nodata == -32768.0 and geotiff._contiguous(info)
    assert len(info.offsets) == (1 if strip_bytes > a.nbytes else -(-67 // max(1, strip_bytes // (131 * a.itemsize))))
    np.testing.assert_array_equal(geotiff.read_array(path, pinned=False), a)
    if dtype in (np.float32, np.uint8, np.int32):                    # the sample types Pillow maps to an image mode
        np.testing.assert_array_equal(np.array(Image.open(path)), a)


def _build_tiff(blocks, tags, bo="<"):
    """A classic TIFF from raw blocks + {tag: (type, values)}: offsets / bytecounts tags are filled in here."""
    fmt = {3: "H", 4: "I"}
    data = b"".join(blocks)
    offs, pos = [], 8
    for b in blocks:
        offs.append(pos)
        pos += len(b)
    tiled = 322 in tags
    tags = dict(tags)
    tags[324 if tiled else 273] = (4, offs)
    tags[325 if tiled else 279] = (4, [len(b) for b in blocks])
    ifd_off = (8 + len(data) + 1) // 2 * 2
    n = len(tags)
    extra_off = ifd_off + 2 + 12 * n + 4
    entries, extras = b"", b""
    for tag in sorted(tags):
        typ, vals = tags[tag]
        raw = struct.pack(bo + fmt[typ] * len(vals), *vals)
        if len(raw) <= 4:
            val = raw.ljust(4, b"\0")
        else:
            val = struct.pack(bo + "I", extra_off + len(extras))
            extras += raw + b"\0" * (len(raw) % 2)
        entries += struct.pack(bo + "HHI", tag, typ, len(vals)) + val
    head = (b"II" if bo == "<" else b"MM") + struct.pack(bo + "HI", 42, ifd_off)
    return head + data + b"\0" * (ifd_off - 8 - len(data)) + struct.pack(bo + "H", n) + entries + struct.pack(bo + "I", 0) + extras


@pytest.mark.parametrize("bo", ["<", ">"])
def test_tiles_deflate_predictor_and_byte_order(tmp_path, bo):
    """A tiled, Deflate-compressed int16 raster with the horizontal predictor, ragged edge tiles, both byte orders (GDAL's
    COMPRESS=DEFLATE PREDICTOR=2 TILED=YES layout), built by hand."""
    a = _rng_raster((70, 100), np.int16, 3)
    tw = th = 32
    blocks = []
    for ty in range(0, 70, th):
        for tx in range(0, 100, tw):
            t = np.zeros((th, tw), dtype=np.int16)
            part = a[ty:ty + th, tx:tx + tw]
            t[:part.shape[0], :part.shape[1]] = part
            d = t.copy()
            d[:, 1:] = t[:, 1:] - t[:, :-1]                           # wraps modulo 2^16
            blocks.append(zlib.compress(d.astype(bo + "i2").tobytes()))
    tags = {256: (4, [100]), 257: (4, [70]), 258: (3, [16]), 259: (3, [8]), 262: (3, [1]), 277: (3, [1]), 284: (3, [1]),
            317: (3, [2]), 322: (3, [tw]), 323: (3, [th]), 339: (3, [2])}
    path = tmp_path / "t.tif"
    path.write_bytes(_build_tiff(blocks, tags, bo))
    got = geotiff.read_array(path, pinned=False)
    assert got.dtype == np.int16
    np.testing.assert_array_equal(got, a)
    # big-endian uncompressed strips: the contiguous fast path + byte swap
    strips = [a[y:y + 16].astype(bo + "i2").tobytes() for y in range(0, 70, 16)]
    tags = {256: (4, [100]), 257: (4, [70]), 258: (3, [16]), 259: (3, [1]), 262: (3, [1]), 277: (3, [1]), 278: (4, [16]),
            284: (3, [1]), 339: (3, [2])}
    path.write_bytes(_build_tiff(strips, tags, bo))
    assert geotiff._contiguous(geotiff.read_info(path))
    np.testing.assert_array_equal(geotiff.read_array(path, pinned=False), a)


def test_array2raster_copies_the_georeference_of_a_gdal_raster(tmp_path):
    """utils_dem.array2raster (utils_dem.py:17-40): float32 output, geotransform + projection from ``rasterfn``.  The
    source tags are the ones of the reference's GDAL-written resources/images/final_dem.tif (golden fixture)."""
    g = load_golden("run_geotiff")
    src = tmp_path / "src.tif"
    ends = np.cumsum(g["tag_raw_len"])
    tags = {int(t): (int(typ), int(cnt), g["tag_raw"][e - n:e].tobytes()) for t, typ, cnt, n, e in
            zip(g["tag_ids"], g["tag_types"], g["tag_counts"], g["tag_raw_len"], ends)}
    geotiff.write_geotiff(src, g["array"], geo_tags=tags)
    np.testing.assert_array_equal(geotiff.read_array(src, pinned=False), g["array"])
    info = geotiff.read_info(src)
    np.testing.assert_allclose(info.geotransform(), g["geotransform"], rtol=0, atol=0)
    final = (g["array"].astype(np.float64) * 1.0000001 + 0.25)       # a float64 result, like hydro_dem_process.py:149-151
    out = tmp_path / "out.tif"
    geotiff.array2raster(str(out), final, str(src))
    oinfo = geotiff.read_info(out)
    assert oinfo.dtype == np.float32 and oinfo.shape == final.shape
    assert oinfo.geo_tags() == info.geo_tags() and len(oinfo.geo_tags()) == len(tags) >= 4
    np.testing.assert_array_equal(geotiff.read_array(out, pinned=False), final.astype(np.float32))
    pil = Image.open(out)
    np.testing.assert_array_equal(np.array(pil), final.astype(np.float32))
    for t, (typ, cnt, raw) in tags.items():                           # libtiff sees the same georeference
        v = pil.tag_v2[t]
        if typ == 12:
            np.testing.assert_array_equal(np.asarray(v, dtype=np.float64), np.frombuffer(raw, dtype="<f8"))
        elif typ == 3:
            np.testing.assert_array_equal(np.asarray(v), np.frombuffer(raw, dtype="<u2"))
    # without a reference raster: no georeference, still a readable float32 raster
    geotiff.array2raster(str(out), final)
    assert geotiff.read_info(out).geo_tags() == {}


def test_unsupported_files_fail_loudly(tmp_path):
    a = _rng_raster((40, 40), np.float32, 4)
    p = tmp_path / "packbits.tif"
    Image.fromarray(a).save(p, format="TIFF", compression="packbits")
    with pytest.raises(geotiff.GeoTiffError, match="compression 32773"):
        geotiff.read_array(p, pinned=False)
    (tmp_path / "x.tif").write_bytes(b"not a tiff at all")
    with pytest.raises(geotiff.GeoTiffError, match="not a TIFF"):
        geotiff.read_info(tmp_path / "x.tif")
    rgb = np.zeros((8, 8, 3), dtype=np.uint8)
    Image.fromarray(rgb).save(tmp_path / "rgb.tif", format="TIFF")
    with pytest.raises(geotiff.GeoTiffError, match="samples per pixel"):
        geotiff.read_info(tmp_path / "rgb.tif")
    with pytest.raises(geotiff.GeoTiffError):
        geotiff.write_geotiff(tmp_path / "y.tif", np.zeros(5))


def test_bigtiff_writer_layout(tmp_path):
    """The > 4 GB layout (BigTIFF: 8-byte offsets, LONG8 strip tables) on a small raster, forced."""
    a = _rng_raster((50, 60), np.float32, 5)
    p = tmp_path / "big.tif"
    geotiff.write_geotiff(p, a, strip_bytes=1000, nodata=0, bigtiff=True)
    assert p.read_bytes()[:4] == b"II+\x00"
    info = geotiff.read_info(p)
    assert info.big and len(info.offsets) > 1 and info.nodata == 0.0
    np.testing.assert_array_equal(geotiff.read_array(p, pinned=False), a)
    np.testing.assert_array_equal(np.array(Image.open(p)), a)


def test_parallel_pieces(tmp_path, monkeypatch):
    """Reads and writes larger than one I/O piece are split over the thread pool (pread into disjoint slices of the
    result; conversion of the output pieces on the pool, writes in order): forced here with a tiny piece size."""
    monkeypatch.setattr(geotiff, "_PIECE", 1000)
    a = _rng_raster((301, 257), np.float32, 9)
    p = tmp_path / "p.tif"
    geotiff.write_geotiff(p, a.astype(np.float64), dtype=np.float32, strip_bytes=3000)     # converted piece by piece
    np.testing.assert_array_equal(np.array(Image.open(p)), a)
    np.testing.assert_array_equal(geotiff.read_array(p, pinned=False), a)
    geotiff.write_geotiff(p, a[:, ::2], strip_bytes=500)                                   # non-contiguous source rows
    np.testing.assert_array_equal(geotiff.read_array(p, pinned=False), a[:, ::2])
    # a truncated file fails loudly instead of returning garbage
    data = p.read_bytes()
    info = geotiff.read_info(p)
    p.write_bytes(data[:int(info.offsets[0]) + 5000] )
    with pytest.raises(geotiff.GeoTiffError):
        geotiff.read_array(p, pinned=False)


@pytest.mark.parametrize("kind", ["smooth", "noise", "constant", "long"])
def test_lzw_decoder_against_libtiff(tmp_path, kind):
    """hd_host_lzw_decode on streams written by libtiff: compressible rasters (long strings, table resets after 4094
    entries, every code width) and incompressible noise; strips of many rows so that one stream crosses several resets."""
    rng = np.random.default_rng(11)
    if kind == "smooth":
        yy, xx = np.mgrid[0:400, 0:700]
        a = np.round(100 + 20 * np.sin(xx / 40.0) + 15 * np.cos(yy / 55.0)).astype(np.int16).view(np.uint16)
    elif kind == "noise":
        a = rng.integers(0, 65535, (300, 500), dtype=np.uint16)
    elif kind == "constant":
        a = np.full((500, 900), 7, dtype=np.uint16)
    else:
        a = np.repeat(rng.integers(0, 4, (1, 250000), dtype=np.uint8), 2, axis=0).reshape(500, 1000).astype(np.uint16)
    p = tmp_path / "l.tif"
    Image.fromarray(a).save(p, format="TIFF", compression="tiff_lzw", tiffinfo={278: a.shape[0]})
    info = geotiff.read_info(p)
    assert info.compression == 5
    np.testing.assert_array_equal(geotiff.read_array(p, pinned=False), a)
    np.testing.assert_array_equal(np.array(Image.open(p)), a)


def test_lzw_rejects_corrupt_streams():
    import ctypes
    from hydrodem_b200 import _lib
    lib = _lib.load()
    out = np.zeros(64, dtype=np.uint8)
    bad = np.array([0x80, 0x3F, 0xFF, 0xFF, 0xFF], dtype=np.uint8)       # Clear, then a code far beyond the table
    assert lib.hd_host_lzw_decode(ctypes.c_void_p(bad.ctypes.data), len(bad), ctypes.c_void_p(out.ctypes.data), 64) < 0
    assert lib.hd_host_lzw_decode(None, 0, ctypes.c_void_p(out.ctypes.data), 64) < 0


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_floating_point_predictor(tmp_path, dtype):
    """PREDICTOR=3 (GDAL's choice for compressed float rasters): bytes grouped by significance, then differenced.  The
    encoder here is the definition written out (TIFF Technical Note 3), strips compressed with Deflate."""
    a = _rng_raster((50, 77), dtype, 13)
    a[3, 4] = np.nan
    es = a.itemsize
    strips = []
    for y in range(0, 50, 16):
        rows = a[y:y + 16]
        be = rows.astype(">f" + str(es)).view(np.uint8).reshape(rows.shape[0], 77, es)
        planes = np.ascontiguousarray(be.transpose(0, 2, 1)).reshape(rows.shape[0], 77 * es)    # most significant bytes first
        d = planes.copy()
        d[:, 1:] = planes[:, 1:] - planes[:, :-1]
        strips.append(zlib.compress(d.tobytes()))
    tags = {256: (4, [77]), 257: (4, [50]), 258: (3, [8 * es]), 259: (3, [8]), 262: (3, [1]), 277: (3, [1]), 278: (4, [16]),
            284: (3, [1]), 317: (3, [3]), 339: (3, [3])}
    p = tmp_path / "fp.tif"
    p.write_bytes(_build_tiff(strips, tags))
    got = geotiff.read_array(p, pinned=False)
    assert got.dtype == np.dtype(dtype)
    np.testing.assert_array_equal(got, a)


@pytest.mark.parametrize("comp", ["tiff_lzw", "tiff_adobe_deflate"])
def test_horizontal_predictor_written_by_libtiff(tmp_path, comp):
    """GDAL's COMPRESS=LZW|DEFLATE PREDICTOR=2 for integer elevation rasters, produced by libtiff itself."""
    yy, xx = np.mgrid[0:300, 0:400]
    a = np.round(100 + 20 * np.sin(xx / 40.0) + 15 * np.cos(yy / 55.0)).astype(np.int16).view(np.uint16)
    p = tmp_path / "p2.tif"
    Image.fromarray(a).save(p, format="TIFF", compression=comp, tiffinfo={317: 2})
    info = geotiff.read_info(p)
    assert info.predictor == 2 and len(info.offsets) > 1
    np.testing.assert_array_equal(geotiff.read_array(p, pinned=False), a)
