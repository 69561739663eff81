"""GeoTIFF ingest / egress throughput, on the GPU box:  python tools/ingest_time.py [n]
Writes synthetic n x n rasters to /tmp (float32 and int16), then times: np.fromfile of the pixel block (the floor: page
cache -> pageable array), geotiff.read_array into pinned memory, geotiff.read_to_device (read chunks + upload
overlapped), dev.upload of an already-read pinned array (the PCIe floor), and array2raster of a float64 result."""
import json, os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hydrodem_b200 import device as dev, geotiff

n = int(sys.argv[1]) if len(sys.argv) > 1 else 18000
rng = np.random.default_rng(1)
base = (rng.standard_normal((n // 8, n)) * 40 + 100).astype(np.float32)
a = np.tile(base, (8, 1))[:n]
out = {"size": n}


def best(fn, reps=3):
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = fn()
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
        del r
    return min(ts)


for name, arr in (("float32", a), ("int16", np.round(a).astype(np.int16))):
    path = f"/tmp/ingest_{name}.tif"
    t0 = time.perf_counter()
    geotiff.write_geotiff(path, arr)
    t_write = time.perf_counter() - t0
    gb = arr.nbytes / 1e9
    info = geotiff.read_info(path)
    t_raw = best(lambda: np.fromfile(path, dtype=arr.dtype, count=arr.size, offset=int(info.offsets[0])))
    t_pin = best(lambda: geotiff.read_array(path))
    t_dev = best(lambda: geotiff.read_to_device(path))
    host = geotiff.read_array(path)
    assert np.array_equal(host, arr)
    t_up = best(lambda: dev.upload(host))
    r = geotiff.read_to_device(path)
    assert np.array_equal(dev.download(r), arr)
    out[name] = {"GB": round(gb, 3), "write_geotiff_GBps": round(gb / t_write, 2), "np_fromfile_GBps": round(gb / t_raw, 2),
                 "read_array_pinned_GBps": round(gb / t_pin, 2), "read_to_device_GBps": round(gb / t_dev, 2),
                 "upload_only_GBps": round(gb / t_up, 2)}
    os.remove(path)
final = a.astype(np.float64)
t0 = time.perf_counter()
geotiff.array2raster("/tmp/final.tif", final)
out["array2raster_f64_to_f32"] = {"GB_written": round(final.size * 4 / 1e9, 3), "GBps": round(final.size * 4 / 1e9 / (time.perf_counter() - t0), 2)}
os.remove("/tmp/final.tif")
out["note"] = "files live in the page cache (just written): disk speed is not part of these numbers"
print(json.dumps(out))
