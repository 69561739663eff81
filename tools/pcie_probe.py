"""PCIe copy rates of the pitched-raster copies used by the host API (one GPU)."""
import ctypes, time, sys
import numpy as np, torch
sys.path.insert(0, ".")
from hydrodem_b200 import _lib, device as dev

lib = _lib.load()
ny = nx = 3601
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def rate(fn, nbytes, n=10):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / n
    return nbytes / dt / 1e9, dt * 1e3


for dt_np in (np.float32, np.float64, np.uint8):
    host = dev.pinned_empty((ny, nx), dt_np)
    host[...] = 1
    r = dev.empty(ny, nx, dev.hd_dtype_of(np.dtype(dt_np)), np.dtype(dt_np))
    es = np.dtype(dt_np).itemsize
    up = lambda: lib.hd_memcpy2d_h2d(r.ptr, r.pitch * es, ctypes.c_void_p(host.ctypes.data), nx * es, nx * es, ny, ctypes.c_void_p(s1.cuda_stream))
    down = lambda: lib.hd_memcpy2d_d2h(ctypes.c_void_p(host.ctypes.data), nx * es, r.ptr, r.pitch * es, nx * es, ny, ctypes.c_void_p(s2.cuda_stream))
    print(np.dtype(dt_np).name, "2D h2d GB/s, ms", rate(up, host.nbytes), "2D d2h", rate(down, host.nbytes))
    flat_d = torch.empty(ny * nx, dtype=torch.from_numpy(host).dtype, device="cuda")
    flat_h = torch.from_numpy(host).reshape(-1)
    print("   1D h2d", rate(lambda: flat_d.copy_(flat_h, non_blocking=True), host.nbytes),
          "1D d2h", rate(lambda: flat_h.copy_(flat_d, non_blocking=True), host.nbytes))
    host2 = dev.pinned_empty((ny, nx), dt_np)
    both = lambda: (up(), lib.hd_memcpy2d_d2h(ctypes.c_void_p(host2.ctypes.data), nx * es, r.ptr, r.pitch * es, nx * es, ny, ctypes.c_void_p(s2.cuda_stream)))
    print("   2D both directions (sum GB/s)", rate(both, 2 * host.nbytes))
