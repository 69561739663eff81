"""Upper bound of the streaming host API: the per-tile PCIe traffic alone, both directions at once, no kernels."""
import ctypes, time, sys
import numpy as np, torch
sys.path.insert(0, ".")
from hydrodem_b200 import _lib, device as dev
lib = _lib.load()
n = 3601 * 3601
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
up = [(dev.pinned_empty((n,), np.float32), torch.empty(n, dtype=torch.float32, device="cuda")),
      (dev.pinned_empty((n,), np.uint8), torch.empty(n, dtype=torch.uint8, device="cuda")),
      (dev.pinned_empty((n,), np.float32), torch.empty(n, dtype=torch.float32, device="cuda"))]
down = [(dev.pinned_empty((n,), np.float32), torch.empty(n, dtype=torch.float32, device="cuda")),
        (dev.pinned_empty((n,), np.float32), torch.empty(n, dtype=torch.float32, device="cuda")),
        (dev.pinned_empty((n,), np.uint8), torch.empty(n, dtype=torch.uint8, device="cuda"))]
def h2d():
    for h, d in up:
        lib.hd_memcpy2d_h2d(ctypes.c_void_p(d.data_ptr()), h.nbytes, ctypes.c_void_p(h.ctypes.data), h.nbytes, h.nbytes, 1, ctypes.c_void_p(s1.cuda_stream))
def d2h():
    for h, d in down:
        lib.hd_memcpy2d_d2h(ctypes.c_void_p(h.ctypes.data), h.nbytes, ctypes.c_void_p(d.data_ptr()), h.nbytes, h.nbytes, 1, ctypes.c_void_p(s2.cuda_stream))
for name, fn in (("h2d only", lambda: h2d()), ("d2h only", lambda: d2h()), ("both", lambda: (h2d(), d2h()))):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20):
        fn()
    torch.cuda.synchronize()
    print(name, (time.perf_counter() - t0) / 20 * 1e3, "ms per tile")
