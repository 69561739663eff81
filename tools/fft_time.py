"""Time the stripe-removal stage (DetectApplyFourier) alone on an n x n device raster: python tools/fft_time.py 36000"""
import ctypes, json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hydrodem_b200 import _lib, device as dev
from hydrodem_b200.filters import custom_filters as cf

n = int(sys.argv[1]) if len(sys.argv) > 1 else 18000
lib = _lib.load()
src = dev.empty(n, n, _lib.F32)
t = src.tensor()
for r0 in range(0, n, 2048):
    r1 = min(n, r0 + 2048)
    t[r0:r1] = 100 + 10 * torch.randn((r1 - r0, n), device="cuda")
daf = cf.DetectApplyFourier()
for rep in range(2):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); out = daf.run_device(src); b.record(); torch.cuda.synchronize()
    print(f"{n}: DetectApplyFourier {a.elapsed_time(b):.2f} ms", flush=True)
    del out
lib.hd_profile_enable(1)
out = daf.run_device(src)
cbuf = ctypes.create_string_buffer(1 << 16)
lib.hd_profile_report(cbuf, 1 << 16)
lib.hd_profile_enable(0)
print(json.dumps({"size": n, "kernels": {l.split()[0]: [int(l.split()[1]), round(float(l.split()[2]), 3)] for l in cbuf.value.decode().splitlines()}}))
