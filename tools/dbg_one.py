import sys, numpy as np, torch
sys.path.insert(0, '.')
from hydrodem_b200.filters import custom_filters as cf, extension_filters as ef
from hydrodem_b200 import device as dev, _lib
which = sys.argv[1]
a = np.random.default_rng(0).random((100, 300)).astype(np.float32)
r = dev.upload(a)
torch.cuda.synchronize(); print("upload ok", flush=True)
if which == "expand":
    out = cf.ExpandFilter(window_size=7).run_device(r)
elif which == "majority":
    out = cf.MajorityFilter(window_size=11).run_device(r)
elif which == "nanfix":
    out = cf.CorrectNANValues().run_device(r)
elif which == "conv":
    out = ef.Convolve().run_device(r)
try:
    torch.cuda.synchronize(); print(which, "kernel ok", flush=True)
except Exception as e:
    print(which, "FAILED", e, flush=True)
