"""Row-band distributed fft2 / ifft2 over real NCCL (torchrun, one rank per GPU): parity with the single-GPU transform
of the same synthetic mosaic and timing.   torchrun --nproc-per-node N tools/band_fft_check.py [n]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from hydrodem_b200 import _lib, device as dev, sharding
from hydrodem_b200.synth import SynthScene
import ctypes

n = int(sys.argv[1]) if len(sys.argv) > 1 else 3601
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
comm = sharding.DistComm()
band = sharding.Band(comm, n, n)
sc = SynthScene(n, n, 1004)
x = dev.upload(sc.srtm((band.r0, band.r1)))
for _ in range(2):
    spec = band.fft2(x); back = band.ifft2(spec)
torch.cuda.synchronize(); dist.barrier()
t0 = time.perf_counter()
for _ in range(5):
    spec = band.fft2(x)
torch.cuda.synchronize(); dist.barrier()
t_f = (time.perf_counter() - t0) / 5
err = float((back.tensor().real - x.tensor()).abs().max())
out = {"n": n, "world": comm.world, "fft2_ms": t_f * 1e3, "roundtrip_max_abs_err": err}
if n <= 8192 and comm.rank == 0:                                  # single-GPU reference of the whole mosaic on rank 0
    lib = _lib.load()
    full = dev.upload(sc.srtm())
    plan = dev.fft_plan(n, n); nb = lib.hd_fft2_workspace_bytes(n, n); work = dev.scratch(nb)
    ref = dev.empty(n, n, _lib.C64, np.complex64)
    _lib.check(lib.hd_fft2_c2c(plan, full.ptr, full.dtype, full.pitch, ref.ptr, ref.pitch, 0, ctypes.c_void_p(work.data_ptr()), nb, dev.stream_ptr()))
    c0, c1 = sharding.band_bounds(n, comm.world)[0]
    d = (spec.tensor() - ref.tensor()[:, c0:c1].t()).abs().max()
    out["max_abs_diff_vs_single_gpu"] = float(d); out["spectrum_max_abs"] = float(ref.tensor().abs().max())
if comm.rank == 0:
    print(json.dumps(out), flush=True)
dist.destroy_process_group()
