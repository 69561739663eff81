"""BASELINE.json configs[3]: sink-fill + D8 on a synthetic 18000 x 18000 mosaic, row-band sharded over the ranks of
one box (torch.distributed.run, one rank per GPU, NCCL halo exchange).  Verifies against the single-GPU result
computed on rank 0 (bit-identical) and prints device-timed Mcells/s (max over ranks).

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/band_fill_scale.py [n]
"""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from hydrodem_b200 import _lib, device as dev, sharding                     # noqa: E402
from hydrodem_b200.filters import new_filters as nf                          # noqa: E402
import gpu_synth                                                             # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 18000
comm = sharding.DistComm()
band = sharding.Band(comm, n, n)
zt = gpu_synth.terrain(n, 1004, rows=(band.r0, band.r1), rounded=True)
z = dev.empty(band.rows, n, _lib.F32)
z.tensor().copy_(zt)
times = []
for rep in range(3):
    torch.cuda.synchronize(); dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    w, d8 = band.sinkfill(z)
    b.record()
    torch.cuda.synchronize(); dist.barrier()
    t = torch.tensor([a.elapsed_time(b)], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    times.append(float(t.item()))
# verification: rank 0 fills the whole mosaic alone and compares every band
ok = True
if rank == 0:
    full = dev.empty(n, n, _lib.F32)
    full.tensor().copy_(gpu_synth.terrain(n, 1004, rounded=True))
    ref_w = nf.SinkFill(want_stats=False).run_device(full)
    ref_d8 = nf.D8FlowDirection().run_device(ref_w)
    bounds = sharding.band_bounds(n, world)
    ok &= bool(torch.equal(w.tensor(), ref_w.tensor()[bounds[0][0]:bounds[0][1]]))
    ok &= bool(torch.equal(d8.tensor(), ref_d8.tensor()[bounds[0][0]:bounds[0][1]]))
    for r in range(1, world):
        r0, r1 = bounds[r]
        bw = torch.empty((r1 - r0, n), dtype=torch.float32, device="cuda")
        bd = torch.empty((r1 - r0, n), dtype=torch.uint8, device="cuda")
        dist.recv(bw, r); dist.recv(bd, r)
        ok &= bool(torch.equal(bw, ref_w.tensor()[r0:r1])) and bool(torch.equal(bd, ref_d8.tensor()[r0:r1]))
    single = []
    for rep in range(2):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        rw = nf.SinkFill(want_stats=False).run_device(full); nf.D8FlowDirection().run_device(rw)
        b.record(); torch.cuda.synchronize()
        single.append(a.elapsed_time(b))
    print(json.dumps({"config": "sink-fill + D8, row-band sharded", "size": n, "n_gpus": world, "identical_to_single_gpu": ok,
                      "ms": min(times[1:]), "mcells_s": n * n / min(times[1:]) / 1e3, "halo_rounds": band.fill_rounds,
                      "single_gpu_ms_same_box": min(single), "single_gpu_mcells_s": n * n / min(single) / 1e3}), flush=True)
else:
    dist.send(w.tensor().contiguous(), 0); dist.send(d8.tensor().contiguous(), 0)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
