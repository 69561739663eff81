"""Per-phase host / device times of the banded chain (rank 0), under torchrun:
    HD_BAND_TRACE=1 python -m torch.distributed.run --nproc-per-node N ... tools/band_phases.py [size]"""
import json, os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hydrodem_b200 import _lib, device as dev, sharding
from hydrodem_b200.synth import DeviceMosaic

n = int(sys.argv[1]) if len(sys.argv) > 1 else 36000
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
band = sharding.Band(sharding.DistComm(), n, n)
mosaic = DeviceMosaic(n, n, 1005)
d_srtm = dev.DeviceRaster(torch.empty((band.rows, n), dtype=torch.float32, device=dev.device()), band.rows, n, _lib.F32, np.float32)
g_ext, h_ext = band.alloc_ext(_lib.U8, np.uint8), band.alloc_ext(_lib.F32, np.float32)
mosaic.band(band.r0, band.r1, out={"srtm": d_srtm.tensor(), "groves": g_ext.owned().tensor(), "hsheds": h_ext.owned().tensor()})
for _ in range(3):
    out = band.conditioning_chain(d_srtm, g_ext, h_ext); del out
sharding.TRACE.report()
agg = {}
reps = 3
for _ in range(reps):
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    out = band.conditioning_chain(d_srtm, g_ext, h_ext); del out
    for name, host_ms, dev_ms in sharding.TRACE.report():
        a = agg.setdefault(name, [0.0, 0.0]); a[0] += host_ms / reps; a[1] += dev_ms / reps
if dist.get_rank() == 0:
    print(json.dumps({"size": n, "world": dist.get_world_size(), "fill_rounds": band.fill_rounds,
                      "phases_host_ms_dev_ms": {k: [round(v[0], 2), round(v[1], 2)] for k, v in agg.items()},
                      "total_host_ms": round(sum(v[0] for v in agg.values()), 2), "total_dev_ms": round(sum(v[1] for v in agg.values()), 2)}))
dist.destroy_process_group()
