"""Real multi-GPU check of the row-band path (run under torch.distributed.run, one rank per GPU):
banded majority + sink-fill/D8 on a synthetic mosaic must equal the single-GPU result of rank 0."""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hydrodem_b200 import device as dev, sharding                      # noqa: E402
from hydrodem_b200.filters import custom_filters as cf, new_filters as nf   # noqa: E402
from hydrodem_b200.synth import SynthScene                             # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
sc = SynthScene(n, n, 1004)
comm = sharding.DistComm()
band = sharding.Band(comm, n, n)
hs = sc.hsheds((band.r0, band.r1))
z = np.round(sc.srtm((band.r0, band.r1)))
d_hs, d_z = dev.upload(hs), dev.upload(z)
for _ in range(2):                                                    # warm-up + timed
    torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
    maj = band.apply(cf.MajorityFilter(window_size=11), d_hs, 5)
    w, d8 = band.sinkfill(d_z)
    torch.cuda.synchronize(); dist.barrier(); dt = time.perf_counter() - t0
got = [dev.download(maj), dev.download(w), dev.download(d8), hs, z]
gathered = [None] * world
dist.all_gather_object(gathered, got)
if rank == 0:
    full_hs, full_z = sc.hsheds(), np.round(sc.srtm())
    ref_maj = cf.MajorityFilter(window_size=11).apply(full_hs)
    ref_w = nf.SinkFill().apply(full_z)
    ref_d8 = nf.D8FlowDirection().apply(ref_w)
    cat = [np.concatenate([g[k] for g in gathered]) for k in range(5)]
    parts = dict(majority=np.array_equal(cat[0], ref_maj), fill=np.array_equal(cat[1], ref_w, equal_nan=True),
                 d8=np.array_equal(cat[2], ref_d8), hs_input=np.array_equal(cat[3], full_hs),
                 z_input=np.array_equal(cat[4], full_z))
    print(parts, "fill diff cells", int((cat[1] != ref_w).sum()), "maj diff", int((cat[0] != ref_maj).sum()), flush=True)
    ok = all(parts.values())
    print(f"band check world={world} n={n}: {'OK' if ok else 'MISMATCH'}  rounds={band.fill_rounds}  "
          f"majority+fill+d8 {dt * 1e3:.1f} ms  ({n * n / dt / 1e6:.0f} Mcells/s)", flush=True)
    if not ok:
        sys.exit(1)
dist.destroy_process_group()
