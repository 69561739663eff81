"""Full conditioning chain on a large synthetic mosaic (one GPU): inputs are generated band by band on the host
and assembled in HBM, the chain runs device-resident, and size-independent properties are checked:
  * the final DEM is integer valued; the fill is idempotent and >= the DEM; every D8 code points strictly downhill;
  * the Fourier blanking mask is point-symmetric about DC (odd sizes) and blanks < 1 % of the spectrum;
  * a crop of the chain input re-run through the CPU oracle agrees on the exact-class stages (lagoons branch).

    python tools/big_chain.py 10801          (C3 / C5 sizes: 10801, 18000, 36000)
"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hydrodem_b200 import _lib, device as dev                                  # noqa: E402
from hydrodem_b200.filters import new_filters as nf                            # noqa: E402
from hydrodem_b200.pipeline import ConditioningChain                           # noqa: E402
from hydrodem_b200.synth import SynthScene                                     # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10801
gpu_synth = "--gpu-synth" in sys.argv
seed = {10801: 1003, 18000: 1004, 36000: 1005}.get(n, 1000)
band = 1024
t0 = time.time()
d_srtm, d_hs, d_gr = dev.empty(n, n, _lib.F32), dev.empty(n, n, _lib.F32), dev.empty(n, n, _lib.U8)
if not gpu_synth:
    sc = SynthScene(n, n, seed)
    for r0 in range(0, n, band):
        r1 = min(n, r0 + band)
        for dst, arr in ((d_srtm, sc.srtm((r0, r1))), (d_hs, sc.hsheds((r0, r1))), (d_gr, sc.groves((r0, r1)))):
            dst.tensor()[r0:r1].copy_(torch.from_numpy(arr))
else:
    # Same recipe as hydrodem_b200/synth.py, generated on the device with torch (the NumPy generator needs ~25 min
    # for 36000^2): trend + 6 octaves of bilinear value noise + stripes + sensor noise + canopy; HydroSHEDS = rounded
    # textured base with constant lagoon plateaus and voids.  Not bit-identical to SynthScene -- this script only
    # checks size-independent properties.
    import torch.nn.functional as F
    g = torch.Generator(device="cuda").manual_seed(seed)
    ts, th, tg = d_srtm.tensor(), d_hs.tensor(), d_gr.tensor()
    coarse = [(2 ** (9 - k), 8.0 * 0.5 ** k,
               torch.randn((1, 1, n // 2 ** (9 - k) + 3, n // 2 ** (9 - k) + 3), generator=g, device="cuda")) for k in range(6)]
    xs = torch.arange(n, device="cuda", dtype=torch.float64)
    for r0 in range(0, n, band):
        r1 = min(n, r0 + band)
        ys = torch.arange(r0, r1, device="cuda", dtype=torch.float64)
        base = (100.0 + 1e-4 * xs[None, :] + 5e-5 * ys[:, None]).float()
        for step, amp, c in coarse:
            gy = (ys / step / (c.shape[2] - 1) * 2 - 1).float()
            gx = (xs / step / (c.shape[3] - 1) * 2 - 1).float()
            grid = torch.stack(torch.broadcast_tensors(gx[None, :], gy[:, None]), dim=-1)[None]
            base += amp * F.grid_sample(c, grid, mode="bilinear", align_corners=True)[0, 0]
        stripes = (0.5 * torch.sin(2 * np.pi * (0.11 * xs[None, :] + 0.07 * ys[:, None]))
                   + 0.3 * torch.sin(2 * np.pi * (0.031 * xs[None, :] - 0.052 * ys[:, None]))).float()
        ts[r0:r1] = base + stripes + 0.3 * torch.randn((r1 - r0, n), generator=g, device="cuda")
        th[r0:r1] = torch.round(base + 0.55 * torch.randn((r1 - r0, n), generator=g, device="cuda"))
    tg.zero_()
    rng = np.random.default_rng(seed)
    for _ in range(n * n // 40000):
        hgt, wid = int(rng.integers(3, 9)), int(rng.integers(20, 120))
        if rng.random() < 0.5:
            hgt, wid = wid, hgt
        y0, x0 = int(rng.integers(0, n - hgt)), int(rng.integers(0, n - wid))
        tg[y0:y0 + hgt, x0:x0 + wid] = 1
        ts[y0:y0 + hgt, x0:x0 + wid] += float(rng.uniform(2, 6))
    yy, xx = torch.meshgrid(torch.arange(-40, 41, device="cuda"), torch.arange(-40, 41, device="cuda"), indexing="ij")
    rr = yy * yy + xx * xx
    for _ in range(n * n // 25000):
        cy, cx, rad = int(rng.integers(41, n - 41)), int(rng.integers(41, n - 41)), float(rng.uniform(6, 40))
        blk = th[cy - 40:cy + 41, cx - 40:cx + 41]
        disc = rr <= rad * rad
        blk[disc] = blk[40, 40] - 2.0
    nv = max(1, int(n * n * 2e-5))
    th[torch.from_numpy(rng.integers(1, n - 1, nv)).cuda(), torch.from_numpy(rng.integers(1, n - 1, nv)).cuda()] = -32768.0
torch.cuda.synchronize()
print(f"inputs {n}x{n} generated + uploaded in {time.time() - t0:.1f} s", flush=True)

chain = ConditioningChain(keep_intermediates=True)
times = []
for rep in range(2):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    res = chain.run_device(d_srtm, d_gr, d_hs)
    b.record()
    torch.cuda.synchronize()
    times.append(a.elapsed_time(b))
cells = n * n
if "--profile" in sys.argv:
    import ctypes
    lib = _lib.load()
    lib.hd_profile_enable(1)
    chain.run_device(d_srtm, d_gr, d_hs)
    cbuf = ctypes.create_string_buffer(1 << 16)
    lib.hd_profile_report(cbuf, 1 << 16)
    lib.hd_profile_enable(0)
    rows = [l.split() for l in cbuf.value.decode().splitlines()]
    tot = sum(float(r[2]) for r in rows)
    prof = {r[0]: {"launches": int(r[1]), "ms": round(float(r[2]), 3), "share": round(float(r[2]) / tot, 4)}
            for r in sorted(rows, key=lambda r: -float(r[2]))}
    print(json.dumps({"size": n, "profile_total_ms": tot, "kernels": prof}), flush=True)
print(f"chain: {times[-1]:.2f} ms  -> {cells / times[-1] / 1e3:.0f} Mcells/s   (first run {times[0]:.2f} ms), "
      f"peak HBM in use {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB", flush=True)

R = res.rasters
final, filled, d8 = R["final"].tensor(), R["filled"].tensor(), R["d8"].tensor()
checks = {}
checks["final_integer_valued"] = bool((final == torch.round(final)).all().item())
checks["filled_ge_final"] = bool((filled >= final.float()).all().item())
refill = nf.SinkFill(want_stats=False).run_device(R["filled"]).tensor()
checks["fill_idempotent"] = bool(torch.equal(refill, filled))
# D8: every coded cell has a strictly lower neighbour in the coded direction
dy = [0, 1, 1, 1, 0, -1, -1, -1]; dx = [1, 1, 0, -1, -1, -1, 0, 1]
ok = True
inner = filled[1:-1, 1:-1]
for k in range(8):
    sel = d8[1:-1, 1:-1] == (1 << k)
    nb = filled[1 + dy[k]:n - 1 + dy[k], 1 + dx[k]:n - 1 + dx[k]]
    ok &= bool((nb[sel] < inner[sel]).all().item())
checks["d8_points_downhill"] = ok
checks["d8_frame_zero"] = bool((d8[0].sum() + d8[-1].sum() + d8[:, 0].sum() + d8[:, -1].sum()).item() == 0)
mask = R["fourier_mask"].tensor()
checks["mask_fraction"] = float(mask.float().mean().item())
if n % 2 == 1:
    checks["mask_point_symmetric"] = bool(torch.equal(mask, torch.flip(mask, (0, 1))))
# exact-class cross-check of a crop against the CPU oracle (lagoons branch is local: halo 14 rows / cols)
from oracle import chain as ochain                                             # noqa: E402
c0, cs = n // 2, 400
crop_hs = d_hs.tensor()[c0:c0 + cs, c0:c0 + cs].cpu().numpy()
want = ochain.lagoons_detection(crop_hs.copy())
got = R["lagoons_values"].tensor()[c0:c0 + cs, c0:c0 + cs].cpu().numpy().astype(np.float64)
m = 20
checks["lagoons_crop_exact"] = bool(np.array_equal(got[m:-m, m:-m], want["TidyingLagoons"][m:-m, m:-m]))
print(json.dumps({"size": n, "ms": times[-1], "mcells_s": cells / times[-1] / 1e3, "checks": checks}), flush=True)
