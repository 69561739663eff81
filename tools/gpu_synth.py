"""Device-side synthetic terrain for the large-mosaic tools (same recipe as hydrodem_b200/synth.py, generated with
torch on the GPU; deterministic for a given seed and GPU type, NOT bit-identical to the NumPy generator)."""
import numpy as np
import torch
import torch.nn.functional as F


def terrain(n, seed, rows=None, band=1024, noise=0.3, rounded=False):
    """(rows r0:r1 of) a float32 n x n surface: trend + 6 octaves of bilinear value noise + white noise."""
    r0, r1 = (0, n) if rows is None else rows
    g = torch.Generator(device="cuda").manual_seed(seed)
    coarse = [(2 ** (9 - k), 8.0 * 0.5 ** k,
               torch.randn((1, 1, n // 2 ** (9 - k) + 3, n // 2 ** (9 - k) + 3), generator=g, device="cuda")) for k in range(6)]
    out = torch.empty((r1 - r0, n), dtype=torch.float32, device="cuda")
    xs = torch.arange(n, device="cuda", dtype=torch.float64)
    for b0 in range(0, n, band):                       # the white-noise stream is consumed for every band: rows-invariant
        b1 = min(n, b0 + band)
        wn = torch.randn((b1 - b0, n), generator=g, device="cuda")
        lo, hi = max(b0, r0), min(b1, r1)
        if lo >= hi:
            continue
        ys = torch.arange(lo, hi, device="cuda", dtype=torch.float64)
        base = (100.0 + 1e-4 * xs[None, :] + 5e-5 * ys[:, None]).float()
        for step, amp, c in coarse:
            gy = (ys / step / (c.shape[2] - 1) * 2 - 1).float()
            gx = (xs / step / (c.shape[3] - 1) * 2 - 1).float()
            grid = torch.stack(torch.broadcast_tensors(gx[None, :], gy[:, None]), dim=-1)[None]
            base += amp * F.grid_sample(c, grid, mode="bilinear", align_corners=True)[0, 0]
        v = base + noise * wn[lo - b0:hi - b0]
        out[lo - r0:hi - r0] = torch.round(v) if rounded else v
    return out
