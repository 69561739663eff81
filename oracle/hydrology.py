"""Oracle (test infrastructure): the NEW hydrology stages -- sink-fill and D8.

Neither exists in the reference: **parity unpinned**.
The definitions below are this repo's own;
sink-fill is cross-checked by two independent algorithms (iterative
Planchon-Darboux here vs priority-flood in oracle/c/hydro_oracle.c).
"""
import numpy as np

from . import clib

D8_DY = (0, 1, 1, 1, 0, -1, -1, -1)               # E SE S SW W NW N NE
D8_DX = (1, 1, 0, -1, -1, -1, 0, 1)
D8_CODES = (1, 2, 4, 8, 16, 32, 64, 128)
INV_SQRT2_F32 = np.float32(0.70710678)


def sinkfill_iterative(z, max_iter=1000000):
    """Planchon & Darboux (2001) with eps = 0, 8-connectivity, Jacobi sweeps.

    W = z on the one-cell frame, +inf inside; NaN cells are outlets at -inf
    (and stay NaN in the output); repeat W(c) = max(z(c), min_N8 W(n))
    wherever that lowers W(c), to the fixed point (unique for eps = 0).
    Returns (W float32, sweeps)."""
    z = np.asarray(z, dtype=np.float32)
    ny, nx = z.shape
    nan = np.isnan(z)
    w = np.full(z.shape, np.inf, dtype=np.float32)
    w[0, :] = z[0, :]; w[-1, :] = z[-1, :]; w[:, 0] = z[:, 0]; w[:, -1] = z[:, -1]
    w[nan] = -np.inf
    interior = np.zeros(z.shape, dtype=bool)
    interior[1:-1, 1:-1] = True
    interior &= ~nan
    zi = np.where(nan, np.float32(-np.inf), z)
    for it in range(max_iter):
        p = np.pad(w, 1, mode='constant', constant_values=np.inf)
        m = np.full(z.shape, np.inf, dtype=np.float32)
        for dy, dx in zip(D8_DY, D8_DX):
            m = np.minimum(m, p[1 + dy:1 + dy + ny, 1 + dx:1 + dx + nx])
        cand = np.maximum(zi, m)
        upd = interior & (cand < w)
        if not upd.any():
            out = w.copy()
            out[nan] = np.nan
            return out, it
        w[upd] = cand[upd]
    raise RuntimeError("sinkfill_iterative did not converge")


def sinkfill(z):
    """Priority-flood fill (C oracle, ``ho_priority_flood``)."""
    return clib.priority_flood(z)


def d8(w):
    """D8 flow direction on a (filled) surface, ESRI codes, uint8.

    For every non-frame cell with a non-NaN value: the neighbour with the
    largest positive drop W(c) - W(n), diagonal drops multiplied by
    0.70710678f in float32; ties keep the first in the order E, SE, S, SW, W,
    NW, N, NE; no positive drop -> 0.  NaN neighbours never win."""
    w = np.asarray(w, dtype=np.float32)
    ny, nx = w.shape
    out = np.zeros(w.shape, dtype=np.uint8)
    c = w[1:-1, 1:-1]
    best = np.zeros(c.shape, dtype=np.float32)
    code = np.zeros(c.shape, dtype=np.uint8)
    with np.errstate(invalid='ignore'):
        for k, (dy, dx) in enumerate(zip(D8_DY, D8_DX)):
            n = w[1 + dy:ny - 1 + dy, 1 + dx:nx - 1 + dx]
            drop = (c - n).astype(np.float32)
            if k & 1:
                drop = (drop * INV_SQRT2_F32).astype(np.float32)
            better = drop > best
            best = np.where(better, drop, best)
            code = np.where(better, np.uint8(D8_CODES[k]), code)
    out[1:-1, 1:-1] = code
    return out
