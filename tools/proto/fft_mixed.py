"""Prototype (NumPy) of the mixed-radix in-place DIT row FFT of csrc/fft.cu and its shared-memory bank model.

n = r_1 r_2 ... r_p.  Position p of the shared array holds x[perm(p)] (digit reversal, built recursively:
perm_N(q M + p') = r_last perm_M(p') + q).  Pass s (radix R = r_s, Lprev = r_1..r_{s-1}) works on groups
{base + j + q Lprev}: x_q *= W_{Lprev R}^(j q), DFT_R in registers, stored back in place.  Output: natural order.
Checks the index algebra against numpy.fft and prints, per pass order, the shared-memory wavefronts per warp access
(64-bit accesses: 16 lanes per wavefront, bank pair = padded index mod 16).
"""
import itertools
import sys

import numpy as np

ALLOWED = [16, 15, 12, 10, 9, 8, 6, 5, 4, 3, 2]


def factorizations(n, maxr=16):
    """All multisets of allowed radices with product n, fewest passes first."""
    out = []

    def rec(rem, start, cur):
        if rem == 1:
            out.append(tuple(cur))
            return
        for i in range(start, len(ALLOWED)):
            r = ALLOWED[i]
            if rem % r == 0:
                rec(rem // r, i, cur + [r])
    rec(n, 0, [])
    out.sort(key=lambda t: (len(t), max(t)))
    return out


def perm_table(radices):
    """perm[p] = input index held at position p before the first pass (radices in pass order)."""
    perm = np.zeros(1, dtype=np.int64)
    for r in radices:
        m = len(perm)
        perm = np.concatenate([r * perm + q for q in range(r)])
    return perm


def fft_mixed(x, radices):
    n = len(x)
    s = np.asarray(x, dtype=np.complex128)[perm_table(radices)]
    lprev = 1
    for r in radices:
        L = lprev * r
        g = np.arange(n // r)
        j, blk = g % lprev, g // lprev
        base = blk * L + j
        idx = base[:, None] + np.arange(r)[None, :] * lprev
        v = s[idx] * np.exp(-2j * np.pi * (j[:, None] * np.arange(r)[None, :]) / L)
        s[idx] = np.fft.fft(v, axis=1)
        lprev = L
    return s


def dif_position(k, radices, n):
    """Where X[k] sits after the DIF passes (radices in pass order): csrc/fft.cu dif_position."""
    pos, L = 0, n
    for r in radices:
        stride = L // r
        pos += (k % r) * stride
        k //= r
        L = stride
    return pos


def fft_mixed_dif(x, radices):
    """In-place DIF: natural order in, X[k] at dif_position(k) (what the CUDA kernel does)."""
    n = len(x)
    s = np.asarray(x, dtype=np.complex128).copy()
    L = n
    for r in radices:
        stride = L // r
        g = np.arange(n // r)
        blk, j = g // stride, g % stride
        idx = (blk * L + j)[:, None] + np.arange(r)[None, :] * stride
        y = np.fft.fft(s[idx], axis=1) * np.exp(-2j * np.pi * (j[:, None] * np.arange(r)[None, :]) / L)
        s[idx] = y
        L = stride
    return s[[dif_position(k, radices, n) for k in range(n)]]


def wavefronts(radices, pad, nt=256):
    """Average wavefronts per warp-wide 64-bit access (ideal 2.0) for every pass, first sweep of all warps."""
    n = int(np.prod(radices))
    res = []
    lprev = 1
    for r in radices:
        g = np.arange(min(n // r, nt))
        j, blk = g % lprev, g // lprev
        base = blk * lprev * r + j
        tot = cnt = 0
        for q in range(r):
            a = pad(base + q * lprev)
            for w0 in range(0, len(g), 16):          # half-warps
                lanes = a[w0:w0 + 16]
                banks = lanes % 16
                worst = max(len(set(lanes[banks == b2])) for b2 in set(banks))
                tot += worst
                cnt += 1
        res.append(2.0 * tot / cnt)
        lprev *= r
    return res


PADS = {"none": lambda i: i, "i>>4": lambda i: i + (i >> 4), "i>>5": lambda i: i + (i >> 5), "i>>3": lambda i: i + (i >> 3)}

if __name__ == "__main__":
    rng = np.random.default_rng(1)
    for n, rad in ((30, (2, 3, 5)), (7200, (8, 9, 10, 10)), (6000, (5, 8, 10, 15)), (360, (15, 4, 6))):
        x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
        err = np.abs(fft_mixed(x, rad) - np.fft.fft(x)).max()
        assert err < 1e-9 * n, (n, err)
        err = np.abs(fft_mixed_dif(x, rad) - np.fft.fft(x)).max()
        assert err < 1e-9 * n, ("dif", n, err)
    print("index algebra ok")
    for n in (int(a) for a in sys.argv[1:] or ("7200", "6000")):
        facs = factorizations(n)
        best = [f for f in facs if len(f) == len(facs[0])]
        rows = []
        for f in best:
            for order in set(itertools.permutations(f)):
                for pname, pad in PADS.items():
                    w = wavefronts(order, pad)
                    rows.append((sum(w), order, pname, [round(v, 2) for v in w]))
        rows.sort(key=lambda t: t[0])
        print(n, "passes", len(facs[0]))
        for r in rows[:12]:
            print("   ", r)
