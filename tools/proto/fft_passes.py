"""Prototype (NumPy) of the register-blocked radix-2^k FFT passes used by csrc/fft.cu.

DIF passes: natural order in -> bit-reversed order out.  DIT passes: bit-reversed in -> natural out.
A pass of radix R = 2^k on sub-transforms of length L handles groups {base + r + q * (L/R)}, q = 0..R-1:
  DIF:  y = DFT_R(x) by k constant-twiddle radix-2 DIF stages (y in bit-reversed register order),
        then y[p] *= W_L^(r * bitrev_k(p))
  DIT:  x[p] *= W_L^(+-r * bitrev_k(p)), then k constant-twiddle radix-2 DIT stages (natural register order out)
This file checks both against numpy.fft and prints the pass plans.
"""
import numpy as np


def bitrev(v, bits):
    r = 0
    for i in range(bits):
        r |= ((v >> i) & 1) << (bits - 1 - i)
    return r


def plan(log2m):
    n = max(1, -(-log2m // 4))
    base, rem = divmod(log2m, n)
    return [base + 1] * rem + [base] * (n - rem)


def dif_regs(x, k, sign):
    R = 1 << k
    x = list(x)
    for t in range(k):
        Lt = R >> t
        half = Lt // 2
        for blk in range(0, R, Lt):
            for j in range(half):
                a, b = x[blk + j], x[blk + j + half]
                x[blk + j] = a + b
                x[blk + j + half] = (a - b) * np.exp(sign * 2j * np.pi * j / Lt)
    return x


def dit_regs(x, k, sign):
    R = 1 << k
    x = list(x)
    for t in range(k):
        half = 1 << t
        Lt = 2 * half
        for blk in range(0, R, Lt):
            for j in range(half):
                a = x[blk + j]
                b = x[blk + j + half] * np.exp(sign * 2j * np.pi * j / Lt)
                x[blk + j] = a + b
                x[blk + j + half] = a - b
    return x


def fft_dif(s, ks, sign=-1):
    s = np.array(s, dtype=np.complex128)
    M = len(s)
    L = M
    for k in ks:
        R = 1 << k
        st = L // R
        for base in range(0, M, L):
            for r in range(st):
                idx = [base + r + q * st for q in range(R)]
                y = dif_regs(s[idx], k, sign)
                for p in range(R):
                    y[p] *= np.exp(sign * 2j * np.pi * r * bitrev(p, k) / L)
                s[idx] = y
        L //= R
    return s


def fft_dit(s, ks, sign=-1):
    """ks is the DIF plan; the DIT passes run it backwards."""
    s = np.array(s, dtype=np.complex128)
    M = len(s)
    L = 1
    for k in reversed(ks):
        R = 1 << k
        st = L
        L *= R
        for base in range(0, M, L):
            for r in range(st):
                idx = [base + r + q * st for q in range(R)]
                x = list(s[idx])
                for p in range(R):
                    x[p] *= np.exp(sign * 2j * np.pi * r * bitrev(p, k) / L)
                s[idx] = dit_regs(x, k, sign)
    return s


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    for log2m in range(1, 12):
        M = 1 << log2m
        ks = plan(log2m)
        x = rng.normal(size=M) + 1j * rng.normal(size=M)
        br = np.array([bitrev(i, log2m) for i in range(M)])
        f = fft_dif(x, ks)
        assert np.allclose(f, np.fft.fft(x)[br]), ("dif", log2m)
        g = fft_dit(np.fft.fft(x)[br], ks, sign=+1) / M
        assert np.allclose(g, x), ("dit inverse", log2m)
        h = fft_dit(x[br], ks, sign=-1)
        assert np.allclose(h, np.fft.fft(x)), ("dit forward", log2m)
        print(log2m, ks, "ok")
