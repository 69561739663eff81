"""Generate hydrodem_b200/csrc/fft_radix.cuh: straight-line in-register DFTs of length R (forward, exp(-2 pi i / R)).

The mixed-radix row FFT of csrc/fft.cu (lengths 2^a 3^b 5^c: 6000, 7200 ... -- the sub-rows of the 18000 / 36000
mosaics) runs passes of radix R in {2, 3, 4, 5, 6, 8, 9, 10, 12, 15, 16}: one thread holds the R points of a
butterfly in registers.  Each dft_fwd<R> is generated here from hand-written 2 / 3 / 5-point kernels combined by
Cooley-Tukey with constant twiddles (trivial ones -- 1, -1, +-i, the eighth roots -- are specialised), on a small
real-valued expression IR with lazy signs (no negation is ever emitted).  Every generated butterfly is evaluated
numerically from the same IR and checked against numpy.fft before the header is written.

    python tools/gen_fft_radix.py            (re)writes the header
    python tools/gen_fft_radix.py --check    verifies only
"""
import math
import os
import sys

import numpy as np

RADICES = list(range(2, 17))           # 7, 11, 13, 14: only as the outer factor n1 of a long row


class Node:
    __slots__ = ("op", "args", "name", "val")

    def __init__(self, op, *args):
        self.op, self.args, self.name, self.val = op, args, None, None


class Builder:
    """Real-valued straight-line code; a value is (sign, Node)."""

    def __init__(self):
        self.order = []

    def node(self, op, *args):
        n = Node(op, *args)
        self.order.append(n)
        return n

    def inp(self, idx, part):
        return (1, self.node("in", idx, part))

    def add(self, a, b):
        (sa, na), (sb, nb) = a, b
        if sa == sb:
            return (sa, self.node("add", na, nb))
        return (1, self.node("sub", na, nb)) if sa > 0 else (1, self.node("sub", nb, na))

    def sub(self, a, b):
        return self.add(a, (-b[0], b[1]))

    def mulc(self, c, a):
        """c * a, c a Python float."""
        if c == 1.0:
            return a
        if c == -1.0:
            return (-a[0], a[1])
        return (1, self.node("mul", c * a[0], a[1]))

    def fmac(self, c, a, b):
        """c * a + b."""
        sa, na = a
        sb, nb = b
        # result sign follows b: sb * (c*sa*sb * na + nb)
        return (sb, self.node("fma", c * sa * sb, na, nb))


def neg(v):
    return (-v[0], v[1])


class Cx:
    """Complex value on a Builder."""

    def __init__(self, b, re, im):
        self.b, self.re, self.im = b, re, im

    def __add__(self, o):
        return Cx(self.b, self.b.add(self.re, o.re), self.b.add(self.im, o.im))

    def __sub__(self, o):
        return Cx(self.b, self.b.sub(self.re, o.re), self.b.sub(self.im, o.im))

    def scale(self, c):
        return Cx(self.b, self.b.mulc(c, self.re), self.b.mulc(c, self.im))

    def fma_scale(self, c, o):
        """c * self + o (c real)."""
        return Cx(self.b, self.b.fmac(c, self.re, o.re), self.b.fmac(c, self.im, o.im))

    def mul_neg_i(self):          # (a + ib)(-i) = b - ia
        return Cx(self.b, self.im, neg(self.re))

    def mul_i(self):
        return Cx(self.b, neg(self.im), self.re)

    def neg(self):
        return Cx(self.b, neg(self.re), neg(self.im))

    def mul_w(self, num, den):
        """self * exp(-2 pi i num / den)."""
        num %= den
        b = self.b
        if num == 0:
            return self
        if 4 * num == den:
            return self.mul_neg_i()
        if 2 * num == den:
            return self.neg()
        if 4 * num == 3 * den:
            return self.mul_i()
        if (8 * num) % den == 0:
            # odd eighth roots: h (1 -+ i) style -- two adds, two multiplies
            h = math.sqrt(0.5)
            k = 8 * num // den
            a, bb = self.re, self.im
            if k == 1:      # h(1 - i): re = h(a + b), im = h(b - a)
                return Cx(b, b.mulc(h, b.add(a, bb)), b.mulc(h, b.sub(bb, a)))
            if k == 3:      # h(-1 - i): re = h(b - a), im = -h(a + b)
                return Cx(b, b.mulc(h, b.sub(bb, a)), b.mulc(-h, b.add(a, bb)))
            if k == 5:      # h(-1 + i): re = -h(a + b), im = h(a - b)
                return Cx(b, b.mulc(-h, b.add(a, bb)), b.mulc(h, b.sub(a, bb)))
            if k == 7:      # h(1 + i): re = h(a - b), im = h(a + b)
                return Cx(b, b.mulc(h, b.sub(a, bb)), b.mulc(h, b.add(a, bb)))
        ang = -2.0 * math.pi * num / den
        c, s = math.cos(ang), math.sin(ang)
        # (a + ib)(c + is) = (ac - bs) + i(as + bc)
        re = b.fmac(c, self.re, b.mulc(-s, self.im))
        im = b.fmac(s, self.re, b.mulc(c, self.im))
        return Cx(b, re, im)


def dft(xs):
    """Forward DFT of the list of Cx, natural order in and out."""
    n = len(xs)
    if n == 1:
        return xs
    if n == 2:
        return [xs[0] + xs[1], xs[0] - xs[1]]
    if n == 3:
        t, d = xs[1] + xs[2], xs[1] - xs[2]
        y0 = xs[0] + t
        m = t.fma_scale(-0.5, xs[0])
        r = d.scale(math.sqrt(3.0) / 2.0).mul_neg_i()          # -i (sqrt3 / 2) d
        return [y0, m + r, m - r]
    if n == 5:
        a1, b1, a2, b2 = xs[1] + xs[4], xs[1] - xs[4], xs[2] + xs[3], xs[2] - xs[3]
        c1, c2 = math.cos(2 * math.pi / 5), math.cos(4 * math.pi / 5)
        s1, s2 = math.sin(2 * math.pi / 5), math.sin(4 * math.pi / 5)
        y0 = xs[0] + a1 + a2
        p1 = a2.fma_scale(c2, a1.fma_scale(c1, xs[0]))
        p2 = a2.fma_scale(c1, a1.fma_scale(c2, xs[0]))
        q1 = b2.fma_scale(s2, b1.scale(s1)).mul_neg_i()        # -i (s1 b1 + s2 b2)
        q2 = b2.fma_scale(-s1, b1.scale(s2)).mul_neg_i()       # -i (s2 b1 - s1 b2)
        return [y0, p1 + q1, p2 + q2, p2 - q2, p1 - q1]
    if all(n % p for p in range(2, n)):
        # other primes (7, 11, 13): direct sum with the symmetric terms paired,
        #   X_k = x_0 + sum_m [cos(2 pi m k / n) (x_m + x_(n-m))  -  i sin(2 pi m k / n) (x_m - x_(n-m))]
        h = (n - 1) // 2
        sp = [xs[m] + xs[n - m] for m in range(1, h + 1)]
        sm = [xs[m] - xs[n - m] for m in range(1, h + 1)]
        y0 = xs[0]
        for v in sp:
            y0 = y0 + v
        out = [y0] + [None] * (n - 1)
        for k in range(1, h + 1):
            ev, od = xs[0], None
            for m in range(1, h + 1):
                ang = 2 * math.pi * ((m * k) % n) / n
                ev = sp[m - 1].fma_scale(math.cos(ang), ev)
                od = sm[m - 1].scale(math.sin(ang)) if od is None else sm[m - 1].fma_scale(math.sin(ang), od)
            od = od.mul_neg_i()
            out[k], out[n - k] = ev + od, ev - od
        return out
    # Cooley-Tukey n = A * B: input index A * n2 + n1, output index k2 + B * k1
    A = 4 if n % 4 == 0 and n > 4 else next(p for p in (2, 3, 5, 7) if n % p == 0)
    if n == 4:
        A = 2
    B = n // A
    sub = [dft([xs[A * n2 + n1] for n2 in range(B)]) for n1 in range(A)]       # sub[n1][k2]
    out = [None] * n
    for k2 in range(B):
        z = dft([sub[n1][k2].mul_w(n1 * k2, n) for n1 in range(A)])
        for k1 in range(A):
            out[k2 + B * k1] = z[k1]
    return out


def build(n):
    b = Builder()
    xs = [Cx(b, b.inp(i, 0), b.inp(i, 1)) for i in range(n)]
    ys = dft(xs)
    return b, ys


def evaluate(b, ys, x):
    for nd in b.order:
        if nd.op == "in":
            v = x[nd.args[0]]
            nd.val = v.real if nd.args[1] == 0 else v.imag
        elif nd.op == "add":
            nd.val = nd.args[0].val + nd.args[1].val
        elif nd.op == "sub":
            nd.val = nd.args[0].val - nd.args[1].val
        elif nd.op == "mul":
            nd.val = nd.args[0] * nd.args[1].val
        elif nd.op == "fma":
            nd.val = nd.args[0] * nd.args[1].val + nd.args[2].val
    return np.array([y.re[0] * y.re[1].val + 1j * (y.im[0] * y.im[1].val) for y in ys])


def live_nodes(b, ys):
    live = set()
    stack = [y.re[1] for y in ys] + [y.im[1] for y in ys]
    while stack:
        nd = stack.pop()
        if id(nd) in live:
            continue
        live.add(id(nd))
        for a in nd.args:
            if isinstance(a, Node):
                stack.append(a)
    return live


def flit(c):
    return repr(float(np.float32(c))) + "f"


def emit(n):
    b, ys = build(n)
    live = live_nodes(b, ys)
    lines, k, ops = [], 0, {"add": 0, "mul": 0, "fma": 0}
    for nd in b.order:
        if id(nd) not in live:
            continue
        if nd.op == "in":
            nd.name = f"x[{nd.args[0]}].{'xy'[nd.args[1]]}"
            continue
        nd.name = f"t{k}"
        k += 1
        if nd.op == "add":
            rhs = f"{nd.args[0].name} + {nd.args[1].name}"
            ops["add"] += 1
        elif nd.op == "sub":
            rhs = f"{nd.args[0].name} - {nd.args[1].name}"
            ops["add"] += 1
        elif nd.op == "mul":
            rhs = f"{flit(nd.args[0])} * {nd.args[1].name}"
            ops["mul"] += 1
        else:
            rhs = f"fmaf({flit(nd.args[0])}, {nd.args[1].name}, {nd.args[2].name})"
            ops["fma"] += 1
        lines.append(f"    const float {nd.name} = {rhs};")
    for i, y in enumerate(ys):
        re = ("-" if y.re[0] < 0 else "") + y.re[1].name
        im = ("-" if y.im[0] < 0 else "") + y.im[1].name
        lines.append(f"    x[{i}] = make_float2({re}, {im});")
    head = (f"// {n}-point: {ops['add']} add, {ops['mul']} mul, {ops['fma']} fma\n"
            f"template <> __device__ __forceinline__ void dft_fwd<{n}>(float2 (&x)[{n}])\n{{\n")
    return head + "\n".join(lines) + "\n}\n"


def check():
    rng = np.random.default_rng(5)
    for n in RADICES:
        b, ys = build(n)
        x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
        got = evaluate(b, ys, x)
        want = np.fft.fft(x)
        err = np.abs(got - want).max()
        assert err < 1e-12, (n, err)
    return True


def main():
    check()
    if "--check" in sys.argv:
        print("all butterflies match numpy.fft")
        return
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "hydrodem_b200", "csrc", "fft_radix.cuh")
    parts = ["// GENERATED by tools/gen_fft_radix.py -- do not edit.  In-register forward DFTs (exp(-2 pi i / R)), natural\n"
             "// order in and out; each was checked against numpy.fft from the same expression graph it is printed from.\n"
             "#pragma once\n#include <cuda_runtime.h>\n\n"
             "template <int R> __device__ __forceinline__ void dft_fwd(float2 (&x)[R]);\n\n"]
    for n in RADICES:
        parts.append(emit(n))
        parts.append("\n")
    with open(out, "w") as f:
        f.write("".join(parts))
    print("wrote", out)


if __name__ == "__main__":
    main()
