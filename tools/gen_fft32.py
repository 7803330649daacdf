"""Emit gypsum_b200/csrc/fft32_gen.cuh: straight-line split-radix forward DFT codelets (length 4..64) on
separate re[]/im[] register arrays, natural order in and out.

Inverse transforms use the same codelet with the re/im arrays swapped (IDFT(x) = swap(DFT(swap(x)))).
Twiddle constants are folded (1, -j, (1-j)/sqrt2 ... are special-cased) and every split-radix butterfly with
non-trivial twiddles is written in the factored ("tangent") form, in which the two twiddle products, their sum /
difference and the four outputs are 16 multiply-adds instead of 20 operations:
    w^k z = c1 (z.re - t1 z.im, z.im + t1 z.re),  t1 = tan,  and likewise w^3k z' = c3 (...), so that
    w^k z +- w^3k z' = c1 (a' +- (c3/c1) b')  and  out = u +- c1 (...)  -- each line one FMA per component.
The multiply-adds are written as fmaf() so that every kernel and the host emulator round identically.
Run:  python tools/gen_fft32.py
"""
import math
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class Emitter:
    def __init__(self):
        self.lines = []
        self.n = 0
        self.flops = 0

    def tmp(self, expr):
        name = f"t{self.n}"
        self.n += 1
        self.lines.append(f"    const float {name} = {expr};")
        self.flops += 1
        return name

    def lit(self, v):
        return f"{v:.9e}f"

    def fma(self, const, x, y):
        """const * x + y as an explicit fused multiply-add: the rounding must not depend on the contraction choices the
        compiler makes in each inlining context (every kernel and the host emulator have to agree bit for bit)."""
        self.flops += 1
        return self.tmp(f"fmaf({self.lit(const)}, {x}, {y})")

    # complex values are (re_name, im_name) string pairs
    def add(self, a, b):
        return (self.tmp(f"{a[0]} + {b[0]}"), self.tmp(f"{a[1]} + {b[1]}"))

    def sub(self, a, b):
        return (self.tmp(f"{a[0]} - {b[0]}"), self.tmp(f"{a[1]} - {b[1]}"))

    def mul_neg_j(self, a):  # a * (-j) = (im, -re)
        return (a[1], f"(-{a[0]})")

    def mulw(self, a, num, den):
        """a * exp(-2 pi i num/den) with special cases folded."""
        num %= den
        g = math.gcd(num, den) if num else den
        num //= g
        den //= g
        if num == 0:
            return a
        if den == 2:  # -1
            return (f"(-{a[0]})", f"(-{a[1]})")
        if den == 4:
            return self.mul_neg_j(a) if num == 1 else (f"(-{a[1]})", a[0])
        if den == 8:
            r = self.lit(math.sqrt(0.5))
            s, d = self.tmp(f"{a[0]} + {a[1]}"), self.tmp(f"{a[1]} - {a[0]}")
            # (x+jy)(c - js)/..: num=1: ((x+y) + j(y-x))/sqrt2 ; num=3: ((y-x) - j(x+y))/sqrt2 ; 5: -num1 ; 7: -num3
            if num == 1:
                return (self.tmp(f"{r} * {s}"), self.tmp(f"{r} * {d}"))
            if num == 3:
                return (self.tmp(f"{r} * {d}"), self.tmp(f"-{r} * {s}"))
            if num == 5:
                return (self.tmp(f"-{r} * {s}"), self.tmp(f"-{r} * {d}"))
            return (self.tmp(f"-{r} * {d}"), self.tmp(f"{r} * {s}"))
        ang = -2.0 * math.pi * num / den
        c, s = self.lit(math.cos(ang)), self.lit(math.sin(ang))
        re = self.tmp(f"{c} * {a[0]} - {s} * {a[1]}")
        im = self.tmp(f"{c} * {a[1]} + {s} * {a[0]}")
        self.flops += 4
        return (re, im)

    def fft(self, x):
        n = len(x)
        if n == 1:
            return x
        if n == 2:
            return [self.add(x[0], x[1]), self.sub(x[0], x[1])]
        u = self.fft(x[0::2])
        z = self.fft(x[1::4])
        zp = self.fft(x[3::4])
        out = [None] * n
        q = n // 4
        for k in range(q):
            if k == 0:
                a, b = z[0], zp[0]
                s = self.add(a, b)
                d = self.mul_neg_j(self.sub(a, b))
                out[k] = self.add(u[k], s)
                out[k + 2 * q] = self.sub(u[k], s)
                out[k + q] = self.add(u[k + q], d)
                out[k + 3 * q] = self.sub(u[k + q], d)
                continue
            # a = w^k z = c1 * a1, b = w^3k z' = c3 * b1 with a1, b1 two operations each
            if 8 * k == n:  # w = (1 - j)/sqrt2, w^3 = (-1 - j)/sqrt2: c1 = c3 = 1/sqrt2
                zr, zi = z[k]
                pr, pi = zp[k]
                a1 = (self.tmp(f"{zr} + {zi}"), self.tmp(f"{zi} - {zr}"))
                b1 = (self.tmp(f"{pi} - {pr}"), self.tmp(f"-{pr} - {pi}"))
                c1 = math.sqrt(0.5)
                s1 = self.add(a1, b1)
                d1 = self.sub(a1, b1)
            else:
                th1, th3 = -2.0 * math.pi * k / n, -2.0 * math.pi * 3 * k / n
                c1, c3 = math.cos(th1), math.cos(th3)
                t1, t3, rho = math.tan(th1), math.tan(th3), c3 / c1
                zr, zi = z[k]
                pr, pi = zp[k]
                a1 = (self.fma(-t1, zi, zr), self.fma(t1, zr, zi))
                b1 = (self.fma(-t3, pi, pr), self.fma(t3, pr, pi))
                s1 = (self.fma(rho, b1[0], a1[0]), self.fma(rho, b1[1], a1[1]))
                d1 = (self.fma(-rho, b1[0], a1[0]), self.fma(-rho, b1[1], a1[1]))
            ur, ui = u[k]
            vr, vi = u[k + q]
            out[k] = (self.fma(c1, s1[0], ur), self.fma(c1, s1[1], ui))
            out[k + 2 * q] = (self.fma(-c1, s1[0], ur), self.fma(-c1, s1[1], ui))
            # -j (a - b) = c1 * (d1.im, -d1.re)
            out[k + q] = (self.fma(c1, d1[1], vr), self.fma(-c1, d1[0], vi))
            out[k + 3 * q] = (self.fma(-c1, d1[1], vr), self.fma(c1, d1[0], vi))
        return out


def emit(n):
    e = Emitter()
    # snapshot inputs so the in-place writes below cannot alias
    xin = []
    for i in range(n):
        e.lines.append(f"    const float xr{i} = re[{i}], xi{i} = im[{i}];")
        xin.append((f"xr{i}", f"xi{i}"))
    y = e.fft(xin)
    for k in range(n):
        e.lines.append(f"    re[{k}] = {y[k][0]}; im[{k}] = {y[k][1]};")
    head = (f"// forward DFT-{n}, X[k] = sum_n x[n] exp(-2 pi i n k / {n}); ~{e.flops} flops\n"
            f"GB_HD GB_INLINE void fft{n}_fwd(float (&re)[{n}], float (&im)[{n}]) {{\n")
    return head + "\n".join(e.lines) + "\n}\n"


def main():
    out = ["// GENERATED by tools/gen_fft32.py -- do not edit.", "#pragma once", '#include "gb_common.cuh"', ""]
    for n in (4, 8, 16, 32, 64):
        out.append(emit(n))
    path = os.path.join(ROOT, "gypsum_b200", "csrc", "fft32_gen.cuh")
    with open(path, "w") as f:
        f.write("\n".join(out))
    print("wrote", path)


if __name__ == "__main__":
    main()
