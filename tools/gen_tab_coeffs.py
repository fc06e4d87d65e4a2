#!/usr/bin/env python
"""Generate the log-spaced value+derivative tables of the gaussianerf / gaussian U/J kernels
(flowvpm.jl_b200/csrc/vpm_tab_coeffs.cuh, committed).

For a family with regularising function g(s) (src/FLOWVPM_kernel.jl:51-66) the pair loop needs
    A = g/r^3 = G/sigma^3            G(s) = g(s)/s^3
    B = (dg/(sigma r) - 3g/r^2)/r^3 = (1/s) dG/ds / sigma^5 = 2 dG/du / sigma^5,   u = s^2.
G is tabulated as a function of u = s^2 on intervals that are uniform in the bits of t = u + c:
interval `row` = the top (11 + LOGN) bits of t minus those of 2^EMIN, i.e. 2^LOGN intervals per
octave of t.  gaussianerf: G is analytic in u, c = 4 = 2^EMIN, rows from u = 0.  gaussian:
G = (1 - exp(-u^1.5))/u^1.5 is NOT analytic at u = 0, so c = 0 (pure log spacing: every interval is
narrow relative to its distance from 0, where the u^1.5 terms are smooth) and the rows start at
u = 2^EMIN = 1/4; the rare pairs closer than s = 1/2 take an exp-based evaluation on the device.
On each interval G(v) ~ p(xi), xi = (t - t_lo)/width - 1/2 in [-1/2, 1/2), a degree-7 polynomial
(Chebyshev interpolant computed at 60 digits); the device evaluates p and dp/dxi with one
joint Horner pass: c0..c4 in FP64, the three highest coefficients in FP32 (their terms are
< 2^-33 of the value).  Packed row (48 bytes = 3 x 16-byte shared-memory chunks):
    [c0 c1] [c2 c3] [c4* (t5,t6)]     c4*: the low 20 bits of c4's mantissa hold t7 as
                                      sign + 8 exponent + 11 mantissa bits (t7 = float(bits << 12))
Beyond the regularised range (g == 1 to < 2e-16: s >= 9 resp. 3.45) G is the pure power law
u^-3/2.  128 rows stored in front of the others tabulate it on the mantissa of u itself (t = u, no
offset), 2^LOGN rows per octave over the two octaves [1, 4): the odd/even exponent parity is a
row-index bit and the remaining scale 2^(-3 (e >> 1)) is an exponent shift.  With them a warp whose
lanes straddle the cut-off evaluates ONE code path (no second rsqrt-based evaluation).
The error quoted in the header is that of the PACKED row evaluated the way the device does
(FP32 tail emulated with numpy float32), against mpmath.
"""
import os
import struct
import sys

import mpmath as mp
import numpy as np

mp.mp.dps = 60
f32 = np.float32
C2 = mp.sqrt(2 / mp.pi)


# ---------------------------------------------------------------- the tabulated functions
def gerf_G(u):
    u = mp.mpf(u)
    if u < mp.mpf("0.75"):
        return C2 * mp.fsum((-1) ** (k + 1) * u ** (k - 1) * 2 * k / ((2 * k + 1) * mp.mpf(2) ** k * mp.factorial(k))
                            for k in range(1, 70))
    s = mp.sqrt(u)
    return (mp.erf(s / mp.sqrt(2)) - C2 * s * mp.exp(-u / 2)) / s ** 3


def gerf_dG(u):  # dG/du = H/2, H = (sqrt(2/pi) e^{-u/2} - 3G)/u
    u = mp.mpf(u)
    if u < mp.mpf("0.75"):
        return C2 * mp.fsum((-1) ** (k + 1) * (k - 1) * u ** (k - 2) * 2 * k / ((2 * k + 1) * mp.mpf(2) ** k * mp.factorial(k))
                            for k in range(2, 70))
    return (C2 * mp.exp(-u / 2) - 3 * gerf_G(u)) / (2 * u)


def gaus_Gu(u):
    return gaus_G(mp.sqrt(mp.mpf(u)))


def gaus_dGu(u):
    s = mp.sqrt(mp.mpf(u))
    return gaus_dG(s) / (2 * s)


def gaus_G(s):
    s = mp.mpf(s)
    if s < mp.mpf("0.3"):
        v = s ** 3
        return mp.fsum((-v) ** k / mp.factorial(k + 1) for k in range(0, 40))
    return (1 - mp.exp(-s ** 3)) / s ** 3


def gaus_dG(s):
    s = mp.mpf(s)
    if s < mp.mpf("0.3"):
        return mp.fsum((-1) ** k * 3 * k * s ** (3 * k - 1) / mp.factorial(k + 1) for k in range(1, 40))
    e = mp.exp(-s ** 3)
    return 3 * (s ** 3 * e - (1 - e)) / s ** 4


FAMILIES = {
    # name: (G(u), dG/du, offset c, EMIN = exponent of the first row's t, LOGN, largest u the table must cover)
    "Gerf": (gerf_G, gerf_dG, 4, 2, 6, 81.0),          # far field (g == 1) from s = 9
    "Gaus": (gaus_Gu, gaus_dGu, 0, -2, 6, 11.9025),    # far field from s = 3.45; u < 1/4 off the table
}
DEG = 7
NHEAD = 5


def cheb_mono(f, deg):
    """monomial coefficients in xi in [-1/2, 1/2] of the Chebyshev interpolant of f(xi)"""
    n = deg + 1
    nodes = [mp.cos(mp.pi * (k + mp.mpf(1) / 2) / n) for k in range(n)]
    fx = [f(t / 2) for t in nodes]
    c = [2 * mp.fsum(fx[k] * mp.cos(mp.pi * j * (k + mp.mpf(1) / 2) / n) for k in range(n)) / n for j in range(n)]
    c[0] /= 2
    T = [[mp.mpf(1)], [mp.mpf(0), mp.mpf(1)]]
    for j in range(2, n):
        cur = [mp.mpf(0)] + [2 * v for v in T[j - 1]]
        for i, v in enumerate(T[j - 2]):
            cur[i] -= v
        T.append(cur)
    mono = [mp.mpf(0)] * n
    for j in range(n):
        for i, v in enumerate(T[j]):
            mono[i] += c[j] * v
    return [mono[i] * mp.mpf(2) ** i for i in range(n)]


def f2u(x):
    return struct.unpack("<I", struct.pack("<f", float(x)))[0]


def u2f(b):
    return struct.unpack("<f", struct.pack("<I", b & 0xffffffff))[0]


def d2u(x):
    return struct.unpack("<Q", struct.pack("<d", float(x)))[0]


def u2d(b):
    return struct.unpack("<d", struct.pack("<Q", b))[0]


def pack_row(co):
    """co: 8 mpf monomial coefficients -> 6 uint64 words of the packed row"""
    c = [float(v) for v in co[:NHEAD]]
    t5, t6, t7 = (f2u(f32(float(v))) for v in co[NHEAD:])
    # t7: round to sign + 8 exponent + 11 mantissa bits, stored in the low 20 bits of c4
    t7r = (t7 + 0x800) >> 12
    c4 = (d2u(c[4]) & ~0xfffff) | (t7r & 0xfffff)
    return [d2u(c[0]), d2u(c[1]), d2u(c[2]), d2u(c[3]), c4, (t6 << 32) | t5]


def eval_packed(words, xi):
    """the device's evaluation order (FP32 tail in numpy float32, FP64 head in Python floats)"""
    c0, c1, c2, c3 = (u2d(w) for w in words[:4])
    c4 = u2d(words[4])
    t7 = f32(u2f((words[4] & 0xfffff) << 12))
    t5, t6 = f32(u2f(words[5] & 0xffffffff)), f32(u2f(words[5] >> 32))
    xf = f32(np.floor((xi + 0.5) * 2 ** 23) / 2 ** 23 - 0.5)  # top 23 bits of the mantissa field
    b6 = f32(xf * t7 + t6)
    b5 = f32(xf * b6 + t5)
    d5 = f32(xf * t7 + b6)
    d4 = f32(xf * d5 + b5)
    b = xi * float(b5) + c4
    d = xi * float(d4) + b
    for c in (c3, c2, c1):
        b = xi * b + c
        d = xi * d + b
    b = xi * b + c0
    return b, d


def far_rows(name, logn):
    """power-law rows u^-3/2 for t = u in [1, 4) (two octaves: the exponent parity is a row bit); the row
    order follows the index bits the device uses: (E & 1) << LOGN | mantissa bits, E the BIASED
    exponent -- an odd biased exponent is an even true exponent."""
    pw = mp.mpf(-3) / 2
    n = 1 << logn
    octaves = [1, 0]   # true-exponent parity of the first / second block of rows
    rows, errG, errD = [], mp.mpf(0), mp.mpf(0)
    for par in octaves:
        for j in range(n):
            t0 = mp.mpf(2) ** par * (1 + mp.mpf(j) / n)
            w = mp.mpf(2) ** par / n
            mid = t0 + w / 2
            co = cheb_mono(lambda xi: (mid + xi * w) ** pw, DEG)
            words = pack_row(co)
            rows.append(words)
            for xi in np.linspace(-0.5, 0.5 - 2.0 ** -30, 9):
                v = mid + mp.mpf(float(xi)) * w
                b, d = eval_packed(words, float(xi))
                errG = max(errG, abs((mp.mpf(b) - v ** pw) / v ** pw))
                td = pw * v ** (pw - 1)
                errD = max(errD, abs((mp.mpf(d) / w - td) / td))
    return rows, errG, errD


def family_rows(name):
    G, dG, c, emin, logn, vmax = FAMILIES[name]
    c = mp.mpf(c)
    first = mp.mpf(2) ** emin
    n = 1 << logn
    nrows = int(mp.ceil(mp.log((vmax + c) / first, 2) * n))
    rows, errG, errD = [], mp.mpf(0), mp.mpf(0)
    for row in range(nrows):
        e, j = divmod(row, n)
        t0 = first * 2 ** e * (1 + mp.mpf(j) / n)
        w = first * 2 ** e / n
        mid = t0 + w / 2
        co = cheb_mono(lambda xi: G(mid + xi * w - c), DEG)
        words = pack_row(co)
        rows.append(words)
        for xi in np.linspace(-0.5, 0.5 - 2.0 ** -30, 9):
            v = mid + mp.mpf(float(xi)) * w - c
            if v <= 0:
                continue
            b, d = eval_packed(words, float(xi))
            tg, td = G(v), dG(v)
            errG = max(errG, abs((mp.mpf(b) - tg) / tg))
            errD = max(errD, abs((mp.mpf(d) / w - td) / td))
    frows, ferrG, ferrD = far_rows(name, logn)
    # device row order: the power-law rows first (their index is a bit mask), then the regularised range
    return dict(offset=int(c), emin=emin, logn=logn, nrows=nrows, rows=frows + rows, errG=errG, errD=errD,
                nfar=len(frows), ferrG=ferrG, ferrD=ferrD)


def main():
    out = ["// GENERATED by tools/gen_tab_coeffs.py -- do not edit by hand.",
           "// Log-spaced value+derivative tables of G = g(s)/s^3 for the gaussianerf and gaussian U/J kernels;",
           "// layout, packing and the accuracy figures are described in the generator's docstring.",
           "#pragma once", "#include <cstdint>", "namespace vpm {", ""]
    for name in FAMILIES:
        r = family_rows(name)
        print(name, "rows", r["nrows"], "errG", mp.nstr(r["errG"], 3), "errD", mp.nstr(r["errD"], 3),
              "far rows", r["nfar"], "errG", mp.nstr(r["ferrG"], 3), "errD", mp.nstr(r["ferrD"], 3), flush=True)
        out.append(f"// {name}: t = u + {r['offset']}, first row at t = 2^{r['emin']}, 2^{r['logn']} intervals per octave, {r['nrows']} rows; max rel err of the packed rows")
        out.append(f"// evaluated as on the device: G {mp.nstr(r['errG'], 3)}, dG/du {mp.nstr(r['errD'], 3)};")
        out.append(f"// before them (rows 0..{r['nfar'] - 1}) the power-law rows u^-3/2 (t = u): G {mp.nstr(r['ferrG'], 3)}, dG/du {mp.nstr(r['ferrD'], 3)}")
        out.append(f"constexpr int kTab{name}Offset = {r['offset']};")
        out.append(f"constexpr int kTab{name}Emin = {r['emin']};")
        out.append(f"constexpr int kTab{name}LogN = {r['logn']};")
        out.append(f"constexpr int kTab{name}Rows = {r['nrows']};     // regularised range")
        out.append(f"constexpr int kTab{name}FarRows = {r['nfar']};  // power-law rows, stored first")
        out.append(f"__device__ static const uint64_t kTab{name}[{r['nrows'] + r['nfar']} * 6] = {{")
        for words in r["rows"]:
            out.append("    " + ", ".join(f"0x{w:016x}ull" for w in words) + ",")
        out.append("};")
        out.append("")
    out.append("}  // namespace vpm")
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "flowvpm.jl_b200", "csrc", "vpm_tab_coeffs.cuh")
    with open(path, "w") as fh:
        fh.write("\n".join(out) + "\n")
    print("wrote", os.path.normpath(path))


if __name__ == "__main__":
    main()
