#!/usr/bin/env python3
"""Generate the standard-normal quantile tables used by the Philox path (DESIGN.md "Z32").

The Philox path turns ten random bits r (bit 9 = sign, bits 0-8 = m) plus a five-bit CLASS c (the low bits of the
draw's Philox block number, i.e. of its position) into a standard normal deviate by one table lookup:

    fine cell f = 32*m + (m odd ? 31-c : c)  of the half-normal, f in [0, 16384):
        [a_f, b_f),  a_f = ndtri(0.5 + f/32768),  b_f = ndtri(0.5 + (f+1)/32768)
    z = +-S[c]*Z1[f],  Z1[f] = sqrt(E[z^2 | a_f <= z < b_f])      (the cell's conditional RMS)

Every fine cell has probability 2^-15 (2^-14 of the half-normal); a class owns 512 cells spread evenly over the
quantile range (one per run of 32, alternating direction so that the classes' moments nearly agree), and the 32
classes partition the table.  Sum_f 2^-14 * Z1[f]^2 = E[z^2] = 1 holds over all cells; the per-class factor S[c]
(within 0.6 % of 1) makes EVERY class a law of mean 0 and variance exactly 1, so the noise level does not depend on
the sample position.  The outermost cell (f = 16383 = class 0, m = 511; |z| > 4.0 sigma, probability 2^-14 counting
both signs) is refined with 13 fresh bits: sub-cell j is
[ndtri(0.5 + (16383 + j/8192)/32768), ndtri(0.5 + (16383 + (j+1)/8192)/32768)) with representative
Z2[j] = S[0] * conditional RMS again, which keeps the variance exact and extends the support to ~5.9 sigma (the
reference's Box-Muller on a 31-bit Lehmer uniform reaches 6.55 sigma, /root/reference src/rand.h:79-94).

Output: squigulator_b200/data/ztable_v3.bin =
    Z32[32768] float32 LE : Z32[(r << 5) | c] = (r bit 9 ? -1 : +1) * S[c] * Z1[f(m, c)].  The 4-byte entry index has
                            the class in its low five bits = the shared-memory bank, so the 32 lanes of a warp
                            (32 consecutive Philox blocks) never collide.  The two tail entries (m = 511, c = 0) hold
                            a quiet NaN: the sample kernel's range check catches it and takes the refinement path.
    Z2[8192]   float32 LE : the refined tail cell (magnitude, scaled by S[0]; the sign comes from r)
The file is data shared by the product (embedded into libsqg.so) and by the oracle (loaded at run time);
tests/test_ztable.py re-derives it independently with mpmath.
"""
import os
import sys

import numpy as np
from scipy.special import ndtri, ndtr

N1 = 16384
TAIL_CELLS = 1
SUB = 8192


def cond_rms(a, b):
    """sqrt(E[z^2 | a<=z<b]) for finite cells, 16-point Gauss-Legendre per cell (no cancellation)."""
    x, w = np.polynomial.legendre.leggauss(16)
    mid = 0.5 * (a + b)[:, None]
    half = 0.5 * (b - a)[:, None]
    z = mid + half * x[None, :]
    pdf = np.exp(-0.5 * z * z)
    num = (w[None, :] * z * z * pdf).sum(axis=1)
    den = (w[None, :] * pdf).sum(axis=1)
    return np.sqrt(num / den)


def cond_rms_inf(a):
    """sqrt(E[z^2 | z>=a]) = sqrt(1 + a*phi(a)/Q(a))"""
    phi = np.exp(-0.5 * a * a) / np.sqrt(2 * np.pi)
    q = ndtr(-a)
    return np.sqrt(1.0 + a * phi / q)


def build():
    e1 = ndtri(0.5 + np.arange(N1 + 1, dtype=np.float64) / 32768.0)  # e1[N1] = ndtri(1.0) = inf
    z1 = np.empty(N1, dtype=np.float64)
    z1[:-1] = cond_rms(e1[:-2], e1[1:-1])
    z1[-1] = cond_rms_inf(e1[-2])

    z2 = np.empty(TAIL_CELLS * SUB, dtype=np.float64)
    first = N1 - TAIL_CELLS
    for t in range(TAIL_CELLS):
        i = first + t
        # upper-tail probability of sub-cell edges, computed as 0.5 - (i + j/SUB)/32768 exactly in binary
        q = 0.5 - (i + np.arange(SUB + 1, dtype=np.float64) / SUB) / 32768.0
        with np.errstate(divide="ignore"):
            e = -ndtri(q)  # q[-1] == 0 for the last cell -> +inf
        zz = np.empty(SUB, dtype=np.float64)
        if np.isinf(e[-1]):
            zz[:-1] = cond_rms(e[:-2], e[1:-1])
            zz[-1] = cond_rms_inf(e[-2])
        else:
            zz[:] = cond_rms(e[:-1], e[1:])
        z2[t * SUB:(t + 1) * SUB] = zz
    return z1.astype(np.float32), z2.astype(np.float32)


def fine_cell(m, c):
    """class c, slot m -> fine half-normal cell"""
    return 32 * m + np.where(m & 1, 31 - c, c)


def tables():
    """(Z32 as float32 incl. NaN sentinels, Z2 float32, per-class scale S float64)"""
    z1, z2 = build()
    z1d, z2d = z1.astype(np.float64), z2.astype(np.float64)
    sq = z1d ** 2
    sq[N1 - 1] = (z2d ** 2).mean()  # second moment of the refined tail cell
    m = np.arange(512)
    scale = np.array([1.0 / np.sqrt(sq[fine_cell(m, c)].mean()) for c in range(32)])
    half = np.empty(N1, dtype=np.float64)  # index (m << 5) | c
    for c in range(32):
        half[(m << 5) | c] = scale[c] * z1d[fine_cell(m, c)]
    z32 = np.concatenate([half, -half]).astype("<f4")
    bits = z32.view("<u4").copy()
    tail_idx = (511 << 5) | 0
    bits[tail_idx] = 0x7FC00000          # tail sentinels (quiet NaN), either sign
    bits[N1 + tail_idx] = 0x7FC00000
    return bits.view("<f4"), (scale[0] * z2d).astype("<f4"), scale


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(
        os.path.dirname(os.path.abspath(__file__)), "..", "squigulator_b200", "data", "ztable_v3.bin")
    z1, _ = build()
    assert np.all(np.diff(z1) > 0)
    z32, z2, scale = tables()
    assert np.all(np.diff(z2) > 0)
    os.makedirs(os.path.dirname(out), exist_ok=True)
    with open(out, "wb") as f:
        f.write(z32.tobytes())
        f.write(z2.tobytes())
    h = z32[:N1].astype(np.float64)
    var = []
    for c in range(32):
        v = h[c::32] ** 2
        if c == 0:
            v[511] = (z2.astype(np.float64) ** 2).mean()
        var.append(v.mean())
    print(f"wrote {out}: Z1[0]={z1[0]:.3e} Z2[0]={z2[0]:.6f} Z2[-1]={z2[-1]!r}; class scale {scale.min():.6f}..{scale.max():.6f}; "
          f"class variance {min(var):.8f}..{max(var):.8f}")


if __name__ == "__main__":
    main()
