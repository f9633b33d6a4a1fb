#!/usr/bin/env python3
"""Generate the standard-normal quantile tables used by the Philox path (DESIGN.md "z16").

The Philox path turns one 16-bit uniform h into a standard normal deviate by table lookup:

    sign = h >> 15,  i = h & 0x7FFF
    cell i of the half-normal = [a_i, b_i),  a_i = ndtri(0.5 + i/65536),  b_i = ndtri(0.5 + (i+1)/65536)
    z = +-Z1[i],  Z1[i] = sqrt(E[z^2 | a_i <= z < b_i])          (the cell's conditional RMS)

so every cell has probability 2^-16 and the discrete law has mean 0 and variance EXACTLY 1
(sum_i 2^-15 * Z1[i]^2 = E[z^2] = 1).  The 2 outermost cells (i >= 32766, z > 4.0 sigma,
probability 2^-14) are refined once more with 13 fresh bits: sub-cell j of tail cell t=i-32766 is
[ndtri(0.5 + (i + j/8192)/65536), ndtri(0.5 + (i + (j+1)/8192)/65536)) with representative
Z2[t*8192+j] = conditional RMS again, which keeps the variance exact and extends the support to
~6.2 sigma (the reference's Box-Muller on a 31-bit Lehmer uniform reaches 6.55 sigma,
/root/reference src/rand.h:79-94).

Output: squigulator_b200/data/ztable_v2.bin =
    Z16[65536] float16 LE : Z16[h] = (h & 0x8000 ? -1 : +1) * fp16(Z1[h & 0x7FFF])  (2-byte entries let the
                            128 KB table sit in shared memory with the sign folded into the index; the
                            fp16 rounding error, <= 2^-12 relative and zero-mean, is far below one ADC step)
    Z2[16384]  float32 LE : the refined tail cells (magnitude; the sign comes from h)
The file is data shared by the product (embedded into libsqg.so) and by the oracle (loaded at run
time); tests/test_ztable.py re-derives it independently with mpmath.
"""
import os
import sys

import numpy as np
from scipy.special import ndtri, ndtr

N1 = 32768
TAIL_CELLS = 2
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
    e1 = ndtri(0.5 + np.arange(N1 + 1, dtype=np.float64) / 65536.0)  # e1[N1] = ndtri(1.0) = inf
    z1 = np.empty(N1, dtype=np.float64)
    z1[:-1] = cond_rms(e1[:-2], e1[1:-1])
    z1[-1] = cond_rms_inf(e1[-2])

    z2 = np.empty(TAIL_CELLS * SUB, dtype=np.float64)
    first = N1 - TAIL_CELLS
    for t in range(TAIL_CELLS):
        i = first + t
        # upper-tail probability of sub-cell edges, computed as 0.5 - (i + j/SUB)/65536 exactly in binary
        q = 0.5 - (i + np.arange(SUB + 1, dtype=np.float64) / SUB) / 65536.0
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


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(
        os.path.dirname(os.path.abspath(__file__)), "..", "squigulator_b200", "data", "ztable_v2.bin")
    z1, z2 = build()
    assert np.all(np.diff(z1) > 0) and np.all(np.diff(z2) > 0)
    os.makedirs(os.path.dirname(out), exist_ok=True)
    z1h = z1.astype(np.float16)
    assert np.all(np.diff(z1h.astype(np.float32)) >= 0) and np.all(np.isfinite(z1h))
    z16 = np.concatenate([z1h, -z1h])  # index = 16-bit uniform h: bit 15 is the sign
    with open(out, "wb") as f:
        f.write(z16.astype("<f2").tobytes())
        f.write(z2.astype("<f4").tobytes())
    vh = body_h = (z1h[:N1 - TAIL_CELLS].astype(np.float64) ** 2).sum() / N1 + (z2.astype(np.float64) ** 2).sum() / (N1 * SUB)
    print(f"variance of the shipped law (fp16 body + fp32 tail) = {vh:.9f}; tail threshold fp16(Z1[{N1 - TAIL_CELLS}]) = {float(z1h[N1 - TAIL_CELLS])!r}")
    v1 = (z1.astype(np.float64) ** 2).mean()
    # exact variance of the two-level law: body cells + refined tail cells
    body = (z1[:N1 - TAIL_CELLS].astype(np.float64) ** 2).sum() / N1
    tail = (z2.astype(np.float64) ** 2).sum() / (N1 * SUB)
    print(f"wrote {out}: Z1[0]={z1[0]:.3e} Z1[-1]={z1[-1]:.4f} Z2[-1]={z2[-1]:.4f} "
          f"var(level1)={v1:.9f} var(two-level)={body + tail:.9f}")


if __name__ == "__main__":
    main()
