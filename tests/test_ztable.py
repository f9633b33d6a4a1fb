"""The quantile tables (squigulator_b200/data/ztable_v3.bin) re-derived independently (mpmath) + their invariants."""
import numpy as np
import pytest

from tests import helpers as H

N1 = 16384


def fine_cell(m, c):
    return 32 * m + (31 - c if m & 1 else c)


@pytest.fixture(scope="module")
def tables():
    raw = H.load_ztable()
    z32 = raw[:32768 * 4].view("<f4")
    z2 = raw[32768 * 4:].view("<f4").astype(np.float64)
    return z32, z2


def test_layout_and_symmetry(tables):
    z32, z2 = tables
    assert z32.size == 32768 and z2.size == 8192
    bits = z32.view("<u4")
    tail = (511 << 5) | 0
    assert bits[tail] == 0x7FC00000 and bits[N1 + tail] == 0x7FC00000      # the two NaN sentinels ...
    assert np.count_nonzero(np.isnan(z32)) == 2                             # ... and no other
    ok = ~np.isnan(z32[:N1])
    assert np.array_equal(z32[N1:][ok], -z32[:N1][ok])                      # bit 14 of the index is the sign
    assert np.all(z32[:N1][ok] > 0)
    for c in range(32):                                                      # every class: increasing with the slot m
        v = z32[c:N1:32].astype(np.float64)
        v = v[~np.isnan(v)]
        assert np.all(np.diff(v) > 0)
    assert np.all(np.diff(z2) > 0)
    assert abs(float(z2[-1]) - 5.9104958) < 1e-6                            # Z_MAX in sqg_device.cuh


def test_every_class_has_mean_zero_and_unit_variance(tables):
    z32, z2 = tables
    for c in range(32):
        v = z32[c:N1:32].astype(np.float64) ** 2
        if c == 0:
            v[511] = (z2 ** 2).mean()           # the refined tail cell
        assert abs(v.mean() - 1.0) < 2e-7, c


def test_cells_match_mpmath(tables):
    mp = pytest.importorskip("mpmath")
    mp.mp.dps = 30
    z32, z2 = tables

    def edge(p):
        return mp.sqrt(2) * mp.erfinv(2 * mp.mpf(p) - 1)

    def rms(a, b):
        num = mp.quad(lambda z: z * z * mp.npdf(z), [a, b])
        den = mp.quad(lambda z: mp.npdf(z), [a, b])
        return float(mp.sqrt(num / den))

    def cell(f):
        return rms(edge(mp.mpf(1) / 2 + mp.mpf(f) / 32768), edge(mp.mpf(1) / 2 + mp.mpf(f + 1) / 32768))

    # the ratio table / conditional RMS is one constant per class (the unit-variance scale), within 0.7 % of 1
    for c in (0, 5, 31):
        ratios = []
        for m in (0, 1, 100, 255, 400, 510):
            ratios.append(float(z32[(m << 5) | c]) / cell(fine_cell(m, c)))
        assert max(ratios) - min(ratios) < 3e-7, (c, ratios)
        assert abs(ratios[0] - 1.0) < 7e-3
        if c == 0:
            s0 = ratios[0]
    for j in (0, 4096, 8190):
        a = edge(mp.mpf(1) / 2 + (mp.mpf(16383) + mp.mpf(j) / 8192) / 32768)
        b = edge(mp.mpf(1) / 2 + (mp.mpf(16383) + mp.mpf(j + 1) / 8192) / 32768)
        assert abs(s0 * rms(a, b) - z2[j]) < 2e-6
    a = edge(mp.mpf(1) / 2 + (mp.mpf(16383) + mp.mpf(8191) / 8192) / 32768)
    assert abs(s0 * rms(a, mp.inf) - z2[-1]) < 2e-6


def test_classes_partition_the_cells():
    """index = (r10 << 5) | class: a class owns 512 half-normal cells, one in every run of 32 fine cells, and the 32
    classes partition the 16384 cells; the class is the shared-memory bank of the 4-byte entry."""
    seen = np.zeros(N1, dtype=np.int32)
    for c in range(32):
        cells = np.array([fine_cell(m, c) for m in range(512)])
        assert np.array_equal(cells >> 5, np.arange(512))
        seen[cells] += 1
        idx = (np.arange(512) << 5) | c
        assert np.all((idx & 31) == c)
    assert np.all(seen == 1)
    assert fine_cell(511, 0) == N1 - 1          # the refined cell is the outermost one
