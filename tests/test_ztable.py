"""The quantile tables (squigulator_b200/data/ztable_v2.bin) re-derived independently (mpmath) + their invariants."""
import numpy as np
import pytest

from tests import helpers as H


@pytest.fixture(scope="module")
def tables():
    raw = H.load_ztable()
    z16 = raw[:65536 * 2].view("<f2").astype(np.float64)
    z2 = raw[65536 * 2:].view("<f4").astype(np.float64)
    return z16, z2


def test_layout_and_symmetry(tables):
    z16, z2 = tables
    assert z16.size == 65536 and z2.size == 2 * 8192
    assert np.array_equal(z16[32768:], -z16[:32768])           # bit 15 of the index is the sign
    assert np.all(np.diff(z16[:32768]) >= 0) and z16[0] > 0     # monotone half-normal quantiles
    assert np.all(np.diff(z2) > 0)
    assert float(z16[32766]) == 4.08203125                      # Z_TAIL_THR in sqg_device.cuh
    assert abs(float(z2[-1]) - 6.0590086) < 1e-6                # Z_MAX in sqg_device.cuh


def test_unit_variance(tables):
    z16, z2 = tables
    body = (z16[:32766] ** 2).sum() / 32768
    tail = (z2 ** 2).sum() / (32768 * 8192)
    assert abs(body + tail - 1.0) < 5e-6      # two-level law, incl. the fp16 rounding of the body
    assert abs((z16[:32768] ** 2).mean() - 1.0) < 5e-6


def test_cells_match_mpmath(tables):
    mp = pytest.importorskip("mpmath")
    mp.mp.dps = 30
    z16, z2 = tables

    def edge(p):
        return mp.sqrt(2) * mp.erfinv(2 * mp.mpf(p) - 1)

    def rms(a, b):
        num = mp.quad(lambda z: z * z * mp.npdf(z), [a, b])
        den = mp.quad(lambda z: mp.npdf(z), [a, b])
        return float(mp.sqrt(num / den))

    for i in (0, 7, 1000, 16384, 30000, 32765):
        a = edge(mp.mpf(1) / 2 + mp.mpf(i) / 65536)
        b = edge(mp.mpf(1) / 2 + mp.mpf(i + 1) / 65536)
        r = rms(a, b)
        assert abs(float(np.float16(r)) - z16[i]) <= abs(r) * 2 ** -10, i   # same value up to one fp16 ulp
    for t, j in ((0, 0), (0, 4096), (1, 8190)):
        i = 32766 + t
        a = edge(mp.mpf(1) / 2 + (mp.mpf(i) + mp.mpf(j) / 8192) / 65536)
        b = edge(mp.mpf(1) / 2 + (mp.mpf(i) + mp.mpf(j + 1) / 8192) / 65536)
        assert abs(rms(a, b) - z2[t * 8192 + j]) < 1e-6
    a = edge(mp.mpf(1) / 2 + (mp.mpf(32767) + mp.mpf(8191) / 8192) / 65536)
    assert abs(rms(a, mp.inf) - z2[-1]) < 1e-6


def test_bank_stratification_covers_the_table():
    """stratify(): replacing bits 1-5 by the block number's low bits keeps 2^11 cells per block class, evenly spread,
    and the 32 classes partition the table."""
    h = np.arange(65536, dtype=np.uint32)
    seen = np.zeros(65536, dtype=np.int32)
    for block in range(32):
        idx = np.unique((h & 0xFFC1) | ((block & 31) << 1))
        assert idx.size == 2048
        assert np.all(((idx >> 1) & 31) == block)          # all in shared-memory bank `block`
        seen[idx] += 1
        # evenly spread over the quantile range: one pair of adjacent cells in every run of 64 cells
        assert np.array_equal(np.unique(idx >> 6), np.arange(1024))
    assert np.all(seen == 1)
