"""Known-answer tests for the oracle's building blocks (CPU)."""
import ctypes as C

import numpy as np
import pytest

from tests import helpers as H


def philox(lib, ctr, key):
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    lib.sqo_philox4x32_10(c, k, o)
    return list(o)


def test_philox4x32_10_known_answers(oracle_lib):
    # Random123 kat_vectors (philox4x32, 10 rounds)
    assert philox(oracle_lib, [0, 0, 0, 0], [0, 0]) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    assert philox(oracle_lib, [0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2) == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert philox(oracle_lib, [0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0]) == \
        [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


def philox_r(lib, ctr, key, rounds):
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    lib.sqo_philox4x32(c, k, rounds, o)
    return list(o)


def test_fma_rz_is_exact(oracle_lib):
    """sqo_fma_rz == the exactly computed x*y+z rounded toward zero to binary32 (PTX fma.rz.f32)."""
    from fractions import Fraction
    import struct

    def rz32(q):
        if q == 0:
            return 0.0
        sgn, a = (-1 if q < 0 else 1), abs(q)
        e = a.numerator.bit_length() - a.denominator.bit_length()
        if Fraction(2) ** e > a:
            e -= 1
        e = max(e, -126)
        m = int(a / Fraction(2) ** (e - 23))  # floor: toward zero
        return sgn * float(Fraction(m) * Fraction(2) ** (e - 23))

    rs = np.random.RandomState(5)
    xs = rs.standard_normal(3000).astype(np.float32)
    ys = (rs.uniform(0.5, 40, 3000)).astype(np.float32)
    zs = (rs.uniform(32768, 36000, 3000)).astype(np.float32)
    zs[::3] = (rs.uniform(-50, 2000, 1000)).astype(np.float32)
    zs[5::7] = np.round(zs[5::7])            # integers: the sum lands just below/above an integer
    xs[11::13] *= np.float32(1e-6)
    for x, y, z in zip(xs, ys, zs):
        want = rz32(Fraction(float(x)) * Fraction(float(y)) + Fraction(float(z)))
        got = oracle_lib.sqo_fma_rz(float(x), float(y), float(z))
        assert struct.pack("<f", got) == struct.pack("<f", want), (x, y, z)


def test_philox_matches_independent_python(oracle_lib):
    def ref(ctr, key, rounds=10):
        c, k = list(ctr), list(key)
        for _ in range(rounds):
            p0, p1 = 0xD2511F53 * c[0], 0xCD9E8D57 * c[2]
            c = [(p1 >> 32) ^ c[1] ^ k[0], p1 & 0xFFFFFFFF, (p0 >> 32) ^ c[3] ^ k[1], p0 & 0xFFFFFFFF]
            k = [(k[0] + 0x9E3779B9) & 0xFFFFFFFF, (k[1] + 0xBB67AE85) & 0xFFFFFFFF]
        return c
    rs = np.random.RandomState(1)
    for _ in range(200):
        ctr = [int(x) for x in rs.randint(0, 2 ** 32, 4, dtype=np.uint64)]
        key = [int(x) for x in rs.randint(0, 2 ** 32, 2, dtype=np.uint64)]
        assert philox(oracle_lib, ctr, key) == ref(ctr, key)
        assert philox_r(oracle_lib, ctr, key, 7) == ref(ctr, key, 7)


def test_lehmer_stream_is_minstd(oracle_lib):
    """reference src/rand.h:79-85: the n-th value is seed*16807^n mod (2^31-1), mapped to (0,1]"""
    m = 2147483647
    for seed in (1, 2, 5, 12345, 4097, 2 ** 31 + 7):
        st = C.c_int64(seed)
        x = seed % m
        for _ in range(2000):
            u = oracle_lib.sqo_lehmer_next(C.byref(st))
            x = (x * 16807) % m
            assert u == (x if x > 0 else m) / m


def test_lehmer_normal_matches_formula(oracle_lib):
    import math
    st, st2 = C.c_int64(77), C.c_int64(77)
    for _ in range(500):
        got = oracle_lib.sqo_lehmer_normal(C.byref(st), 10.0, 3.0)
        u = oracle_lib.sqo_lehmer_next(C.byref(st2))
        t = 2.0 * 3.14159265 * oracle_lib.sqo_lehmer_next(C.byref(st2))
        assert got == math.sqrt(-2.0 * math.log(u)) * math.cos(t) * 3.0 + 10.0


def test_kmer_ranks(oracle_lib):
    # reference src/seq.h:31-42: first base most significant; IUPAC folding; unknown -> 0
    assert oracle_lib.sqo_kmer_rank(b"AAAAAA", 6) == 0
    assert oracle_lib.sqo_kmer_rank(b"TTTTTT", 6) == 4095
    assert oracle_lib.sqo_kmer_rank(b"ACGTAC", 6) == int("012301", 4)
    assert oracle_lib.sqo_kmer_rank(b"acgtNU", 6) == int("012303", 4)
    assert oracle_lib.sqo_kmer_rank(b"RYKMSW", 6) == int("012020", 4)
    assert oracle_lib.sqo_kmer_rank(b"BDHVxX", 6) == int("100000", 4)
    assert oracle_lib.sqo_kmer_rank(b"ACGTACGTA", 9) == int("012301230", 4)
    # reference src/seq.h:62-74: A,C,G,M,T upper case only
    assert oracle_lib.sqo_meth_kmer_rank(b"ACGMTA", 6) == int("012340", 5)
    assert oracle_lib.sqo_meth_kmer_rank(b"acgmtN", 6) == 0
    assert oracle_lib.sqo_meth_kmer_rank(b"TTTTTTTTT", 9) == 5 ** 9 - 1


def test_short_read_rule(oracle_lib, ztable):
    """reference src/gensig.c:242-245: reads shorter than k become 5 k-mers of "ACGTACGTACGT"."""
    prof, flags = H.PRESETS["dna-r9-prom"]
    model = H.random_model(4096)
    o = H.Oracle(oracle_lib, prof, flags | H.SQ_IDEAL, 6, 4096, model, 1, H.RNG_LEGACY)
    a = o.gen_sig(b"AC", want_ss=True)
    b = o.gen_sig(b"ACGTACGTAC", want_ss=True)
    assert len(a["ss"]) == 5 and np.array_equal(a["sig"], b["sig"])
    o.close()
