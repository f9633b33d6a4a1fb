"""The amplitude law of the Philox scheme, tested as a LAW: what one draw can produce, that no
sample position is tied to a subset of it, and that a long run of samples is indistinguishable - by Kolmogorov-Smirnov,
fourth moment and tail counts - from the reference's own rand.h path (src/rand.h:87-94) at >= 10^7 samples.

A draw = 10 random bits (sign + 9-bit slot) and a CLASS in 0..31; within a class the 1024 values are equiprobable atoms
(conditional RMS of 512 half-normal cells, scaled to unit variance; the outermost cell of class 0 is refined by 13 more
bits).  The class of the sample at emitted position q is  (chunk & 31) XOR hash(unit, group, read)  with chunk = q >> 3 (a unit = 3 groups of 32 chunks):
inside a group of 32 chunks every class occurs once (one shared-memory bank per lane), across groups a position meets
all 32 classes, so the law of a sample at ANY position is the pooled law: 32768 atoms, |z| up to 5.91.
"""
import ctypes as C
import os

import numpy as np
import pytest

from tests import helpers as H

N1 = 16384
REF_SO = os.path.join(H.ROOT, "oracle", "_ref", "libsqref.so")


def class_atoms(z32, z2, c):
    """(values, probabilities) of |z| within class c"""
    v = z32[c:N1:32].astype(np.float64)
    w = np.full(512, 1.0 / 512)
    if c == 0:
        v = np.concatenate([v[:511], z2.astype(np.float64)])
        w = np.concatenate([w[:511], np.full(z2.size, 1.0 / 512 / z2.size)])
    return v, w


def test_per_class_and_pooled_law_from_the_table():
    from scipy.stats import norm
    raw = H.load_ztable()
    z32, z2 = raw[:32768 * 4].view("<f4"), raw[32768 * 4:].view("<f4")
    kurt, p3, p4, zmax = [], [], [], []
    for c in range(32):
        v, w = class_atoms(z32, z2, c)
        m2, m4 = (w * v ** 2).sum(), (w * v ** 4).sum()
        assert abs(m2 - 1.0) < 2e-7
        kurt.append(m4 / m2 ** 2)
        p3.append(w[v > 3].sum())
        p4.append(w[v > 4].sum())
        zmax.append(v.max())
    # one class by itself is a coarse law: kurtosis 2.95 .. 3.28, no atom beyond 3.1 sigma in the lightest class ...
    assert 2.9 < min(kurt) and max(kurt) < 3.3
    assert 3.0 < min(zmax) < 3.2 and abs(max(zmax) - 5.9104958) < 1e-6
    # ... which is why the class must not be a function of the sample position alone (see test_class_rotation_*): the
    # POOLED law is the law of every sample, and it is N(0,1) to 5e-4 in kurtosis, 2 % in P(|z|>3), 10 % in P(|z|>4)
    assert abs(np.mean(kurt) - 3.0) < 1e-3
    assert abs(np.mean(p3) / (2 * norm.sf(3)) - 1.0) < 0.02
    assert abs(np.mean(p4) / (2 * norm.sf(4)) - 1.0) < 0.11


def amp_class(q, r_lo):
    """class of the sample at emitted position q of read r_lo (oracle/sqg_oracle.c, sqg_signal.cuh amp_class4)"""
    cq = q >> 3
    u, g = cq // 96, (cq % 96) >> 5
    x = ((u * 0x9E3779B1) + ((r_lo * 0x85EBCA6B) & 0xFFFFFFFF)) & 0xFFFFFFFF
    x ^= x >> 15
    h = (((x * 0x2C1B3C6D) & 0xFFFFFFFF) >> (27 - 5 * g)) & 31
    return (cq & 31) ^ h


def test_class_rotation_covers_every_class_at_every_position():
    q = np.arange(0, 8 * 32 * 4096, 8, dtype=np.int64)          # first sample of 131072 chunks = 4096 groups
    for r_lo in (0, 1, 12345, 0xFFFFFFFF):
        c = amp_class(q, r_lo).reshape(4096, 32)
        assert np.all(np.sort(c, axis=1) == np.arange(32))      # inside a group: a bijection of the lanes (one bank each)
        for lane in (0, 7, 31):                                  # a fixed position within the group, over the groups:
            cnt = np.bincount(c[:, lane], minlength=32)          # uniform over the 32 classes (chi-square, 31 dof)
            chi2 = ((cnt - 128.0) ** 2 / 128.0).sum()
            assert chi2 < 31 + 5 * np.sqrt(62), (r_lo, lane, chi2)
        # and not periodic: the class at a position repeats 1, 2, 3 groups (256, 512, 768 samples) later with probability
        # 1/32 (all positions of a group move together - class = lane ^ hash - so a group is ONE observation)
        for lag in (1, 2, 3):
            same = (c[lag:, 0] == c[:-lag, 0]).mean()
            assert abs(same - 1 / 32) < 5 * np.sqrt((1 / 32) * (31 / 32) / (4096 - lag)), (r_lo, lag, same)


def flat_model(num_kmer, mean=95.0, stdv=2.5):
    m = np.empty(2 * num_kmer, dtype=np.float32)
    m[0::2], m[1::2] = mean, stdv
    return m


def residuals(sig, offset, prof, mean, stdv, dither=None):
    """standardised residual of every sample of a flat-model read: ((raw + 0.5 + offset) * range / dig - mean) / stdv
    (+0.5: centre of the truncation cell; the truncation leaves a uniform error of +-0.5 ADC steps = +-0.03 sigma here).
    dither: a RandomState - the +0.5 becomes U(0,1), which turns the per-read lattice of ADC steps (0.055 sigma wide, at
    a different phase in every read) into a continuous variable, the same way on both sides of a two-sample test."""
    d = 0.5 if dither is None else dither.random_sample(sig.size)
    return ((sig.astype(np.float64) + d + offset) * prof["range"] / prof["digitisation"] - mean) / stdv


def law_checks(z, label):
    """moments and tail counts of n standardised residuals against N(0,1) + the uniform truncation term"""
    from scipy.stats import norm
    n = z.size
    q2 = np.var(z) - 1.0                                          # truncation adds (step/stdv)^2 / 12 ~ 3e-4 .. 2e-3
    assert abs(z.mean()) < 5 / np.sqrt(n), (label, z.mean())
    assert -1e-3 < q2 < 4e-3, (label, np.var(z))
    k = np.mean((z - z.mean()) ** 4) / np.var(z) ** 2
    assert abs(k - 3.0) < 5 * np.sqrt(24.0 / n) + 2e-3, (label, k)
    for t in (3.0, 4.0):
        p = 2 * norm.sf(t)
        cnt = np.count_nonzero(np.abs(z) > t)
        # the table's pooled tail mass sits within 2 % (3 sigma) / 11 % (4 sigma) of the normal one: see the table test
        tol = (0.03 if t == 3.0 else 0.13) * p * n + 5 * np.sqrt(p * n)
        assert abs(cnt - p * n) < tol, (label, t, cnt, p * n)


def lag_checks(z, label):
    """|residual| > 3 events must not recur with the period of the group structure (256 samples) or of the chunk (8)"""
    e = (np.abs(z) > 3.0).astype(np.float64)
    p = e.mean()
    for lag in (8, 256, 768):
        both = np.mean(e[lag:] * e[:-lag])
        sd = np.sqrt(p * p * (1 - p * p) / (e.size - lag))
        assert abs(both - p * p) < 5 * sd, (label, lag, both, p * p)
    # and their positions are uniform over the 256 positions of a group (chi-square, 255 dof)
    pos = np.nonzero(e)[0] & 255
    cnt = np.bincount(pos, minlength=256)
    chi2 = ((cnt - cnt.mean()) ** 2 / cnt.mean()).sum()
    assert chi2 < 255 + 5 * np.sqrt(510), (label, chi2)


def ref_residuals(n_samples, prof, flags, k, mean, stdv, seed, dither=None):
    """>= n_samples residuals of the reference's own nrng path (compiled, unmodified reference): a homopolymer read, so
    that ONE k-mer stream supplies all the draws and the small-seed artefact of its first draw does not matter"""
    from tests.test_oracle_vs_ref import ref_gen
    lib = C.CDLL(REF_SO)
    lib.sqref_open.restype = C.c_void_p
    lib.sqref_open.argtypes = [C.POINTER(H.Profile), C.c_uint32, C.c_int64, C.c_int32, C.c_float, C.c_int, C.c_char_p, C.c_char_p, C.c_int]
    lib.sqref_gen_sig.restype = C.c_int64
    lib.sqref_gen_sig.argtypes = [C.c_void_p, C.c_char_p, C.c_int32, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double),
                                  C.POINTER(C.POINTER(C.c_int16)), C.POINTER(C.POINTER(C.c_int32)), C.POINTER(C.c_int64)]
    lib.sqref_free_buf.argtypes = [C.c_void_p]
    lib.sqref_close.argtypes = [C.c_void_p]
    import tempfile
    # the reference takes its table from a model file (--kmer-model, src/model.c:40-142): a flat one
    with tempfile.NamedTemporaryFile("w", suffix=".model", delete=False) as f:
        f.write(f"#model_name\tflat\n#k\t{k}\nkmer\tlevel_mean\tlevel_stdv\tsd_mean\tsd_stdv\tweight\n")   # ("#k": src/model.c:75-91)
        import itertools
        for t in itertools.product("ACGT", repeat=k):
            f.write(f"{''.join(t)}\t{mean:.6f}\t{stdv:.6f}\t0\t0\t0\n")
        path = f.name
    p = H.make_profile(prof)
    h = lib.sqref_open(C.byref(p), flags, seed, 1, 1.0, 0, path.encode(), None, 0)
    out, got = [], 0
    while got < n_samples:
        r = ref_gen(lib, h, b"A" * 200000)
        out.append(residuals(r["sig"], r["offset"], prof, mean, stdv, dither))
        got += out[-1].size
    lib.sqref_close(h)
    os.unlink(path)
    return np.concatenate(out)


def test_oracle_law_moments_tails_and_periodicity(oracle_lib, ztable):
    """CPU: 4 x 10^6 samples of the oracle's Philox path (the GPU test below runs the same checks on 4 x 10^7)"""
    prof, flags = H.PRESETS["dna-r10-prom"]
    o = H.Oracle(oracle_lib, prof, flags, 9, 4 ** 9, flat_model(4 ** 9), 5, H.RNG_PHILOX, ztable=ztable)
    z = []
    for i in range(12):
        r = o.gen_sig(b"ACGT" * 6500, read_index=1000 + i)
        z.append(residuals(r["sig"], r["offset"], prof, 95.0, 2.5))
    o.close()
    z = np.concatenate(z)
    assert z.size > 3.5e6
    law_checks(z, "oracle")
    lag_checks(z, "oracle")


@pytest.mark.skipif(not os.path.exists(REF_SO), reason="oracle/_ref not built")
def test_oracle_vs_reference_ks(oracle_lib, ztable):
    """Two-sample Kolmogorov-Smirnov, 10^7 samples each: oracle Philox path against the reference's rand.h path.  Both
    sides carry the same ADC truncation, so the statistic compares the generators."""
    from scipy.stats import ks_2samp
    prof, flags = H.PRESETS["dna-r10-prom"]
    zr = ref_residuals(10_000_000, prof, flags, 9, 95.0, 2.5, seed=4, dither=np.random.RandomState(1))
    o = H.Oracle(oracle_lib, prof, flags, 9, 4 ** 9, flat_model(4 ** 9), 5, H.RNG_PHILOX, ztable=ztable)
    z, rs = [], np.random.RandomState(2)
    for i in range(31):
        r = o.gen_sig(b"ACGT" * 6250, read_index=i)
        z.append(residuals(r["sig"], r["offset"], prof, 95.0, 2.5, rs))
    o.close()
    z = np.concatenate(z)[:zr.size]
    assert z.size >= 9_500_000
    ks = ks_2samp(z, zr[:z.size])
    # D ~ 1.36 * sqrt(2/n) = 6e-4 at the 5 % level; require the 0.1 % level (1.95 * sqrt(2/n))
    assert ks.statistic < 1.95 * np.sqrt(2.0 / z.size), ks


def test_parameters_in_use_match_the_reference_expression_to_1e_4():
    """north_star: per-k-mer current mean / stdv within 1e-4 relative of the reference's.  The parameters the kernels use
    are binary32 roundings of the reference's double expressions (src/gensig.c:266,270; src/sim.c:249):
    A' = (stdv*amp_noise)*scale, M = mean*scale, B = M - offset on the 2^-8 grid."""
    model = H.real_model("dna-r10-prom")
    if model is None:
        model = H.random_model(4 ** 9)
    prof = H.PRESETS["dna-r10-prom"][0]
    scale = prof["digitisation"] / prof["range"]
    mean, stdv = model[0::2].astype(np.float64), model[1::2].astype(np.float64)
    A = (model[1::2] * np.float32(1.0)) * np.float32(scale)
    M = model[0::2] * np.float32(scale)
    assert np.max(np.abs(A.astype(np.float64) - stdv * scale) / (stdv * scale)) < 2e-7
    assert np.max(np.abs(M.astype(np.float64) - mean * scale) / (mean * scale)) < 2e-7
    for off in (prof["offset_mean"], prof["offset_mean"] + 12 * prof["offset_std"], prof["offset_mean"] - 12 * prof["offset_std"]):
        c_r = np.float32(32768.0) - np.float32(off)
        Bq = (M + c_r) - np.float32(32768.0)                                   # what the sample arithmetic adds to z*A'
        exact = mean * scale - off
        assert np.max(np.abs(Bq.astype(np.float64) - exact)) < 2 ** -8 + 1e-4  # the 2^-8 grid: 0.004 of an ADC step
        assert np.max(np.abs(Bq.astype(np.float64) - exact) / np.abs(exact)) < 1e-4


# ---- the same law checks through the GPU, at 10x the size ----

@pytest.mark.gpu
def test_gpu_law_moments_tails_periodicity_and_ks(oracle_lib, ztable):
    import squigulator_b200 as sq
    from scipy.stats import ks_2samp
    prof, flags = H.PRESETS["dna-r10-prom"]
    gen = sq.SignalGenerator(dict(prof), flat_model(4 ** 9), 9, flags=flags, seed=5)
    got = gen.gen_batch([b"ACGT" * 25000] * 31, first_read_index=0)
    gen.close()
    z = np.concatenate([residuals(r["sig"], r["offset"], prof, 95.0, 2.5) for r in got])
    assert z.size > 3.9e7
    law_checks(z, "gpu")
    lag_checks(z, "gpu")
    if os.path.exists(REF_SO):
        zr = ref_residuals(10_000_000, prof, flags, 9, 95.0, 2.5, seed=4, dither=np.random.RandomState(1))
        rs = np.random.RandomState(2)
        zd = np.concatenate([residuals(r["sig"], r["offset"], prof, 95.0, 2.5, rs) for r in got[:8]])
        ks = ks_2samp(zd[:zr.size], zr)
        assert ks.statistic < 1.95 * np.sqrt(2.0 / zr.size), ks
