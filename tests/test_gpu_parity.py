"""GPU parity: libsqg.so (through the C ABI) against the CPU oracle and the reference goldens.

Bar: bit-exact int16 signals, bit-exact per-read doubles, equal dwell arrays."""
import os

import numpy as np
import pytest

from tests import helpers as H

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sq():
    import squigulator_b200 as s
    s.load_library()
    return s


def run_pair(sq, oracle_lib, ztable, profile, flags, k, reads, seed=11, meth=False, amp_noise=1.0, first=0, want_ss=True,
             model=None, tuning=0):
    num_kmer = (5 if meth else 4) ** k
    model = H.random_model(num_kmer) if model is None else model
    gen = sq.SignalGenerator(dict(profile), model, k, flags=flags, seed=seed, meth=meth, amp_noise=amp_noise, tuning=tuning)
    got = gen.gen_batch(reads, first_read_index=first, want_ss=want_ss)
    gen.close()
    o = H.Oracle(oracle_lib, profile, flags, k, num_kmer, model, seed, H.RNG_PHILOX, meth=int(meth), amp_noise=amp_noise,
                 ztable=ztable)
    for i, read in enumerate(reads):
        exp = o.gen_sig(read, read_index=first + i, want_ss=want_ss)
        assert len(got[i]["sig"]) == len(exp["sig"]), f"read {i}: length {len(got[i]['sig'])} != {len(exp['sig'])}"
        if want_ss:
            np.testing.assert_array_equal(got[i]["ss"], exp["ss"], err_msg=f"read {i} dwell")
        bad = np.nonzero(got[i]["sig"] != exp["sig"])[0]
        assert bad.size == 0, f"read {i}: {bad.size} samples differ, first at {bad[:5]} got {got[i]['sig'][bad[:5]]} exp {exp['sig'][bad[:5]]}"
        assert got[i]["offset"] == exp["offset"] and got[i]["median_before"] == exp["median_before"], f"read {i} per-read draws"
    o.close()
    return got


CASES = [
    # name, preset, extra flags, k, meth, read alphabet, mean len, n
    ("dna_r9_prom", "dna-r9-prom", 0, 6, False, b"ACGT", 3000, 40),
    ("dna_r9_min", "dna-r9-min", 0, 6, False, b"ACGT", 2000, 20),
    ("dna_r10_prom", "dna-r10-prom", 0, 9, False, b"ACGT", 5000, 40),
    ("rna_r9_prom", "rna-r9-prom", 0, 5, False, b"ACGT", 1200, 30),
    ("rna004_prom", "rna004-prom", 0, 9, False, b"ACGT", 1300, 30),
    ("dna_ideal", "dna-r9-prom", H.SQ_IDEAL, 6, False, b"ACGT", 3000, 20),
    ("dna_ideal_time", "dna-r10-prom", H.SQ_IDEAL_TIME, 9, False, b"ACGT", 3000, 20),
    ("dna_ideal_amp", "dna-r10-prom", H.SQ_IDEAL_AMP, 9, False, b"ACGT", 3000, 20),
    ("rna_ideal", "rna004-prom", H.SQ_IDEAL, 9, False, b"ACGT", 1000, 20),
    ("rna_ideal_amp", "rna-r9-prom", H.SQ_IDEAL_AMP, 5, False, b"ACGT", 900, 20),
    ("r9_meth", "dna-r9-prom", 0, 6, True, b"ACGTM", 3000, 20),
    ("r10_meth", "dna-r10-prom", 0, 9, True, b"ACGTM", 3000, 12),
    ("iupac_lower_n", "dna-r9-prom", 0, 6, False, b"ACGTacgtNnRYKMSWBDHVU", 2000, 20),
    ("dna_prefix", "dna-r9-prom", H.SQ_PREFIX, 6, False, b"ACGT", 2000, 12),
    ("rna_prefix", "rna-r9-prom", H.SQ_PREFIX, 5, False, b"ACGT", 900, 12),
    ("rna004_prefix", "rna004-prom", H.SQ_PREFIX, 9, False, b"ACGT", 900, 12),
]


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_gpu_equals_oracle(case, sq, oracle_lib, ztable):
    name, preset, xflags, k, meth, alpha, mean_len, n = case
    prof, pflags = H.PRESETS[preset]
    reads = H.random_reads(n, mean_len, seed=hash(name) & 0xFFFF, alphabet=alpha)
    run_pair(sq, oracle_lib, ztable, prof, pflags | xflags, k, reads, meth=meth, first=1000)


REAL = [("dna-r9-prom", "dna-r9-prom", 6, False, b"ACGT"), ("dna-r10-prom", "dna-r10-prom", 9, False, b"ACGT"),
        ("rna004-prom", "rna004-prom", 9, False, b"ACGT"), ("rna-r9-prom", "rna-r9-prom", 5, False, b"ACGT"),
        ("dna-r9-prom-meth", "dna-r9-prom", 6, True, b"ACGTM"), ("dna-r10-prom-meth", "dna-r10-prom", 9, True, b"ACGTM")]


@pytest.mark.parametrize("case", REAL, ids=[c[0] for c in REAL])
def test_gpu_equals_oracle_with_the_reference_tables(case, sq, oracle_lib, ztable):
    """Philox mode with the reference's own built-in pore models (src/model.h, src/methmodel.c; dumped at build time by
    oracle/dump_models.py through the compiled reference) instead of a random table."""
    name, preset, k, meth, alpha = case
    model = H.real_model(name)
    if model is None:
        pytest.skip("oracle/_ref/models not built (needs /root/reference at build time)")
    prof, pflags = H.PRESETS[preset]
    reads = H.random_reads(24, 2500, seed=hash(name) & 0xFFFF, alphabet=alpha)
    run_pair(sq, oracle_lib, ztable, prof, pflags, k, reads, meth=meth, first=31, model=model)


def test_edge_cases(sq, oracle_lib, ztable):
    prof, pflags = H.PRESETS["dna-r10-prom"]
    # empty read, reads shorter than k (the "ACGTACGTACGT" rule, reference src/gensig.c:242-245), len == k, len == k+1
    reads = [b"", b"A", b"ACGTACG", b"ACGTACGTA", b"ACGTACGTAC", b"T" * 5000, b"ACGT" * 3000]
    run_pair(sq, oracle_lib, ztable, prof, pflags, 9, reads, first=7)
    prof, pflags = H.PRESETS["rna-r9-prom"]
    run_pair(sq, oracle_lib, ztable, prof, pflags, 5, [b"", b"AC", b"ACGTA", b"ACGTAC", b"G" * 3000], first=3)
    run_pair(sq, oracle_lib, ztable, prof, pflags | H.SQ_PREFIX, 5, [b"", b"AC", b"G" * 700], first=3)


def test_window_fallback_paths(sq, oracle_lib, ztable):
    """The signal kernel keeps a sliding window of k-mers per warp; a run of tiles is CUT when the next tile does not fit
    behind what the window still holds, and a tile that could never fit is generated sample by sample (slow path).
    Neither happens with sane profiles at default sizes, so the testing knob (sqg_config_t::reserved: low 16 bits =
    samples per tile the fast path accepts, high bits = k-mers the window holds) shrinks the window until they do."""
    reads = H.random_reads(10, 4000, seed=77) + [b"ACGT" * 2000, b"A" * 300]
    for preset, k, xflags in (("dna-r10-prom", 9, 0), ("dna-r9-prom", 6, 0), ("rna-r9-prom", 5, 0), ("rna004-prom", 9, 0),
                              ("rna-r9-prom", 5, H.SQ_PREFIX), ("dna-r10-prom", 9, H.SQ_IDEAL_AMP)):
        prof, pflags = H.PRESETS[preset]
        for tuning in ((262 << 16), (300 << 16) | 3000, 1500, (258 << 16) | 40000):
            run_pair(sq, oracle_lib, ztable, prof, pflags | xflags, k, reads, first=5, tuning=tuning)


def test_empty_batch(sq):
    gen = sq.SignalGenerator("dna-r9-prom", H.random_model(4096), 6)
    assert gen.gen_batch([]) == []
    gen.close()


def test_large_read_index_and_seed(sq, oracle_lib, ztable):
    prof, pflags = H.PRESETS["dna-r9-prom"]
    reads = H.random_reads(6, 1500, seed=5)
    run_pair(sq, oracle_lib, ztable, prof, pflags, 6, reads, seed=(1 << 40) + 12345, first=(1 << 33) + 5)


def test_dwell_extremes(sq, oracle_lib, ztable):
    # large spread: many folded (negative) draws; dwell_std = 0 with non-integer mean; --amp-noise scaling
    prof = dict(H.PRESETS["dna-r9-prom"][0])
    prof["dwell_mean"], prof["dwell_std"] = 4.0, 9.0
    run_pair(sq, oracle_lib, ztable, prof, 0, 6, H.random_reads(10, 1500, seed=9))
    prof["dwell_mean"], prof["dwell_std"] = 7.5, 0.0
    run_pair(sq, oracle_lib, ztable, prof, 0, 6, H.random_reads(10, 1500, seed=10))
    prof["dwell_mean"], prof["dwell_std"] = 120.0, 60.0
    run_pair(sq, oracle_lib, ztable, prof, 0, 6, H.random_reads(4, 800, seed=12), amp_noise=0.5)
    run_pair(sq, oracle_lib, ztable, H.PRESETS["dna-r10-min"][0], H.SQ_R10, 9, H.random_reads(6, 2500, seed=13), amp_noise=0.0)


def test_sample_range_paths(sq, oracle_lib, ztable):
    """The sample kernel reads the int16 out of the mantissa of fma.rz(z, A', B' + 32768) and sends everything outside
    [0, 32768) - negative values, values that wrap around int16, tail cells of the table - through its exact path;
    profiles it cannot bound at all run in `wide` mode.  All of them must equal the oracle bit for bit."""
    reads = H.random_reads(8, 2500, seed=31)
    base = dict(H.PRESETS["dna-r10-prom"][0])
    neg = dict(base, offset_mean=1500.0)                      # every sample negative (truncation toward zero, not floor)
    run_pair(sq, oracle_lib, ztable, neg, H.SQ_R10, 9, reads)
    straddle = dict(base, offset_mean=700.0, offset_std=5.0)  # samples on both sides of zero inside one chunk
    run_pair(sq, oracle_lib, ztable, straddle, H.SQ_R10, 9, reads)
    big = dict(base, digitisation=65536.0, range=140.0)       # beyond int16: the reference's 16-bit wrap (wide mode)
    run_pair(sq, oracle_lib, ztable, big, H.SQ_R10, 9, reads[:4])
    run_pair(sq, oracle_lib, ztable, base, H.SQ_R10, 9, reads[:4], amp_noise=900.0)   # wide mode by the noise scale
    edge = dict(base, digitisation=8192.0, range=30.0)        # values around 32768: fast path and exact path mixed
    run_pair(sq, oracle_lib, ztable, edge, H.SQ_R10, 9, reads[:4])


def test_svb_zd_streams(sq, oracle_lib):
    """SQG_WANT_SVB: the GPU's svb-zd stream of every read == the oracle's restatement of slow5lib's encoder run on the
    raw signal of the same read (which tests/test_svb_zd.py pins to the compiled slow5lib), and decodes back to it."""
    reads = H.random_reads(30, 2500, seed=41, min_len=0) + [b"", b"A", b"ACGTACGTAC"]
    model = H.random_model(4 ** 9)
    for prof, kw in (("dna-r10-prom", {}), ("dna-r10-prom", dict(flags=H.SQ_IDEAL_AMP)), ("rna004-prom", {})):
        gen = sq.SignalGenerator(prof, model, 9, seed=3, **kw)
        raw = gen.gen_batch(reads, first_read_index=50)
        svb = gen.gen_batch(reads, first_read_index=50, want_svb=True, want_ss=True)
        gen.close()
        for a, b in zip(raw, svb):
            assert "sig" not in b and b["n_samples"] == len(a["sig"])
            want = H.oracle_svb_zd(oracle_lib, a["sig"])
            assert np.array_equal(b["svb"], want), (len(a["sig"]), b["svb"][:12], want[:12])
            dec, used = H.svb_zd_decode(b["svb"])
            assert used == len(b["svb"]) and np.array_equal(dec, a["sig"])
    # extreme deltas (3-byte codes): the wrap-around profile of test_sample_range_paths
    big = dict(H.PRESETS["dna-r10-prom"][0], digitisation=65536.0, range=140.0)
    gen = sq.SignalGenerator(big, model, 9, flags=H.SQ_R10, seed=3)
    raw = gen.gen_batch(reads[:6]); svb = gen.gen_batch(reads[:6], want_svb=True)
    gen.close()
    for a, b in zip(raw, svb):
        assert np.array_equal(b["svb"], H.oracle_svb_zd(oracle_lib, a["sig"]))


def test_ss_text(sq, oracle_lib):
    """SQG_WANT_SS_TEXT: the `ss:Z:` value of every read, formatted on the GPU == the oracle's restatement of
    src/format.c:69-75 applied to the dwell array of the same read (DNA in k-mer order, RNA last k-mer first)."""
    reads = H.random_reads(20, 2500, seed=43, min_len=0) + [b"", b"A", b"ACGTACGTAC", b"ACGT" * 4000]
    for prof, k, kw in (("dna-r10-prom", 9, {}), ("rna-r9-prom", 5, {}), ("rna004-prom", 9, {}),
                        ("dna-r9-prom", 6, dict(flags=H.SQ_IDEAL_TIME))):
        gen = sq.SignalGenerator(prof, H.random_model(4 ** k), k, seed=9, **kw)
        a = gen.gen_batch(reads, first_read_index=5, want_ss=True)
        b = gen.gen_batch(reads, first_read_index=5, want_ss_text=True)
        rna = gen_is_rna = prof.startswith("rna")
        gen.close()
        for x, y in zip(a, b):
            assert np.array_equal(x["sig"], y["sig"])
            assert y["ss_text"] == H.oracle_ss_text(oracle_lib, x["ss"], rna), (prof, len(x["ss"]))
    # large dwells (3-4 digits; the library caps mean + 6 std at 1200 samples per k-mer)
    prof = dict(H.PRESETS["dna-r9-prom"][0], dwell_mean=850.0, dwell_std=55.0)
    gen = sq.SignalGenerator(prof, H.random_model(4096), 6, seed=9)
    a = gen.gen_batch(reads[:4], want_ss=True); b = gen.gen_batch(reads[:4], want_ss_text=True, want_svb=True)
    gen.close()
    for x, y in zip(a, b):
        assert y["ss_text"] == H.oracle_ss_text(oracle_lib, x["ss"], False)
        assert np.array_equal(y["svb"], H.oracle_svb_zd(oracle_lib, x["sig"]))


def test_random_profiles(sq, oracle_lib, ztable):
    """Seeded sweep over the parameter space the fixed cases do not visit: random ADC scaling, offsets (incl. signals
    around zero and near the int16 limits), noise scales up to the wide-mode threshold and beyond, dwell laws from 1-2
    samples per k-mer (many k-mers per 8-sample chunk) to hundreds (map capacity), every k, DNA and RNA order, prefix
    junctions, IUPAC letters.  GPU == oracle bit for bit for each."""
    rs = np.random.RandomState(20261017)
    for it in range(60):
        k = int(rs.choice([5, 6, 9]))
        rna = bool(rs.randint(0, 2))
        meth = (not rna) and k != 5 and rs.rand() < 0.2     # base-5 (CpG) model
        flags = (H.SQ_RNA if rna else 0) | (H.SQ_R10 if k == 9 else 0)
        if rs.rand() < 0.25:
            flags |= H.SQ_PREFIX
        if rs.rand() < 0.15:
            flags |= int(rs.choice([H.SQ_IDEAL_TIME, H.SQ_IDEAL_AMP, H.SQ_IDEAL]))
        dm = float(rs.choice([1.0, 2.0, 3.5, 9.0, 13.0, 31.0, 80.0, 400.0]))
        ds = float(rs.choice([0.0, 0.5, 1.0, 4.0, 0.4 * dm]))
        ds = min(ds, (1150.0 - dm) / 6.0)
        dig = float(rs.choice([2048.0, 8192.0, 1024.0]))
        rng_ = float(rs.uniform(150.0, 1500.0))
        prof = dict(digitisation=dig, sample_rate=4000.0, bps=400.0, range=rng_,
                    offset_mean=float(rs.choice([-250.0, 10.0, 900.0 * rng_ / dig * -1, 120.0 * dig / rng_])), offset_std=float(rs.uniform(0, 20)),
                    median_before_mean=200.0, median_before_std=float(rs.uniform(0, 20)), dwell_mean=dm, dwell_std=ds)
        amp = float(rs.choice([0.0, 0.3, 1.0, 1.0, 3.0, 40.0, 2000.0]))
        alpha = b"ACGT" if rs.rand() < 0.7 else b"ACGTacgtNRYKMSWU"
        if meth:
            alpha = b"ACGTM" if rs.rand() < 0.7 else b"ACGTMacgtN"
        n_reads = 6 if dm < 100 else 3
        mean_len = int(rs.choice([30, 300, 2500])) if dm < 100 else 400
        reads = H.random_reads(n_reads, mean_len, seed=int(rs.randint(1 << 30)), alphabet=alpha, min_len=0)
        run_pair(sq, oracle_lib, ztable, prof, flags, k, reads, seed=int(rs.randint(1, 1 << 40)), amp_noise=amp,
                 first=int(rs.randint(0, 1 << 35)), meth=meth)


def test_c_caller(sq, tmp_path):
    """The boundary from plain C99 (tests/c_abi_demo.c): sqg_init, the gen_sig-shaped call, one batch with svb-zd and
    ss:Z: text; its counts agree with the Python binding for the same read, model and seed."""
    import subprocess
    from tests.test_abi import build_c_demo
    r = subprocess.run([build_c_demo(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.startswith("ok "), (r.returncode, r.stdout, r.stderr)
    n_samples, n_kmers, svb_bytes, text_bytes = map(int, r.stdout.split()[1:5])
    read = b"ACGTTGCATGCATGCAAACCCGGGTTTACGATCGATCGATTAGCTAGCTAGGATCGATCGGCTAGCTAGCATCGACTGACTAGCTAGCATCGATCGA"
    model = np.empty(2 * 4096, dtype=np.float32)
    model[0::2] = 60.0 + (np.arange(4096) % 70)
    model[1::2] = 1.0 + (np.arange(4096) % 3)
    gen = sq.SignalGenerator(dict(H.PRESETS["dna-r9-prom"][0]), model, 6, seed=1)
    g = gen.gen_batch([read], want_ss=True)[0]
    v = gen.gen_batch([read], want_svb=True, want_ss_text=True)[0]
    gen.close()
    assert (n_samples, n_kmers, svb_bytes, text_bytes) == (len(g["sig"]), len(g["ss"]), len(v["svb"]), len(v["ss_text"]))


def test_batch_split_and_api_variants_agree(sq):
    """Output depends only on (seed, global read index, bases): not on batching, nor on which entry point is used."""
    reads = H.random_reads(24, 2500, seed=21, min_len=0)
    model = H.random_model(4 ** 9)
    gen = sq.SignalGenerator("dna-r10-prom", model, 9, seed=5, n_slots=2)
    whole = gen.gen_batch(reads, first_read_index=100, want_ss=True)
    a = gen.gen_batch(reads[:9], first_read_index=100, want_ss=True)
    b = gen.gen_batch(reads[9:], first_read_index=109, want_ss=True)
    for x, y in zip(whole, a + b):
        np.testing.assert_array_equal(x["sig"], y["sig"])
        np.testing.assert_array_equal(x["ss"], y["ss"])
        assert x["offset"] == y["offset"] and x["median_before"] == y["median_before"]
    # per-read drop-in (gen_sig shape)
    for i in (0, 5, 23):
        r = gen.gen_sig(reads[i], read_index=100 + i, want_ss=True)
        np.testing.assert_array_equal(r["sig"], whole[i]["sig"])
        np.testing.assert_array_equal(r["ss"], whole[i]["ss"])
        assert r["offset"] == whole[i]["offset"]
    # asynchronous dispatcher: three batches in flight over two slots
    from squigulator_b200.api import _pack_reads, WANT_SS
    parts = [(reads[:8], 100), (reads[8:16], 108), (reads[16:], 116)]
    packed = [(_pack_reads(r), f) for r, f in parts]
    tickets = [gen.submit(p[0], p[1], first_read_index=f, want=WANT_SS) for p, f in packed[:2]]
    got = []
    res = gen.wait(tickets[0]); got += gen._unpack(res); gen.release(tickets[0])
    tickets.append(gen.submit(packed[2][0][0], packed[2][0][1], first_read_index=packed[2][1], want=WANT_SS))
    for t in tickets[1:]:
        res = gen.wait(t); got += gen._unpack(res); gen.release(t)
    for x, y in zip(whole, got):
        np.testing.assert_array_equal(x["sig"], y["sig"])
    # device-resident batch (what bench.py times)
    (bases, off) = _pack_reads(reads)
    db = gen.dev_batch(bases, off, first_read_index=100)
    gen.dev_batch_run(db, 2)
    for x, y in zip(whole, gen.dev_batch_fetch(db)):
        np.testing.assert_array_equal(x["sig"], y["sig"])
    assert gen.dev_batch_info(db)["samples"] == sum(len(x["sig"]) for x in whole)
    gen.dev_batch_destroy(db)
    assert gen.launch_count() > 0
    gen.close()


@pytest.mark.parametrize("name", ["dna_ideal"])
def test_gpu_reproduces_reference_ideal_golden(name, sq):
    """--ideal uses no random numbers, so the GPU output must equal the reference's own golden file
    (reference test/dna_ideal_slow5.exp via tests/golden/dna_ideal.npz) bit for bit."""
    g = H.Golden(os.path.join(H.GOLDEN_DIR, name + ".npz"))
    c = g.cfg
    gen = sq.SignalGenerator(c["profile"], g.dense_model(), c["kmer_size"], flags=c["flags"], seed=c["seed"],
                             meth=bool(c["meth"]), amp_noise=c["amp_noise"])
    got = gen.gen_batch(g.reads)
    gen.close()
    for i in range(len(g.reads)):
        assert len(got[i]["sig"]) == g.sig_len[i]
        assert H.sha256_i16(got[i]["sig"]) == g.sha_of(i)
        if i < g.n_full:
            np.testing.assert_array_equal(got[i]["sig"], g.sig_full[i])
        assert abs(got[i]["offset"] - g.offset[i]) < 1e-6 and abs(got[i]["median_before"] - g.median_before[i]) < 1e-6


def test_golden_structure_philox(sq):
    """Goldens with one noise source off pin structure even under Philox: --ideal-time fixes every length."""
    g = H.Golden(os.path.join(H.GOLDEN_DIR, "dna_ideal_time.npz"))
    c = g.cfg
    gen = sq.SignalGenerator(c["profile"], g.dense_model(), c["kmer_size"], flags=c["flags"], seed=c["seed"])
    got = gen.gen_batch(g.reads)
    gen.close()
    assert [len(x["sig"]) for x in got] == list(g.sig_len)
