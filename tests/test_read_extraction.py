"""Reads named by coordinates: the sequence work of the reference's gen_read() for accepted reads —
copy, 'N' replacement, reverse complement, CpG -> 'M' marking (src/genread.c:125-281, src/seq.h:77-112).

CPU: the oracle's restatement against golden vectors made from the UNMODIFIED reference (tests/golden/gen_read.json,
scripts/make_golden_reads.py) and, where oracle/_ref is built, against the reference itself on fresh draws.
GPU: sqg_gen_batch_coords against the oracle, byte for byte, and its signal against the host-bases path."""
import ctypes as C
import hashlib
import json
import os
import sys

import numpy as np
import pytest

from tests import helpers as H

SO = os.path.join(H.ROOT, "oracle", "_ref", "libsqref.so")


@pytest.fixture(scope="module")
def oracle_lib():
    return H.load_oracle()


def meth_state_for(seed):
    return C.c_int64(seed + 6)  # rand_meth of thread 0, src/sim.c:253


def check_against(oracle_lib, reads, contigs, marr, seed, meth):
    """reads: (contig, pos, len, strand, expected bytes or sha256 hex); returns total rand_meth draws"""
    st = meth_state_for(seed) if meth else None
    total = 0
    for i, (c, pos, ln, strand, want) in enumerate(reads):
        got, draws = H.oracle_extract_read(oracle_lib, contigs[c], marr[c] if meth else None, pos, ln, strand, st)
        total += draws
        if isinstance(want, bytes):
            assert got == want, f"read {i}: {(c, pos, ln, strand)}"
        else:
            assert hashlib.sha256(got).hexdigest() == want, f"read {i}: {(c, pos, ln, strand)}"
    return total


def test_oracle_equals_golden_reads(oracle_lib):
    doc = json.load(open(os.path.join(H.GOLDEN_DIR, "gen_read.json")))
    assert len(doc) == 3
    for case in doc:
        contigs, marr = H.synthetic_genome(seed=case["genome_seed"], with_meth=True)
        assert hashlib.sha256(b"".join(contigs)).hexdigest() == case["genome_sha256"]
        reads = [(r["contig"], r["pos"], r["len"], r["strand"], r["sha256"]) for r in case["reads"]]
        draws = check_against(oracle_lib, reads, contigs, marr, case["seed"], case["meth"])
        if case["meth"]:
            assert draws > 1000  # the marking really ran (and stayed in step over 60 reads: one slip breaks every later read)
        assert {r["strand"] for r in case["reads"]} == ({"+"} if case["name"].startswith("rna") else {"+", "-"})


@pytest.mark.skipif(not os.path.exists(SO), reason="oracle/_ref/libsqref.so not built")
@pytest.mark.parametrize("preset,seed,meth", [("dna-r10-prom", 5, True), ("dna-r9-prom", 99, False), ("rna-r9-prom", 17, False)])
def test_oracle_equals_reference_gen_read(oracle_lib, preset, seed, meth):
    sys.path.insert(0, os.path.join(H.ROOT, "scripts"))
    import make_golden_reads as G
    contigs, marr = H.synthetic_genome(seed=seed + 100, n_contigs=4, mean_len=2500)
    reads = G.reference_reads(G.load_ref(), preset, seed, meth, contigs, marr, 80)
    check_against(oracle_lib, reads, contigs, marr, seed, meth)


def test_lehmer_jump(oracle_lib):
    for seed in (0, 1, 7, 48, 127773, 2147483646, 2147483647):
        st = C.c_int64(seed)
        seq = [oracle_lib.sqo_lehmer_next(C.byref(st)) for _ in range(300)]
        for n in (0, 1, 2, 31, 32, 33, 255, 298):
            j = C.c_int64(oracle_lib.sqo_lehmer_jump(seed, n))
            assert oracle_lib.sqo_lehmer_next(C.byref(j)) == seq[n], (seed, n)


def test_extraction_edge_cases(oracle_lib):
    ctg = b"NNCGNACGTnCGCGN"
    m = np.full(len(ctg), 255, dtype=np.uint8)
    # whole contig, forward: every upper-case N replaced from the stream seeded 100; both sites after position 0 marked
    st = C.c_int64(1 + 6)
    got, draws = H.oracle_extract_read(oracle_lib, ctg, m, 0, len(ctg), "+", st)
    assert len(got) == len(ctg) and b"N" not in got and got[9:10] == b"n"
    assert draws == ctg.count(b"CG") == 4 and got.count(b"M") == 4
    # a site cut by the read's end is not a site (i+1 < rlen, src/genread.c:211)
    got, draws = H.oracle_extract_read(oracle_lib, ctg, m, 0, 3, "+", C.c_int64(7))
    assert draws == 0 and got[2:3] == b"C"
    # reverse strand: M lands on the complement of the G
    got, draws = H.oracle_extract_read(oracle_lib, b"AACGTT", np.full(6, 255, np.uint8), 0, 6, "-", C.c_int64(7))
    assert got == b"AAMGTT" and draws == 1
    # empty read
    got, draws = H.oracle_extract_read(oracle_lib, ctg, m, 4, 0, "-", C.c_int64(7))
    assert got == b"" and draws == 0


# ---------------------------------------------------------------------------------------------- GPU

def _coords(contigs, n, seed, min_len=1):
    rs = np.random.RandomState(seed)
    out = []
    for _ in range(n):
        c = int(rs.randint(0, len(contigs)))
        L = len(contigs[c])
        ln = int(min(L, max(min_len, rs.gamma(2.0, 400))))
        pos = int(rs.randint(0, L - ln + 1))
        out.append((c, pos, ln, "+-"[int(rs.randint(0, 2))]))
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("meth", [False, True])
def test_coords_batch_equals_oracle_and_host_path(oracle_lib, meth):
    import squigulator_b200 as sq
    from squigulator_b200 import api
    k = 6
    nk = (5 if meth else 4) ** k
    model = H.random_model(nk, seed=5)
    seed = 21
    contigs, marr = H.synthetic_genome(seed=31, n_contigs=5, mean_len=4000)
    coords = _coords(contigs, 300, seed=8) + [(0, 0, len(contigs[0]), "-"), (1, 5, 0, "+"), (2, len(contigs[2]) - 1, 1, "-"),
                                              (3, 0, 33, "+"), (3, 1, 32, "-"), (4, 7, 31, "-")]
    # whole contigs and piece-boundary lengths on both strands (reads are cut into 2048-position pieces on the GPU)
    coords += [(c, 0, len(contigs[c]), st) for c in range(5) for st in "+-"]
    coords += [(c, 3, ln, st) for c in (0, 2) for ln in (2047, 2048, 2049, 2176) for st in "+-" if ln + 3 <= len(contigs[c])]
    g = sq.SignalGenerator("dna-r9-prom", model, k, seed=seed, meth=meth)
    g.load_genome(contigs, meth=marr)
    out, draws = g.gen_batch_coords(coords, first_read_index=1000, want=api.WANT_BASES | api.WANT_SS)
    # bytes == oracle, read by read, the rand_meth stream running across the batch
    st = meth_state_for(seed) if meth else None
    total = 0
    for i, (c, pos, ln, strand) in enumerate(coords):
        want, d = H.oracle_extract_read(oracle_lib, contigs[c], marr[c] if meth else None, pos, ln, strand, st)
        total += d
        assert out[i]["bases"] == want, f"read {i}: {coords[i]}"
    assert draws == total and (total > 500 if meth else total == 0)
    # the signal is what the host-bases path gives for those reads
    ref = g.gen_batch([o["bases"] for o in out], first_read_index=1000, want_ss=True)
    for a, b in zip(out, ref):
        np.testing.assert_array_equal(a["sig"], b["sig"])
        np.testing.assert_array_equal(a["ss"], b["ss"])
        assert a["offset"] == b["offset"]
    # two batches chained by meth_draws == one batch (addressable draws)
    h1, d1 = g.gen_batch_coords(coords[:117], first_read_index=1000, meth_draw_base=0, want=api.WANT_BASES)
    h2, d2 = g.gen_batch_coords(coords[117:], first_read_index=1117, meth_draw_base=d1, want=api.WANT_BASES)
    assert d1 + d2 == draws
    for a, b in zip(h1 + h2, out):
        assert a["bases"] == b["bases"]
        np.testing.assert_array_equal(a["sig"], b["sig"])
    # asynchronous submission gives the same
    arr = g.pack_coords(coords)
    t = g.submit_coords(arr, first_read_index=1000, meth_draw_base=0, want=api.WANT_BASES)
    res = g.wait(t)
    got = g._unpack(res)
    assert int(res.meth_draws) == draws
    for a, b in zip(got, out):
        assert a["bases"] == b["bases"]
        np.testing.assert_array_equal(a["sig"], b["sig"])
    g.release(t)
    g.close()


@pytest.mark.gpu
def test_coords_pieces_without_n_take_the_vector_path(oracle_lib):
    """pieces without N of reads without methylation are copied / reverse-complemented 16 bytes per lane: every
    alignment of source and destination, lengths around the 16-byte and 2048-position boundaries, lower case and IUPAC
    letters (complement: T, src/seq.h:77-101), next to pieces that do hold N's (the per-byte path) in the same batch"""
    import squigulator_b200 as sq
    from squigulator_b200 import api
    rs = np.random.RandomState(77)
    ln = 30000
    a = np.frombuffer(b"ACGT", dtype=np.uint8)[rs.randint(0, 4, ln)].copy()
    a[1000:1400] |= 0x20
    a[rs.randint(0, ln, 40)] = np.frombuffer(b"RYKMSWBDHVUacgtnryk", dtype=np.uint8)[rs.randint(0, 19, 40)]
    a[20000:20010] = ord("N")   # only pieces covering these take the per-byte path
    b = np.frombuffer(b"ACGT", dtype=np.uint8)[rs.randint(0, 4, 5000)].copy()
    contigs = [a.tobytes(), b.tobytes()]
    coords = []
    for pos in list(range(0, 20)) + [4093, 4094, 4095, 4096, 4097]:
        for n in (0, 1, 15, 16, 17, 31, 32, 33, 47, 100, 2047, 2048, 2049, 2063, 2064, 2065, 4096, 5000):
            for st in "+-":
                coords.append((0, pos, n, st))
    coords += [(0, 0, ln, "+"), (0, 0, ln, "-"), (0, 13, ln - 13, "-"), (1, 0, 5000, "-"), (1, 1, 4999, "+"), (0, 19000, 3000, "-"), (0, 19990, 25, "+")]
    g = sq.SignalGenerator("dna-r9-prom", H.random_model(4 ** 6, seed=5), 6, seed=3)
    g.load_genome(contigs)
    out, draws = g.gen_batch_coords(coords, want=api.WANT_BASES)
    assert draws == 0
    for i, (c, pos, n, strand) in enumerate(coords):
        want, _ = H.oracle_extract_read(oracle_lib, contigs[c], None, pos, n, strand, None)
        assert out[i]["bases"] == want, f"read {i}: {coords[i]}"
    g.close()


@pytest.mark.gpu
@pytest.mark.parametrize("meth", [False, True])
def test_rank_shards_through_the_library_equal_one_job(meth):
    """the N-rank decomposition run through libsqg itself: every rank's shard of a coordinate job -
    first_read_index = its first global read, meth_draw_base from the host-side CpG prefix index, no exchange between
    ranks - gives the bytes, signals and per-read draws of the one-job call; so does the host-bases path per shard"""
    import squigulator_b200 as sq
    from squigulator_b200 import api
    from squigulator_b200.shard import shard_range, CpgIndex
    k = 6
    contigs, marr = H.synthetic_genome(seed=41, n_contigs=4, mean_len=5000)
    coords = _coords(contigs, 101, seed=9, min_len=40)
    g = sq.SignalGenerator("dna-r9-prom", H.random_model((5 if meth else 4) ** k, seed=6), k, seed=17, meth=meth)
    g.load_genome(contigs, meth=marr if meth else None)
    first = 5000
    whole, draws = g.gen_batch_coords(coords, first_read_index=first, want=api.WANT_BASES)
    bases = CpgIndex(contigs).draw_bases(coords) if meth else np.zeros(len(coords) + 1, dtype=np.int64)
    assert int(bases[-1]) == draws
    for world in (2, 3):
        got = []
        for rank in range(world):
            lo, hi = shard_range(len(coords), rank, world)
            part, d = g.gen_batch_coords(coords[lo:hi], first_read_index=first + lo, meth_draw_base=int(bases[lo]), want=api.WANT_BASES)
            assert d == int(bases[hi] - bases[lo])
            # the same shard with its bases handed over by the host
            again = g.gen_batch([o["bases"] for o in part], first_read_index=first + lo)
            for a, b in zip(part, again):
                np.testing.assert_array_equal(a["sig"], b["sig"])
            got += part
        assert len(got) == len(whole)
        for a, b in zip(got, whole):
            assert a["bases"] == b["bases"] and a["offset"] == b["offset"] and a["median_before"] == b["median_before"]
            np.testing.assert_array_equal(a["sig"], b["sig"])
    g.close()


@pytest.mark.gpu
def test_coords_golden_reads_through_gpu():
    """the reference's own gen_read() output (golden) out of the GPU extraction, RNA included"""
    import squigulator_b200 as sq
    from squigulator_b200 import api
    doc = json.load(open(os.path.join(H.GOLDEN_DIR, "gen_read.json")))
    for case in doc:
        contigs, marr = H.synthetic_genome(seed=case["genome_seed"], with_meth=True)
        meth = case["meth"]
        k = 5
        g = sq.SignalGenerator(case["preset"], H.random_model((5 if meth else 4) ** k, seed=2), k, seed=case["seed"], meth=meth)
        g.load_genome(contigs, meth=marr if meth else None)
        coords = [(r["contig"], r["pos"], r["len"], r["strand"]) for r in case["reads"]]
        out, _ = g.gen_batch_coords(coords, want=api.WANT_BASES)
        for o, r in zip(out, case["reads"]):
            assert hashlib.sha256(o["bases"]).hexdigest() == r["sha256"]
        g.close()


@pytest.mark.gpu
def test_coords_errors():
    import squigulator_b200 as sq
    g = sq.SignalGenerator("dna-r9-prom", H.random_model(4 ** 6), 6)
    with pytest.raises(sq.SqgError) as e:
        g.gen_batch_coords([(0, 0, 10, "+")])
    assert e.value.code == -5  # SQG_ERR_STATE: no genome
    g.load_genome([b"ACGT" * 50, b"GGCC" * 10])
    for bad in [(2, 0, 10, "+"), (0, 195, 10, "+"), (1, -1, 4, "+"), (0, 0, -1, "-"), (0, 0, 10, "x")]:
        with pytest.raises(sq.SqgError) as e:
            g.gen_batch_coords([bad])
        assert e.value.code == -1, bad
    out, d = g.gen_batch_coords([])
    assert out == [] and d == 0
    out, d = g.gen_batch_coords([(1, 36, 4, "-")], want=0x8)
    assert out[0]["bases"] == b"GGCC" and d == 0
    g.close()
