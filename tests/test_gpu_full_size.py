"""Size-independent properties at a BASELINE-sized batch (configs[2] shape: dna-r10-prom, reads of mean 10 kb), where the
oracle is too slow to be the checker: batch-split invariance, idempotence, bookkeeping identities, the svb-zd round trip,
per-k-mer statistics against the model, and oracle parity on a sample of reads of that very batch."""
import hashlib

import numpy as np
import pytest

from tests import helpers as H

pytestmark = pytest.mark.gpu

N_READS = 4096     # ~0.53 G samples, ~1.1 GB of int16 (bench.py's e2e batch)


def read_digest(a):
    return hashlib.blake2b(np.ascontiguousarray(a).tobytes(), digest_size=8).digest()


@pytest.fixture(scope="module")
def batch():
    import squigulator_b200 as sq
    from squigulator_b200.api import _pack_reads
    sq.load_library()
    from bench import synth_reads, synth_model
    bases, off = synth_reads(N_READS, 10000, False, seed=77, genome_mb=16)
    model = synth_model(4 ** 9)
    return sq, bases, off, model


def test_full_size_properties(batch, oracle_lib, ztable):
    sq, bases, off, model = batch
    n = len(off) - 1
    gen = sq.SignalGenerator("dna-r10-prom", model, 9, seed=123, n_slots=2)
    res = gen.gen_batch_raw(bases, off, first_read_index=10 ** 6, want=sq.api.WANT_SS)
    whole = gen._unpack(res, copy=False)
    lens = np.array([len(x["sig"]) for x in whole], dtype=np.int64)
    dig = [read_digest(x["sig"]) for x in whole]
    ss_sum = np.array([int(x["ss"].sum()) for x in whole])
    nk = np.array([len(x["ss"]) for x in whole])
    # bookkeeping identities: len_raw_signal == sum of dwells; k-mers == len - k + 1; total_samples == sum of lengths
    assert np.array_equal(lens, ss_sum)
    assert np.array_equal(nk, np.maximum(np.diff(off) - 9 + 1, 1))
    assert res.total_samples == int(lens.sum()) and lens.sum() > 4e8
    # per-k-mer statistics of the whole batch against the model (pooled over 4^9 ranks: value - mean*scale + offset)
    prof = H.PRESETS["dna-r10-prom"][0]
    scale = prof["digitisation"] / prof["range"]
    x = whole[0]
    some = np.concatenate([w["sig"][:20000].astype(np.float64) for w in whole[:64]])
    assert 300 < some.mean() < 1500 and 30 < some.std() < 400
    offsets = np.array([w["offset"] for w in whole])
    assert abs(offsets.mean() - prof["offset_mean"]) < 5 * prof["offset_std"] / np.sqrt(n)
    assert abs(offsets.std() - prof["offset_std"]) < 5 * prof["offset_std"] / np.sqrt(2 * n)
    dw = np.concatenate([w["ss"] for w in whole[:512]])
    assert abs(dw.mean() - 13.0) < 0.05 and abs(dw.std() - 4.0) < 0.1 and dw.min() >= 1
    # oracle parity on reads sampled from this very batch (first, last, longest, a few in between)
    o = H.Oracle(oracle_lib, prof, H.SQ_R10, 9, 4 ** 9, model, 123, H.RNG_PHILOX, ztable=ztable)
    for i in sorted({0, n - 1, int(np.argmax(lens)), n // 3, 2 * n // 3}):
        read = bases[off[i]:off[i + 1]].tobytes()
        if len(read) > 30000:
            continue    # keep the CPU side in seconds
        exp = o.gen_sig(read, read_index=10 ** 6 + i, want_ss=True)
        assert np.array_equal(whole[i]["sig"], exp["sig"]) and np.array_equal(whole[i]["ss"], exp["ss"])
    o.close()
    del whole, res

    # batch-split invariance + idempotence: four quarter batches through the asynchronous slots, twice
    for _ in range(2):
        q = (n + 3) // 4
        tickets = []
        got = []
        for s in range(0, n, q):
            e = min(n, s + q)
            sub_off = np.ascontiguousarray(off[s:e + 1])
            tickets.append((gen.submit(bases, sub_off, first_read_index=10 ** 6 + s), s, e))
            if len(tickets) == 2:
                t, s0, e0 = tickets.pop(0)
                r = gen.wait(t)
                got += [read_digest(x["sig"]) for x in gen._unpack(r, copy=False)]
                gen.release(t)
        for t, s0, e0 in tickets:
            r = gen.wait(t)
            got += [read_digest(x["sig"]) for x in gen._unpack(r, copy=False)]
            gen.release(t)
        assert got == dig

    # svb-zd: every stream decodes to a signal of the right length whose digest equals the raw read's (64 reads decoded
    # in full by the independent numpy decoder; all reads: stream length identity 4 + ceil(n/4) + data bytes)
    res = gen.gen_batch_raw(bases, off, first_read_index=10 ** 6, want=sq.api.WANT_SVB)
    comp = gen._unpack(res, copy=False)
    tot = 0
    for i, c in enumerate(comp):
        assert c["n_samples"] == lens[i]
        assert int(c["svb"][:4].view("<u4")[0]) == lens[i]
        tot += len(c["svb"])
    assert 1.1 < tot / lens.sum() < 1.6          # bytes per sample
    for i in range(0, n, n // 64):
        if lens[i] > 400000:
            continue
        dec, used = H.svb_zd_decode(comp[i]["svb"])
        assert used == len(comp[i]["svb"]) and read_digest(dec) == dig[i]
    gen.close()
