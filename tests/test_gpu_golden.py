"""The reference's golden files through the GPU path (SQG_RNG_LEGACY): results identical to the reference's on the
same inputs.  tests/golden/*.npz carry the reads, options and sparse model of every scripts/test.sh case together
with the reference's own output (see scripts/make_golden.py)."""
import os

import numpy as np
import pytest

from tests import helpers as H

pytestmark = pytest.mark.gpu
PATHS = H.golden_paths()


@pytest.fixture(scope="module")
def sq():
    import squigulator_b200 as s
    s.load_library()
    return s


def check(g, got, lo=0):
    for j, r in enumerate(got):
        i = lo + j
        assert len(r["sig"]) == g.sig_len[i], (g.name, i)
        if i < g.n_full:
            bad = np.nonzero(r["sig"] != g.sig_full[i])[0]
            assert bad.size == 0, f"{g.name} read {i}: {bad.size} samples differ, first {bad[:5]}"
        assert H.sha256_i16(r["sig"]) == g.sha_of(i), (g.name, i)
        assert abs(r["offset"] - g.offset[i]) < 1e-6 and abs(r["median_before"] - g.median_before[i]) < 1e-6
        if g.ss is not None:
            np.testing.assert_array_equal(r["ss"], g.ss[i])


@pytest.mark.parametrize("path", PATHS, ids=[os.path.basename(p)[:-4] for p in PATHS])
def test_gpu_reproduces_reference_golden(path, sq):
    g = H.Golden(path)
    c = g.cfg
    gen = sq.SignalGenerator(c["profile"], g.dense_model(), c["kmer_size"], flags=c["flags"], seed=c["seed"],
                             meth=bool(c["meth"]), amp_noise=c["amp_noise"], rng_mode=sq.RNG_LEGACY)
    check(g, gen.gen_batch(g.reads, want_ss=g.ss is not None))
    gen.close()


@pytest.mark.parametrize("name", ["dna_basic", "rna_basic_prefix", "r9_meth"])
def test_stream_state_carries_across_batches(name, sq):
    """Reference semantics: the per-k-mer streams persist across reads (src/sim.c:215-258), so -K must not matter."""
    g = H.Golden(os.path.join(H.GOLDEN_DIR, name + ".npz"))
    c = g.cfg
    gen = sq.SignalGenerator(c["profile"], g.dense_model(), c["kmer_size"], flags=c["flags"], seed=c["seed"],
                             meth=bool(c["meth"]), amp_noise=c["amp_noise"], rng_mode=sq.RNG_LEGACY)
    n = len(g.reads)
    cuts = [0, 1, max(2, n // 2), n]
    for a, b in zip(cuts[:-1], cuts[1:]):
        if b > a:
            check(g, gen.gen_batch(g.reads[a:b], first_read_index=a), lo=a)
    gen.close()
