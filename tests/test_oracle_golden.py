"""Pin the oracle: in legacy-RNG mode it must reproduce the reference's own golden files bit for bit
(reference scripts/test.sh + test/*.exp, carried here as tests/golden/*.npz by scripts/make_golden.py)."""
import os

import numpy as np
import pytest

from tests import helpers as H

PATHS = H.golden_paths()


def test_fixtures_present():
    assert len(PATHS) >= 20


@pytest.mark.parametrize("path", PATHS, ids=[os.path.basename(p)[:-4] for p in PATHS])
def test_oracle_reproduces_reference_golden(path, oracle_lib):
    g = H.Golden(path)
    c = g.cfg
    o = H.Oracle(oracle_lib, c["profile"], c["flags"], c["kmer_size"], c["num_kmer"], g.dense_model(), c["seed"],
                 H.RNG_LEGACY, meth=c["meth"], amp_noise=c["amp_noise"])
    start = 0
    for i, read in enumerate(g.reads):
        r = o.gen_sig(read, read_index=i, want_ss=g.ss is not None)
        assert len(r["sig"]) == g.sig_len[i], (g.name, i)
        if i < g.n_full:
            np.testing.assert_array_equal(r["sig"], g.sig_full[i], err_msg=f"{g.name} read {i}")
        assert H.sha256_i16(r["sig"]) == g.sha_of(i), (g.name, i)
        # SLOW5 prints doubles with 6 decimals
        assert abs(r["offset"] - g.offset[i]) < 1e-6 and abs(r["median_before"] - g.median_before[i]) < 1e-6
        if g.ss is not None:
            np.testing.assert_array_equal(r["ss"], g.ss[i])
        assert g.start_time[i] == start  # aux start_time = samples emitted before this read (src/sim.c:602)
        start += len(r["sig"])
    o.close()
