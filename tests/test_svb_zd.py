"""svb-zd (zig-zag delta + StreamVByte), slow5lib's signal compression: the oracle's restatement pinned to the compiled
reference, and its round-trip property.  SURVEY.md 8(f)-1."""
import ctypes as C
import os

import numpy as np
import pytest

from tests import helpers as H

REF = os.path.join(H.ROOT, "oracle", "_ref", "libsqref.so")


def signals():
    rs = np.random.RandomState(3)
    yield np.zeros(0, np.int16)
    yield np.array([5], np.int16)
    yield np.array([-32768, 32767, -32768, 0, 1, -1, 127, 128, -128, -129], np.int16)     # 1-, 2- and 3-byte deltas
    for n in (3, 4, 5, 7, 8, 9, 1000, 4097):
        yield (600 + 25 * rs.standard_normal(n)).astype(np.int16)                           # signal-like
        yield rs.randint(-32768, 32768, n).astype(np.int16)                                 # full range
    yield np.repeat((500 + 80 * rs.standard_normal(300)).astype(np.int16), 13)              # --ideal-amp like


@pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref not built")
def test_oracle_equals_slow5lib(oracle_lib):
    ref = C.CDLL(REF)
    ref.sqref_svb_zd.restype = C.c_int64
    ref.sqref_svb_zd.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64]
    for sig in signals():
        if sig.size == 0:
            continue    # slow5lib is never handed an empty signal (src/sim.c:558 asserts len > 0)
        want = np.empty(8 * sig.size + 64, np.uint8)
        n = ref.sqref_svb_zd(sig.ctypes.data_as(C.c_void_p), sig.size, want.ctypes.data_as(C.c_void_p), want.size)
        assert n > 0
        got = H.oracle_svb_zd(oracle_lib, sig)
        assert got.size == n and np.array_equal(got, want[:n]), sig[:8]


def test_round_trip(oracle_lib):
    for sig in signals():
        enc = H.oracle_svb_zd(oracle_lib, sig)
        dec, used = H.svb_zd_decode(enc)
        assert used == enc.size
        assert np.array_equal(dec, sig)


def test_ss_text_equals_reference_goldens(oracle_lib):
    """the dwell string of PAF/SAM records.  The oracle's restatement of src/format.c:69-75, fed with
    the dwell arrays of the reference's PAF goldens, reproduces their `ss:Z:` values verbatim (RNA: last k-mer first)."""
    import json
    texts = json.load(open(os.path.join(H.GOLDEN_DIR, "ss_text.json")))
    assert set(texts) == {"dna_r10_paf", "rna_r9_paf"}
    for name, lines in texts.items():
        g = H.Golden(os.path.join(H.GOLDEN_DIR, name + ".npz"))
        rna = bool(g.cfg["flags"] & H.SQ_RNA)
        assert len(lines) == len(g.ss)
        for ss, want in zip(g.ss, lines):
            assert H.oracle_ss_text(oracle_lib, ss, rna) == want.encode()
    assert H.oracle_ss_text(oracle_lib, np.array([1, 10, 100, 1000, 7], np.int32), False) == b"1,10,100,1000,7,"
    assert H.oracle_ss_text(oracle_lib, np.array([1, 10, 100], np.int32), True) == b"100,10,1,"
    assert H.oracle_ss_text(oracle_lib, np.zeros(0, np.int32), False) == b""
