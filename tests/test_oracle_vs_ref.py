"""The oracle's legacy mode against the UNMODIFIED reference (oracle/_ref/libsqref.so, built from the sources where
they lie): random reads, every preset, every flag combination on the hot path, several thread stream sets.
Skipped where the compiled reference is absent (it is present in the build container and on the GPU box)."""
import ctypes as C
import os

import numpy as np
import pytest

from tests import helpers as H

SO = os.path.join(H.ROOT, "oracle", "_ref", "libsqref.so")
pytestmark = pytest.mark.skipif(not os.path.exists(SO), reason="oracle/_ref/libsqref.so not built")


@pytest.fixture(scope="module")
def ref():
    lib = C.CDLL(SO)
    lib.sqref_profile.argtypes = [C.c_char_p, C.POINTER(H.Profile), C.POINTER(C.c_uint32)]
    lib.sqref_open.restype = C.c_void_p
    lib.sqref_open.argtypes = [C.POINTER(H.Profile), C.c_uint32, C.c_int64, C.c_int32, C.c_float, C.c_int, C.c_char_p,
                               C.c_char_p, C.c_int]
    lib.sqref_get_model.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
    lib.sqref_kmer_size.restype = C.c_uint32
    lib.sqref_kmer_size.argtypes = [C.c_void_p]
    lib.sqref_num_kmer.restype = C.c_uint32
    lib.sqref_num_kmer.argtypes = [C.c_void_p]
    lib.sqref_gen_sig.restype = C.c_int64
    lib.sqref_gen_sig.argtypes = [C.c_void_p, C.c_char_p, C.c_int32, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double),
                                  C.POINTER(C.POINTER(C.c_int16)), C.POINTER(C.POINTER(C.c_int32)), C.POINTER(C.c_int64)]
    lib.sqref_free_buf.argtypes = [C.c_void_p]
    lib.sqref_close.argtypes = [C.c_void_p]
    return lib


def ref_gen(lib, h, read, tid=0):
    off, mb = C.c_double(), C.c_double()
    sig = C.POINTER(C.c_int16)()
    ss = C.POINTER(C.c_int32)()
    ss_n = C.c_int64()
    n = lib.sqref_gen_sig(h, read, len(read), tid, C.byref(off), C.byref(mb), C.byref(sig), C.byref(ss), C.byref(ss_n))
    out = dict(offset=off.value, median_before=mb.value, sig=np.ctypeslib.as_array(sig, shape=(n,)).copy(),
               ss=np.ctypeslib.as_array(ss, shape=(ss_n.value,)).copy())
    lib.sqref_free_buf(sig)
    lib.sqref_free_buf(ss)
    return out


CASES = [
    ("dna-r9-prom", 0, 0, 1.0, b"ACGT"), ("dna-r9-min", 0, 0, 1.0, b"ACGTNacgtRYKMSWBDHV"),
    ("dna-r10-prom", 0, 0, 1.0, b"ACGT"), ("dna-r10-min", 0, 0, 0.5, b"ACGT"),
    ("rna-r9-prom", 0, 0, 1.0, b"ACGT"), ("rna-r9-min", H.SQ_PREFIX, 0, 1.0, b"ACGT"),
    ("rna004-prom", 0, 0, 1.0, b"ACGT"), ("rna004-min", H.SQ_PREFIX, 0, 1.0, b"ACGT"),
    ("dna-r9-prom", H.SQ_PREFIX, 0, 1.0, b"ACGT"), ("dna-r9-prom", H.SQ_IDEAL, 0, 1.0, b"ACGT"),
    ("dna-r10-prom", H.SQ_IDEAL_TIME, 0, 1.0, b"ACGT"), ("dna-r10-prom", H.SQ_IDEAL_AMP, 0, 1.0, b"ACGT"),
    ("dna-r9-prom", 0, 1, 1.0, b"ACGTM"), ("dna-r10-prom", 0, 1, 1.0, b"ACGTM"),
]


@pytest.mark.parametrize("case", CASES, ids=[f"{c[0]}-f{c[1]:x}-m{c[2]}-a{c[3]}" for c in CASES])
def test_oracle_equals_reference(case, ref, oracle_lib):
    preset, xflags, meth, amp, alpha = case
    p, f = H.Profile(), C.c_uint32()
    assert ref.sqref_profile(preset.encode(), C.byref(p), C.byref(f)) == 0
    flags = f.value | xflags
    seed, nthreads = 42, 2
    h = ref.sqref_open(C.byref(p), flags, seed, nthreads, amp, meth, None, None, 0)
    assert h
    k, n = ref.sqref_kmer_size(h), ref.sqref_num_kmer(h)
    model = np.zeros(2 * n, dtype=np.float32)
    ref.sqref_get_model(h, model.ctypes.data_as(C.POINTER(C.c_float)))
    prof = {fld: getattr(p, fld) for fld in H.PROFILE_FIELDS}
    o = H.Oracle(oracle_lib, prof, flags, k, n, model, seed, H.RNG_LEGACY, meth=meth, amp_noise=amp, num_thread=nthreads)
    reads = H.random_reads(6, 700, seed=len(preset) + xflags, alphabet=alpha) + [b"", b"ACG", b"T" * 300]
    for i, r in enumerate(reads):
        tid = i % nthreads  # streams of different threads are independent (reference src/sim.c:236-257)
        a, b = ref_gen(ref, h, r, tid), o.gen_sig(r, read_index=i, tid=tid, want_ss=True)
        np.testing.assert_array_equal(a["sig"], b["sig"], err_msg=f"read {i}")
        np.testing.assert_array_equal(a["ss"], b["ss"])
        assert a["offset"] == b["offset"] and a["median_before"] == b["median_before"]
    o.close()
    ref.sqref_close(h)
