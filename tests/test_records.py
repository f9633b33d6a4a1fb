"""SQG_WANT_RECORDS: the GPU's BLOW5 records against the UNMODIFIED slow5lib + the reference's own field setters
(compiled into oracle/_ref/libsqref.so): byte for byte per record, and a file written from the GPU's bytes read back by
slow5lib with identical signals.  slow5lib/src/slow5.c:3815-4010, src/gensig.c:130-217."""
import ctypes as C
import os

import numpy as np
import pytest

from tests import helpers as H

pytestmark = pytest.mark.gpu
REF_SO = os.path.join(H.ROOT, "oracle", "_ref", "libsqref.so")
needs_ref = pytest.mark.skipif(not os.path.exists(REF_SO), reason="oracle/_ref not built")


def ref_lib():
    lib = C.CDLL(REF_SO)
    lib.sqref_rec_open.restype = C.c_void_p
    lib.sqref_rec_open.argtypes = [C.c_char_p, C.POINTER(H.Profile), C.c_uint32, C.c_int]
    lib.sqref_rec_encode.restype = C.c_int64
    lib.sqref_rec_encode.argtypes = [C.c_void_p, C.POINTER(H.Profile), C.c_char_p, C.c_double, C.c_void_p, C.c_int64, C.c_double,
                                     C.c_int32, C.c_uint64, C.c_int, C.c_void_p, C.c_int64]
    lib.sqref_rec_write_bytes.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
    lib.sqref_rec_close.argtypes = [C.c_void_p]
    return lib


@needs_ref
@pytest.mark.parametrize("preset,k,ont", [("dna-r10-prom", 9, False), ("rna004-prom", 9, True), ("dna-r9-min", 6, False)])
def test_records_equal_slow5lib_and_read_back(preset, k, ont, tmp_path):
    import squigulator_b200 as sq
    from integration.blow5_stat import Rec
    prof, flags = H.PRESETS[preset]
    reads = H.random_reads(40, 3000, seed=21) + [b"ACGTAC", b"ACGT" * 700, b"T" * 2047, b"G" * 2049]
    ids = [f"S1_{i + 1}!contig{'x' * (i % 7)}!{i * 17}!{i * 17 + len(r)}!{'+-'[i & 1]}" for i, r in enumerate(reads)]
    first, st0 = 1000, 123456789012
    gen = sq.SignalGenerator(dict(prof), H.random_model(4 ** k), k, flags=flags, seed=3)
    plain = gen.gen_batch(reads, first_read_index=first)
    blob, recs = gen.gen_batch_records(reads, ids, first_read_index=first, start_time0=st0, ont_friendly=ont)
    gen.close()
    lib = ref_lib()
    p = H.make_profile(prof)
    path = str(tmp_path / "gpu.blow5")
    sp = lib.sqref_rec_open(path.encode(), C.byref(p), flags, int(ont))
    assert sp
    # 1. record by record: the bytes slow5_encode makes of the same read
    start = st0
    buf = np.empty(64 + 4 * max(len(r["sig"]) for r in plain) + 300, dtype=np.uint8)
    pos = 0
    for i, (r, rec) in enumerate(zip(plain, recs)):
        sig = np.ascontiguousarray(r["sig"], dtype=np.int16)
        n = lib.sqref_rec_encode(sp, C.byref(p), ids[i].encode(), r["offset"], sig.ctypes.data_as(C.c_void_p), sig.size,
                                 r["median_before"], first + i, start, int(ont), buf.ctypes.data_as(C.c_void_p), buf.size)
        assert 0 < n <= buf.size
        want = buf[:n].tobytes()
        got = rec["svb"].tobytes()
        assert got == want, f"read {i}: record differs ({len(got)} vs {len(want)} bytes; first difference at " \
                            f"{next((j for j in range(min(len(got), len(want))) if got[j] != want[j]), -1)})"
        assert blob[pos:pos + n] == want        # records lie back to back
        pos += n
        start += sig.size
    assert pos == len(blob)
    # 2. a BLOW5 file = slow5lib's header + the GPU's bytes, written with ONE call, read back by slow5lib
    assert lib.sqref_rec_write_bytes(sp, blob, len(blob)) >= 0
    lib.sqref_rec_close(sp)
    lib.slow5_open.restype = C.c_void_p
    lib.slow5_open.argtypes = [C.c_char_p, C.c_char_p]
    lib.slow5_get_next.argtypes = [C.POINTER(C.POINTER(Rec)), C.c_void_p]
    lib.slow5_rec_free.argtypes = [C.POINTER(Rec)]
    lib.slow5_close.argtypes = [C.c_void_p]
    rp = lib.slow5_open(path.encode(), b"r")
    assert rp
    rec = C.POINTER(Rec)()
    i = 0
    while lib.slow5_get_next(C.byref(rec), rp) >= 0:
        c = rec.contents
        assert c.read_id == ids[i].encode() and c.len_raw_signal == len(plain[i]["sig"]) and c.offset == plain[i]["offset"]
        assert np.array_equal(np.ctypeslib.as_array(c.raw_signal, shape=(c.len_raw_signal,)), plain[i]["sig"]), i
        i += 1
    lib.slow5_rec_free(rec)
    lib.slow5_close(rp)
    assert i == len(reads)
