#!/usr/bin/env python3
"""Golden vectors for the read-extraction step: the UNMODIFIED reference's gen_read()
(src/genread.c:357, reached through oracle/_ref/libsqref.so) over a small synthetic genome.

    make -C oracle ref && python scripts/make_golden_reads.py      -> tests/golden/gen_read.json

The genome is regenerated from its seed by tests/helpers.synthetic_genome(); the file keeps the coordinates
the reference drew and a SHA-256 of every read it returned, so the oracle's restatement is pinned even where
oracle/_ref is not built."""
import ctypes as C
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import helpers as H  # noqa: E402


def load_ref():
    lib = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libsqref.so"))
    lib.sqref_profile.argtypes = [C.c_char_p, C.POINTER(H.Profile), C.POINTER(C.c_uint32)]
    lib.sqref_open.restype = C.c_void_p
    lib.sqref_open.argtypes = [C.POINTER(H.Profile), C.c_uint32, C.c_int64, C.c_int32, C.c_float, C.c_int, C.c_char_p,
                               C.c_char_p, C.c_int]
    lib.sqref_set_genome.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_void_p, C.c_void_p]
    lib.sqref_gen_read.restype = C.c_int32
    lib.sqref_gen_read.argtypes = [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_char_p, C.c_char_p, C.c_int32]
    return lib


def reference_reads(lib, preset, seed, meth, contigs, meth_arrays, n_reads):
    """n_reads accepted reads from thread 0 of the reference: list of (contig, pos, len, strand, bytes)"""
    p, f = H.Profile(), C.c_uint32()
    assert lib.sqref_profile(preset.encode(), C.byref(p), C.byref(f)) == 0
    h = lib.sqref_open(C.byref(p), f.value, seed, 1, 1.0, 1 if meth else 0, None, None, 0)
    assert h
    seq = b"".join(contigs)
    off = np.zeros(len(contigs) + 1, dtype=np.int64)
    off[1:] = np.cumsum([len(c) for c in contigs])
    m = np.concatenate(meth_arrays).astype(np.uint8) if meth else None
    lib.sqref_set_genome(h, len(contigs), seq, off.ctypes.data_as(C.c_void_p), m.ctypes.data_as(C.c_void_p) if meth else None)
    cap = max(len(c) for c in contigs) + 1
    buf = C.create_string_buffer(cap)
    out = []
    for _ in range(n_reads):
        ctg, pos, st = C.c_int32(), C.c_int32(), C.c_char()
        n = lib.sqref_gen_read(h, C.byref(ctg), C.byref(pos), C.cast(C.byref(st), C.c_char_p), buf, cap)
        assert n >= 0 and ctg.value >= 0
        out.append((ctg.value, pos.value, n, st.value.decode(), buf.raw[:n]))
    return out  # the handle is left open on purpose: sqref_close() does not know about the shim's genome


CASES = [  # (name, preset, seed, meth, genome seed, reads)
    ("dna-r9-meth", "dna-r9-prom", 42, True, 11, 60),
    ("dna-r10-plain", "dna-r10-prom", 7, False, 12, 40),
    ("rna004", "rna004-prom", 3, False, 13, 25),
]

if __name__ == "__main__":
    lib = load_ref()
    doc = []
    for name, preset, seed, meth, gseed, n in CASES:
        contigs, marr = H.synthetic_genome(seed=gseed, with_meth=True)
        reads = reference_reads(lib, preset, seed, meth, contigs, marr, n)
        doc.append(dict(name=name, preset=preset, seed=seed, meth=meth, genome_seed=gseed,
                        genome_sha256=hashlib.sha256(b"".join(contigs)).hexdigest(),
                        reads=[dict(contig=c, pos=p, len=l, strand=s, sha256=hashlib.sha256(b).hexdigest(),
                                    n_M=b.count(b"M"), head=b[:24].decode()) for c, p, l, s, b in reads]))
    path = os.path.join(ROOT, "tests", "golden", "gen_read.json")
    with open(path, "w") as fh:
        json.dump(doc, fh, indent=0)
    print(path, sum(len(d["reads"]) for d in doc), "reads")
