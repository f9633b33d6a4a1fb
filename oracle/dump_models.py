#!/usr/bin/env python3
"""Dump the reference's built-in pore-model tables (src/model.h, src/methmodel.c) as raw float32 files under
oracle/_ref/models/ - run by __graft_entry__.build() in the container that has /root/reference, through the compiled,
unmodified reference (oracle/_ref/libsqref.so).  The files are git-ignored (they derive from the reference's sources)
but travel to the GPU box, where bench.py and the GPU tests read them as plain data: interleaved
(level_mean, level_stdv), rank order, exactly what sqg_init takes.

    python oracle/dump_models.py            # -> oracle/_ref/models/<name>.f32
"""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from tests import helpers as H  # noqa: E402

# name -> (preset, meth)
TABLES = {"dna-r9-prom": ("dna-r9-prom", 0), "dna-r10-prom": ("dna-r10-prom", 0), "rna004-prom": ("rna004-prom", 0),
          "rna-r9-prom": ("rna-r9-prom", 0), "dna-r9-prom-meth": ("dna-r9-prom", 1), "dna-r10-prom-meth": ("dna-r10-prom", 1)}


def main():
    so = os.path.join(HERE, "_ref", "libsqref.so")
    lib = C.CDLL(so)
    lib.sqref_open.restype = C.c_void_p
    lib.sqref_open.argtypes = [C.POINTER(H.Profile), C.c_uint32, C.c_int64, C.c_int32, C.c_float, C.c_int, C.c_char_p, C.c_char_p, C.c_int]
    lib.sqref_num_kmer.restype = C.c_uint32
    lib.sqref_num_kmer.argtypes = [C.c_void_p]
    lib.sqref_kmer_size.restype = C.c_uint32
    lib.sqref_kmer_size.argtypes = [C.c_void_p]
    lib.sqref_get_model.argtypes = [C.c_void_p, C.c_void_p]
    lib.sqref_close.argtypes = [C.c_void_p]
    out = os.path.join(HERE, "_ref", "models")
    os.makedirs(out, exist_ok=True)
    for name, (preset, meth) in TABLES.items():
        prof, flags = H.PRESETS[preset]
        p = H.make_profile(prof)
        h = lib.sqref_open(C.byref(p), flags, 1, 1, 1.0, meth, None, None, 0)
        assert h, name
        n, k = lib.sqref_num_kmer(h), lib.sqref_kmer_size(h)
        m = np.empty(2 * n, dtype=np.float32)
        lib.sqref_get_model(h, m.ctypes.data_as(C.c_void_p))
        lib.sqref_close(h)
        assert n == (5 if meth else 4) ** k
        m.tofile(os.path.join(out, f"{name}.f32"))
        print(f"{name}: k={k} num_kmer={n} mean(level_mean)={m[0::2].mean():.3f} mean(level_stdv)={m[1::2].mean():.3f}")


if __name__ == "__main__":
    main()
