#!/usr/bin/env python3
"""Reads a BLOW5 file back through the UNMODIFIED slow5lib compiled into oracle/_ref/libsqref.so (slow5_open /
slow5_get_next) and prints records, total samples and a checksum of the signals - test infrastructure."""
import ctypes as C
import os
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


class Rec(C.Structure):   # slow5_rec_t, slow5lib/include/slow5/slow5.h (leading fields)
    _fields_ = [("read_id_len", C.c_uint16), ("read_id", C.c_char_p), ("read_group", C.c_uint32), ("digitisation", C.c_double),
                ("offset", C.c_double), ("range", C.c_double), ("sampling_rate", C.c_double), ("len_raw_signal", C.c_uint64),
                ("raw_signal", C.POINTER(C.c_int16))]


def stat(path):
    lib = C.CDLL(os.path.join(os.path.dirname(HERE), "oracle", "_ref", "libsqref.so"))
    lib.slow5_open.restype = C.c_void_p
    lib.slow5_open.argtypes = [C.c_char_p, C.c_char_p]
    lib.slow5_get_next.argtypes = [C.POINTER(C.POINTER(Rec)), C.c_void_p]
    lib.slow5_rec_free.argtypes = [C.POINTER(Rec)]
    lib.slow5_close.argtypes = [C.c_void_p]
    sp = lib.slow5_open(path.encode(), b"r")
    assert sp, path
    rec = C.POINTER(Rec)()
    n = tot = 0
    crc = 0
    while lib.slow5_get_next(C.byref(rec), sp) >= 0:
        r = rec.contents
        n += 1
        tot += r.len_raw_signal
        crc = zlib.crc32(np.ctypeslib.as_array(r.raw_signal, shape=(r.len_raw_signal,)).tobytes(), crc)
    lib.slow5_rec_free(rec)
    lib.slow5_close(sp)
    return n, tot, crc


if __name__ == "__main__":
    for p in sys.argv[1:]:
        n, tot, crc = stat(p)
        print(f"{os.path.basename(p)}: {n} records, {tot} samples, crc32 of all signals {crc:08x}, {os.path.getsize(p)} bytes")
