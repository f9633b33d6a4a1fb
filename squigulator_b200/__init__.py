"""squigulator_b200 — B200 (sm_100a) signal-generation path of squigulator behind a C ABI.

The product is ``libsqg.so`` (``include/sqg.h``, sources in ``squigulator_b200/csrc``).  This package is
only the thin ctypes binding used by the tests and by ``bench.py``; it mirrors the reference's call
shapes (``gen_sig`` per read, one call per batch) and contains no compute of its own.  There is no CPU
fallback: loading fails loudly when the CUDA library is missing.
"""
from .api import (SQ_RNA, SQ_IDEAL, SQ_IDEAL_TIME, SQ_IDEAL_AMP, SQ_PREFIX, SQ_R10, RNG_PHILOX, RNG_LEGACY,
                  PROFILES, Profile, SignalGenerator, SqgError, lib_path, load_library)

__all__ = ["SQ_RNA", "SQ_IDEAL", "SQ_IDEAL_TIME", "SQ_IDEAL_AMP", "SQ_PREFIX", "SQ_R10", "RNG_PHILOX", "RNG_LEGACY",
           "PROFILES", "Profile", "SignalGenerator", "SqgError", "lib_path", "load_library"]
