"""ctypes binding of include/sqg.h.  Names follow the reference (profile_t, model_t, gen_sig, SQ_* flags)."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

# reference src/sq.h:33-43
SQ_RNA, SQ_FULL_CONTIG, SQ_IDEAL, SQ_IDEAL_TIME, SQ_IDEAL_AMP, SQ_PREFIX, SQ_R10 = 0x001, 0x002, 0x004, 0x008, 0x010, 0x020, 0x040
RNG_PHILOX, RNG_LEGACY = 0, 1
WANT_SS = 0x1
WANT_SVB = 0x2
WANT_BASES = 0x8
WANT_SS_TEXT = 0x4
WANT_RECORDS = 0x10

PROFILE_FIELDS = ("digitisation", "sample_rate", "bps", "range", "offset_mean", "offset_std",
                  "median_before_mean", "median_before_std", "dwell_mean", "dwell_std")


class Profile(C.Structure):
    """== profile_t (reference src/sq.h:47-58)"""
    _fields_ = [(f, C.c_double) for f in PROFILE_FIELDS]

    @classmethod
    def from_dict(cls, d):
        p = cls()
        for f in PROFILE_FIELDS:
            setattr(p, f, float(d[f]))
        return p

    def to_dict(self):
        return {f: getattr(self, f) for f in PROFILE_FIELDS}


# -x presets: the struct literals of reference src/sim.c:55-150 and the flags set_profile() adds (src/sim.c:152-188)
PROFILES = {
    "dna-r9-min": (dict(digitisation=8192, sample_rate=4000, bps=450, range=1443.030273, offset_mean=13.7222605, offset_std=10.25279688, median_before_mean=200.815801, median_before_std=20.48933762, dwell_mean=9.0, dwell_std=4.0), 0),
    "dna-r9-prom": (dict(digitisation=2048, sample_rate=4000, bps=450, range=748.5801, offset_mean=-237.4102, offset_std=14.1575, median_before_mean=214.2890337, median_before_std=18.0127916, dwell_mean=9.0, dwell_std=4.0), 0),
    "rna-r9-min": (dict(digitisation=8192, sample_rate=3012, bps=70, range=1126.47, offset_mean=4.65491888, offset_std=4.115262472, median_before_mean=242.6584118, median_before_std=10.60230888, dwell_mean=43.0, dwell_std=35.0), SQ_RNA),
    "rna-r9-prom": (dict(digitisation=2048, sample_rate=3000, bps=70, range=548.788269, offset_mean=-231.9440589, offset_std=12.87185278, median_before_mean=238.5286796, median_before_std=21.1871794, dwell_mean=43.0, dwell_std=35.0), SQ_RNA),
    "dna-r10-prom": (dict(digitisation=2048, sample_rate=5000, bps=400, range=281.345551, offset_mean=-127.5655735, offset_std=19.377283387665, median_before_mean=189.87607393756, median_before_std=15.788097978713, dwell_mean=13.0, dwell_std=4.0), SQ_R10),
    "dna-r10-min": (dict(digitisation=8192, sample_rate=5000, bps=400, range=1536.598389, offset_mean=13.380569389019, offset_std=16.311471649012, median_before_mean=202.15407438804, median_before_std=13.406139241768, dwell_mean=13.0, dwell_std=4.0), SQ_R10),
    "rna004-prom": (dict(digitisation=2048, sample_rate=4000, bps=130, range=299.432068, offset_mean=-259.421128, offset_std=16.010841823643, median_before_mean=205.63935594369, median_before_std=8.3994882799157, dwell_mean=31.0, dwell_std=0.0), SQ_R10 | SQ_RNA),
    "rna004-min": (dict(digitisation=8192, sample_rate=4000, bps=130, range=1437.976685, offset_mean=12.47686423863, offset_std=10.442126577137, median_before_mean=205.08496731088, median_before_std=8.6671292866233, dwell_mean=31.0, dwell_std=0.0), SQ_R10 | SQ_RNA),
}


class Config(C.Structure):
    """== sqg_config_t"""
    _fields_ = [("profile", Profile), ("flags", C.c_uint32), ("kmer_size", C.c_uint32), ("num_kmer", C.c_uint32),
                ("meth", C.c_int32), ("amp_noise", C.c_float), ("seed", C.c_int64), ("rng_mode", C.c_int32),
                ("device", C.c_int32), ("n_slots", C.c_int32), ("reserved", C.c_int32)]


class Result(C.Structure):
    """== sqg_result_t"""
    _fields_ = [("n_reads", C.c_int64), ("total_samples", C.c_int64), ("signal", C.POINTER(C.c_int16)),
                ("sig_off", C.POINTER(C.c_int64)), ("len_raw_signal", C.POINTER(C.c_int64)),
                ("offset", C.POINTER(C.c_double)), ("median_before", C.POINTER(C.c_double)),
                ("ss", C.POINTER(C.c_int32)), ("ss_off", C.POINTER(C.c_int64)),
                ("svb", C.POINTER(C.c_uint8)), ("svb_off", C.POINTER(C.c_int64)), ("svb_len", C.POINTER(C.c_int64)),
                ("ss_text", C.POINTER(C.c_char)), ("ss_text_off", C.POINTER(C.c_int64)),
                ("bases", C.POINTER(C.c_char)), ("bases_off", C.POINTER(C.c_int64)), ("meth_draws", C.c_int64)]


class RecordInfo(C.Structure):
    """== sqg_record_info_t"""
    _fields_ = [("read_ids", C.c_void_p), ("id_off", C.c_void_p), ("start_time0", C.c_uint64), ("ont_friendly", C.c_int32),
                ("reserved", C.c_int32)]


# == sqg_coord_t, as a numpy record (one row per read)
COORD_DTYPE = np.dtype([("contig", np.int32), ("len", np.int32), ("pos", np.int64), ("strand", np.int32),
                        ("reserved", np.int32)])


class SqgError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libsqg error {code}: {msg}")
        self.code = code


# every symbol include/sqg.h declares, with its ctypes signature
_SIGNATURES = {
    "sqg_init": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(Config), C.c_void_p]),
    "sqg_init_device_model": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(Config), C.c_void_p]),
    "sqg_destroy": (None, [C.c_void_p]),
    "sqg_last_error": (C.c_char_p, [C.c_void_p]),
    "sqg_version": (C.c_char_p, []),
    "sqg_host_alloc": (C.c_void_p, [C.c_size_t]),
    "sqg_host_free": (None, [C.c_void_p]),
    "sqg_gen_batch": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_uint32, C.POINTER(Result)]),
    "sqg_gen_batch_records": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(RecordInfo), C.c_uint32, C.POINTER(Result)]),
    "sqg_submit_records": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(RecordInfo), C.c_uint32, C.POINTER(C.c_int64)]),
    "sqg_submit": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_uint32, C.POINTER(C.c_int64)]),
    "sqg_genome_load": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "sqg_gen_batch_coords": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int64, C.c_uint32, C.POINTER(Result)]),
    "sqg_submit_coords": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int64, C.c_uint32, C.POINTER(C.c_int64)]),
    "sqg_wait": (C.c_int, [C.c_void_p, C.c_int64, C.POINTER(Result)]),
    "sqg_release": (C.c_int, [C.c_void_p, C.c_int64]),
    "sqg_gen_sig": (C.POINTER(C.c_int16), [C.c_void_p, C.c_char_p, C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_double),
                                           C.POINTER(C.c_int64), C.c_int64, C.POINTER(C.POINTER(C.c_int32)), C.POINTER(C.c_int64)]),
    "sqg_dev_batch_create": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_uint32, C.POINTER(C.c_void_p)]),
    "sqg_dev_batch_run": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "sqg_dev_batch_info": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "sqg_dev_batch_fetch": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(Result)]),
    "sqg_dev_batch_destroy": (None, [C.c_void_p, C.c_void_p]),
    "sqg_bench_store": (C.c_int, [C.c_void_p, C.c_size_t, C.c_int32, C.POINTER(C.c_float)]),
    "sqg_launch_count": (C.c_int64, [C.c_void_p]),
}

_lib = None


def lib_path():
    # SQG_LIB: a differently tuned build of the same library (kernel experiments, scripts/perf_matrix.py)
    return os.environ.get("SQG_LIB") or os.path.join(_HERE, "libsqg.so")


def load_library():
    """dlopen libsqg.so (built in-tree by __graft_entry__.build()).  No fallback of any kind."""
    global _lib
    if _lib is None:
        path = lib_path()
        if not os.path.exists(path):
            raise ImportError(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(squigulator_b200 has no CPU fallback)")
        lib = C.CDLL(path)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if include/sqg.h and the library ever diverge
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def _pack_reads(reads):
    lens = np.fromiter((len(r) for r in reads), dtype=np.int64, count=len(reads))
    off = np.zeros(len(reads) + 1, dtype=np.int64)
    np.cumsum(lens, out=off[1:])
    bases = np.frombuffer(b"".join(reads), dtype=np.uint8) if len(reads) else np.zeros(0, np.uint8)
    return np.ascontiguousarray(bases), off


class SignalGenerator:
    """One sqg_ctx_t: what init_core()/init_rand() set up for the hot path (reference src/sim.c:215-326)."""

    def __init__(self, profile, model, kmer_size, flags=0, seed=1, meth=False, amp_noise=1.0, rng_mode=RNG_PHILOX,
                 device=0, n_slots=0, device_model_ptr=None, tuning=0):
        self.lib = load_library()
        if isinstance(profile, str):
            d, f = PROFILES[profile]
            profile, flags = Profile.from_dict(d), flags | f
        elif isinstance(profile, dict):
            profile = Profile.from_dict(profile)
        num_kmer = (5 if meth else 4) ** kmer_size
        self.cfg = Config(profile, flags & 0xFFFFFFFF, kmer_size, num_kmer, 1 if meth else 0, amp_noise, seed, rng_mode,
                          device, n_slots, tuning)
        self.h = C.c_void_p()
        if device_model_ptr is not None:
            rc = self.lib.sqg_init_device_model(C.byref(self.h), C.byref(self.cfg), C.c_void_p(device_model_ptr))
        else:
            m = np.ascontiguousarray(model, dtype=np.float32).reshape(-1)
            if m.size != 2 * num_kmer:
                raise ValueError(f"model must hold {num_kmer} (level_mean, level_stdv) pairs")
            rc = self.lib.sqg_init(C.byref(self.h), C.byref(self.cfg), m.ctypes.data_as(C.c_void_p))
        if rc != 0:
            raise SqgError(rc, self.lib.sqg_last_error(None).decode())

    # -- helpers
    def _check(self, rc):
        if rc != 0:
            raise SqgError(rc, self.lib.sqg_last_error(self.h).decode())

    @staticmethod
    def _unpack(res, copy=True):
        n = res.n_reads
        if n == 0:
            return []
        off = np.ctypeslib.as_array(res.sig_off, shape=(n,))
        ln = np.ctypeslib.as_array(res.len_raw_signal, shape=(n,))
        o = np.ctypeslib.as_array(res.offset, shape=(n,))
        mb = np.ctypeslib.as_array(res.median_before, shape=(n,))
        span = int((off + ln).max())
        sig = np.ctypeslib.as_array(res.signal, shape=(max(span, 1),)) if res.signal else None
        so = np.ctypeslib.as_array(res.ss_off, shape=(n + 1,))
        ss = np.ctypeslib.as_array(res.ss, shape=(max(int(so[-1]), 1),)) if res.ss else None
        if res.svb:  # SQG_WANT_SVB: every read's signal as slow5lib's svb-zd stream instead of raw int16
            vo = np.ctypeslib.as_array(res.svb_off, shape=(n + 1,))
            vl = np.ctypeslib.as_array(res.svb_len, shape=(n,))
            svb = np.ctypeslib.as_array(res.svb, shape=(max(int(vo[-1]), 1),))
        if res.ss_text:  # SQG_WANT_SS_TEXT: the PAF/SAM `ss:Z:` value of every read
            to = np.ctypeslib.as_array(res.ss_text_off, shape=(n + 1,))
            txt = C.string_at(res.ss_text, int(to[-1]))
        if res.bases:  # SQG_WANT_BASES (coordinate batches): the reads the GPU cut out of the genome
            bo = np.ctypeslib.as_array(res.bases_off, shape=(n + 1,))
            bases = C.string_at(res.bases, int(bo[-1]))
        out = []
        for i in range(n):
            d = dict(offset=float(o[i]), median_before=float(mb[i]), n_samples=int(ln[i]))
            if res.ss_text:
                d["ss_text"] = txt[to[i]:to[i + 1]]
            if sig is not None:
                s = sig[off[i]:off[i] + ln[i]]
                d["sig"] = s.copy() if copy else s
            if res.svb:
                d["svb"] = svb[vo[i]:vo[i] + vl[i]].copy()
            if ss is not None:
                d["ss"] = ss[so[i]:so[i + 1]].copy()
            if res.bases:
                d["bases"] = bases[bo[i]:bo[i + 1]]
            out.append(d)
        return out

    # -- the batch call (process_db's fan-out, reference src/sim.c:622)
    def gen_batch(self, reads, first_read_index=0, want_ss=False, want_svb=False, want_ss_text=False):
        bases, off = _pack_reads(reads)
        res = Result()
        self._check(self.lib.sqg_gen_batch(self.h, len(reads), bases.ctypes.data_as(C.c_void_p), off.ctypes.data_as(C.c_void_p),
                                           first_read_index, (WANT_SS if want_ss else 0) | (WANT_SVB if want_svb else 0) | (WANT_SS_TEXT if want_ss_text else 0),
                                           C.byref(res)))
        return self._unpack(res)

    def gen_batch_records(self, reads, read_ids, first_read_index=0, start_time0=0, ont_friendly=False):
        """finished BLOW5 records (record compression NONE, signal svb-zd): returns (bytes of all records back to back,
        per-read dicts with 'svb' = the record of the read)"""
        bases, off = _pack_reads(reads)
        ids, ioff = _pack_reads([i if isinstance(i, bytes) else i.encode() for i in read_ids])
        info = RecordInfo(ids.ctypes.data, ioff.ctypes.data, start_time0, 1 if ont_friendly else 0, 0)
        res = Result()
        self._check(self.lib.sqg_gen_batch_records(self.h, len(reads), bases.ctypes.data_as(C.c_void_p), off.ctypes.data_as(C.c_void_p),
                                                   first_read_index, C.byref(info), 0, C.byref(res)))
        total = int(res.svb_off[res.n_reads]) if res.n_reads else 0
        blob = C.string_at(res.svb, total) if total else b""
        return blob, self._unpack(res)

    def gen_batch_raw(self, bases, off, first_read_index=0, want=0):
        """host numpy buffers in, sqg_result_t (views into pinned memory) out — what bench.py's e2e leg times"""
        res = Result()
        self._check(self.lib.sqg_gen_batch(self.h, len(off) - 1, bases.ctypes.data_as(C.c_void_p), off.ctypes.data_as(C.c_void_p),
                                           first_read_index, want, C.byref(res)))
        return res

    # -- device-resident genome, reads by coordinates (what gen_read does for accepted reads, src/genread.c:357)
    def load_genome(self, contigs, meth=None, contig_has_meth=None):
        """contigs: list of bytes; meth: None or list of uint8 arrays (one per contig, same lengths)"""
        seq, off = _pack_reads(contigs)
        m = None
        if meth is not None:
            m = np.ascontiguousarray(np.concatenate([np.asarray(x, dtype=np.uint8) for x in meth]))
            assert m.size == seq.size
        f = None if contig_has_meth is None else np.ascontiguousarray(contig_has_meth, dtype=np.uint8)
        self._check(self.lib.sqg_genome_load(self.h, len(contigs), seq.ctypes.data_as(C.c_void_p), off.ctypes.data_as(C.c_void_p),
                                             None if m is None else m.ctypes.data_as(C.c_void_p),
                                             None if f is None else f.ctypes.data_as(C.c_void_p)))

    @staticmethod
    def pack_coords(coords):
        """coords: iterable of (contig, pos, len, strand) with strand '+' or '-'"""
        a = np.zeros(len(coords), dtype=COORD_DTYPE)
        for i, (c, pos, ln, st) in enumerate(coords):
            a[i] = (c, ln, pos, ord(st), 0)
        return a

    def gen_batch_coords(self, coords, first_read_index=0, meth_draw_base=0, want=0):
        """returns (per-read dicts, rand_meth draws the batch consumed)"""
        a = coords if isinstance(coords, np.ndarray) else self.pack_coords(coords)
        res = Result()
        self._check(self.lib.sqg_gen_batch_coords(self.h, len(a), a.ctypes.data_as(C.c_void_p) if len(a) else None,
                                                  first_read_index, meth_draw_base, want, C.byref(res)))
        return self._unpack(res), int(res.meth_draws)

    def submit_coords(self, coords, first_read_index=0, meth_draw_base=0, want=0):
        t = C.c_int64()
        self._check(self.lib.sqg_submit_coords(self.h, len(coords), coords.ctypes.data_as(C.c_void_p), first_read_index,
                                               meth_draw_base, want, C.byref(t)))
        return t.value

    # -- async dispatcher
    def submit(self, bases, off, first_read_index=0, want=0):
        t = C.c_int64()
        self._check(self.lib.sqg_submit(self.h, len(off) - 1, bases.ctypes.data_as(C.c_void_p), off.ctypes.data_as(C.c_void_p),
                                        first_read_index, want, C.byref(t)))
        return t.value

    def wait(self, ticket):
        res = Result()
        self._check(self.lib.sqg_wait(self.h, ticket, C.byref(res)))
        return res

    def release(self, ticket):
        self._check(self.lib.sqg_release(self.h, ticket))

    # -- the per-read call with gen_sig's shape (reference src/gensig.c:346)
    def gen_sig(self, read, read_index=0, want_ss=False):
        off, mb, n = C.c_double(), C.c_double(), C.c_int64()
        ss, ss_n = C.POINTER(C.c_int32)(), C.c_int64()
        p = self.lib.sqg_gen_sig(self.h, read, len(read), C.byref(off), C.byref(mb), C.byref(n), read_index,
                                 C.byref(ss) if want_ss else None, C.byref(ss_n))
        if not p:
            raise SqgError(-1, self.lib.sqg_last_error(self.h).decode())
        libc = C.CDLL(None)
        libc.free.argtypes = [C.c_void_p]
        out = dict(offset=off.value, median_before=mb.value,
                   sig=np.ctypeslib.as_array(p, shape=(max(n.value, 1),))[:n.value].copy())
        libc.free(p)
        if want_ss:
            out["ss"] = np.ctypeslib.as_array(ss, shape=(max(ss_n.value, 1),))[:ss_n.value].copy()
            libc.free(ss)
        return out

    # -- device-resident batches (bench.py)
    def dev_batch(self, bases, off, first_read_index=0, want=0):
        b = C.c_void_p()
        self._check(self.lib.sqg_dev_batch_create(self.h, len(off) - 1, bases.ctypes.data_as(C.c_void_p),
                                                  off.ctypes.data_as(C.c_void_p), first_read_index, want, C.byref(b)))
        return b

    def dev_batch_run(self, b, steps):
        t, tk = C.c_float(), C.c_float()
        self._check(self.lib.sqg_dev_batch_run(self.h, b, steps, C.byref(t), C.byref(tk)))
        return t.value, tk.value

    def dev_batch_info(self, b):
        s, k, nb, l = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int64()
        self._check(self.lib.sqg_dev_batch_info(self.h, b, C.byref(s), C.byref(k), C.byref(nb), C.byref(l)))
        return dict(samples=s.value, kmers=k.value, bases=nb.value, launches=l.value)

    def dev_batch_fetch(self, b):
        res = Result()
        self._check(self.lib.sqg_dev_batch_fetch(self.h, b, C.byref(res)))
        return self._unpack(res)

    def dev_batch_destroy(self, b):
        self.lib.sqg_dev_batch_destroy(self.h, b)

    def bench_store(self, nbytes, steps):
        t = C.c_float()
        self._check(self.lib.sqg_bench_store(self.h, nbytes, steps, C.byref(t)))
        return t.value

    def launch_count(self):
        return int(self.lib.sqg_launch_count(self.h))

    def close(self):
        if getattr(self, "h", None) and self.h:
            self.lib.sqg_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
