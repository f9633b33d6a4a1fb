"""Test-side bindings: the CPU oracle (oracle/libsqg_oracle.so), the compiled reference
(oracle/_ref/libsqref.so, optional) and the golden fixtures.  Test infrastructure only."""
import ctypes as C
import glob
import hashlib
import json
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
ZTABLE = os.path.join(ROOT, "squigulator_b200", "data", "ztable_v3.bin")

PROFILE_FIELDS = ["digitisation", "sample_rate", "bps", "range", "offset_mean", "offset_std",
                  "median_before_mean", "median_before_std", "dwell_mean", "dwell_std"]

SQ_RNA, SQ_IDEAL, SQ_IDEAL_TIME, SQ_IDEAL_AMP, SQ_PREFIX, SQ_R10 = 0x001, 0x004, 0x008, 0x010, 0x020, 0x040
RNG_PHILOX, RNG_LEGACY = 0, 1

# -x presets (reference src/sim.c:55-150), restated as data for tests that must not need oracle/_ref
PRESETS = {
    "dna-r9-min": (dict(digitisation=8192, sample_rate=4000, bps=450, range=1443.030273, offset_mean=13.7222605, offset_std=10.25279688, median_before_mean=200.815801, median_before_std=20.48933762, dwell_mean=9.0, dwell_std=4.0), 0),
    "dna-r9-prom": (dict(digitisation=2048, sample_rate=4000, bps=450, range=748.5801, offset_mean=-237.4102, offset_std=14.1575, median_before_mean=214.2890337, median_before_std=18.0127916, dwell_mean=9.0, dwell_std=4.0), 0),
    "rna-r9-min": (dict(digitisation=8192, sample_rate=3012, bps=70, range=1126.47, offset_mean=4.65491888, offset_std=4.115262472, median_before_mean=242.6584118, median_before_std=10.60230888, dwell_mean=43.0, dwell_std=35.0), SQ_RNA),
    "rna-r9-prom": (dict(digitisation=2048, sample_rate=3000, bps=70, range=548.788269, offset_mean=-231.9440589, offset_std=12.87185278, median_before_mean=238.5286796, median_before_std=21.1871794, dwell_mean=43.0, dwell_std=35.0), SQ_RNA),
    "dna-r10-prom": (dict(digitisation=2048, sample_rate=5000, bps=400, range=281.345551, offset_mean=-127.5655735, offset_std=19.377283387665, median_before_mean=189.87607393756, median_before_std=15.788097978713, dwell_mean=13.0, dwell_std=4.0), SQ_R10),
    "dna-r10-min": (dict(digitisation=8192, sample_rate=5000, bps=400, range=1536.598389, offset_mean=13.380569389019, offset_std=16.311471649012, median_before_mean=202.15407438804, median_before_std=13.406139241768, dwell_mean=13.0, dwell_std=4.0), SQ_R10),
    "rna004-prom": (dict(digitisation=2048, sample_rate=4000, bps=130, range=299.432068, offset_mean=-259.421128, offset_std=16.010841823643, median_before_mean=205.63935594369, median_before_std=8.3994882799157, dwell_mean=31.0, dwell_std=0.0), SQ_R10 | SQ_RNA),
    "rna004-min": (dict(digitisation=8192, sample_rate=4000, bps=130, range=1437.976685, offset_mean=12.47686423863, offset_std=10.442126577137, median_before_mean=205.08496731088, median_before_std=8.6671292866233, dwell_mean=31.0, dwell_std=0.0), SQ_R10 | SQ_RNA),
}


class Profile(C.Structure):
    _fields_ = [(f, C.c_double) for f in PROFILE_FIELDS]


class OracleConfig(C.Structure):
    _fields_ = [("profile", Profile), ("flags", C.c_uint32), ("kmer_size", C.c_uint32), ("num_kmer", C.c_uint32),
                ("meth", C.c_int32), ("amp_noise", C.c_float), ("seed", C.c_int64), ("rng_mode", C.c_int32),
                ("num_thread", C.c_int32)]


def build_oracle():
    so = os.path.join(ORACLE_DIR, "libsqg_oracle.so")
    src = os.path.join(ORACLE_DIR, "sqg_oracle.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.run(["make", "-C", ORACLE_DIR, "libsqg_oracle.so"], check=True, capture_output=True)
    return so


def load_oracle():
    lib = C.CDLL(build_oracle())
    lib.sqo_open.restype = C.c_void_p
    lib.sqo_open.argtypes = [C.POINTER(OracleConfig), C.POINTER(C.c_float), C.c_void_p]
    lib.sqo_close.argtypes = [C.c_void_p]
    lib.sqo_gen_sig.restype = C.c_int64
    lib.sqo_gen_sig.argtypes = [C.c_void_p, C.c_char_p, C.c_int32, C.c_int64, C.c_int, C.POINTER(C.c_double),
                                C.POINTER(C.c_double), C.POINTER(C.POINTER(C.c_int16)),
                                C.POINTER(C.POINTER(C.c_int32)), C.POINTER(C.c_int64)]
    lib.sqo_free_buf.argtypes = [C.c_void_p]
    lib.sqo_philox4x32_10.argtypes = [C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    lib.sqo_philox4x32.argtypes = [C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.c_int, C.POINTER(C.c_uint32)]
    lib.sqo_z32.restype = C.c_float
    lib.sqo_z32.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32), C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32]
    lib.sqo_fma_rz.restype = C.c_float
    lib.sqo_fma_rz.argtypes = [C.c_float, C.c_float, C.c_float]
    lib.sqo_ss_text.restype = C.c_int64
    lib.sqo_ss_text.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_void_p]
    lib.sqo_svb_zd_encode.restype = C.c_int64
    lib.sqo_svb_zd_encode.argtypes = [C.c_void_p, C.c_int64, C.c_void_p]
    lib.sqo_kmer_rank.restype = C.c_uint32
    lib.sqo_kmer_rank.argtypes = [C.c_char_p, C.c_uint32]
    lib.sqo_meth_kmer_rank.restype = C.c_uint32
    lib.sqo_meth_kmer_rank.argtypes = [C.c_char_p, C.c_uint32]
    lib.sqo_extract_read.restype = C.c_int64
    lib.sqo_extract_read.argtypes = [C.c_char_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int32, C.c_char,
                                     C.POINTER(C.c_int64), C.c_void_p]
    lib.sqo_lehmer_jump.restype = C.c_int64
    lib.sqo_lehmer_jump.argtypes = [C.c_int64, C.c_uint64]
    lib.sqo_lehmer_next.restype = C.c_double
    lib.sqo_lehmer_next.argtypes = [C.POINTER(C.c_int64)]
    lib.sqo_lehmer_normal.restype = C.c_double
    lib.sqo_lehmer_normal.argtypes = [C.POINTER(C.c_int64), C.c_double, C.c_double]
    return lib


def load_ztable():
    t = np.fromfile(ZTABLE, dtype=np.uint8)  # Z32[32768] binary32 ++ Z2[8192] binary32
    assert t.size == 32768 * 4 + 8192 * 4
    return np.ascontiguousarray(t)


def make_profile(d):
    p = Profile()
    for f in PROFILE_FIELDS:
        setattr(p, f, float(d[f]))
    return p


class Oracle:
    """One oracle handle (sqo_open .. sqo_close)."""

    def __init__(self, lib, profile, flags, k, num_kmer, model, seed, rng_mode, meth=0, amp_noise=1.0,
                 num_thread=1, ztable=None):
        self.lib = lib
        cfg = OracleConfig(make_profile(profile), flags, k, num_kmer, meth, amp_noise, seed, rng_mode, num_thread)
        self._model = np.ascontiguousarray(model, dtype=np.float32)
        assert self._model.size == 2 * num_kmer
        self._zt = ztable
        zp = ztable.ctypes.data_as(C.c_void_p) if ztable is not None else None
        self.h = lib.sqo_open(C.byref(cfg), self._model.ctypes.data_as(C.POINTER(C.c_float)), zp)
        assert self.h

    def gen_sig(self, read, read_index=0, tid=0, want_ss=False):
        off, mb = C.c_double(), C.c_double()
        sig = C.POINTER(C.c_int16)()
        ss = C.POINTER(C.c_int32)()
        ss_n = C.c_int64()
        n = self.lib.sqo_gen_sig(self.h, read, len(read), read_index, tid, C.byref(off), C.byref(mb), C.byref(sig),
                                 C.byref(ss) if want_ss else None, C.byref(ss_n))
        out = np.ctypeslib.as_array(sig, shape=(n,)).copy() if n > 0 else np.zeros(0, np.int16)
        self.lib.sqo_free_buf(sig)
        res = dict(offset=off.value, median_before=mb.value, sig=out)
        if want_ss:
            res["ss"] = np.ctypeslib.as_array(ss, shape=(ss_n.value,)).copy()
            self.lib.sqo_free_buf(ss)
        return res

    def close(self):
        if self.h:
            self.lib.sqo_close(self.h)
            self.h = None

    def __del__(self):
        self.close()


class Golden:
    """A tests/golden/*.npz fixture (made by scripts/make_golden.py from the reference's .exp files)."""

    def __init__(self, path):
        z = np.load(path)
        self.cfg = json.loads(bytes(z["cfg"]).decode())
        self.name = self.cfg["name"]
        bases, off = bytes(z["bases"]), z["base_off"]
        self.reads = [bases[off[i]:off[i + 1]] for i in range(len(off) - 1)]
        self.offset, self.median_before = z["offset"], z["median_before"]
        self.sig_len = z["sig_len"]
        self.start_time = z["start_time"]
        self.sha = bytes(z["sig_sha256"])
        self.n_full = self.cfg["n_full"]
        full, fo = z["sig_full"], np.cumsum([0] + list(self.sig_len[:self.n_full]))
        self.sig_full = [full[fo[i]:fo[i + 1]] for i in range(self.n_full)]
        self.model_rank, self.model_mean, self.model_stdv = z["model_rank"], z["model_mean"], z["model_stdv"]
        self.ss = None
        if "ss" in z.files:
            so = z["ss_off"]
            self.ss = [z["ss"][so[i]:so[i + 1]] for i in range(len(so) - 1)]

    def dense_model(self):
        m = np.zeros(2 * self.cfg["num_kmer"], dtype=np.float32)
        m[2 * self.model_rank] = self.model_mean
        m[2 * self.model_rank + 1] = self.model_stdv
        return m

    def sha_of(self, i):
        return self.sha[32 * i:32 * (i + 1)]


def sha256_i16(a):
    return hashlib.sha256(np.ascontiguousarray(a, dtype="<i2").tobytes()).digest()


def golden_paths():
    return sorted(glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def real_model(name):
    """The reference's built-in table `name` (oracle/dump_models.py wrote it through the compiled reference), or None."""
    p = os.path.join(ORACLE_DIR, "_ref", "models", name + ".f32")
    return np.fromfile(p, dtype=np.float32) if os.path.exists(p) else None


def random_model(num_kmer, seed=7):
    """Synthetic pore model of the right shape: level_mean ~ U(60,130) pA, level_stdv ~ U(1,4) pA."""
    rs = np.random.RandomState(seed)
    m = np.empty(2 * num_kmer, dtype=np.float32)
    m[0::2] = rs.uniform(60, 130, num_kmer)
    m[1::2] = rs.uniform(1.0, 4.0, num_kmer)
    return m


def random_reads(n, mean_len, seed=3, alphabet=b"ACGT", min_len=1):
    rs = np.random.RandomState(seed)
    lens = np.maximum(min_len, rs.gamma(2.0, mean_len / 2.0, n).astype(np.int64))
    al = np.frombuffer(alphabet, dtype=np.uint8)
    return [al[rs.randint(0, len(al), l)].tobytes() for l in lens]


def oracle_svb_zd(lib, sig):
    """slow5lib's svb-zd stream of one int16 signal, by the oracle's restatement"""
    sig = np.ascontiguousarray(sig, dtype=np.int16)
    out = np.empty(4 + (sig.size + 3) // 4 + 4 * sig.size + 16, dtype=np.uint8)
    n = lib.sqo_svb_zd_encode(sig.ctypes.data_as(C.c_void_p), sig.size, out.ctypes.data_as(C.c_void_p))
    return out[:n].copy()


def svb_zd_decode(buf):
    """independent decoder (numpy): the inverse of slow5lib's svb-zd, for round-trip properties"""
    buf = np.asarray(buf, dtype=np.uint8)
    n = int(buf[:4].view("<u4")[0])
    keys = buf[4:4 + (n + 3) // 4]
    codes = ((keys[:, None] >> (2 * np.arange(4))) & 3).reshape(-1)[:n].astype(np.int64)
    nb = codes + 1
    start = np.concatenate([[0], np.cumsum(nb)[:-1]]) + 4 + (n + 3) // 4
    v = np.zeros(n, dtype=np.uint64)
    for b in range(4):
        sel = nb > b
        v[sel] |= buf[start[sel] + b].astype(np.uint64) << np.uint64(8 * b)
    v = v.astype(np.int64)
    d = (v >> 1) ^ -(v & 1)
    return np.cumsum(d).astype(np.int16), int(start[-1] + nb[-1]) if n else 4


def oracle_ss_text(lib, ss, rna):
    """the `ss:Z:` value of a PAF/SAM record for one read's dwell array, by the oracle's restatement of src/format.c"""
    ss = np.ascontiguousarray(ss, dtype=np.int32)
    out = np.empty(12 * ss.size + 16, dtype=np.uint8)
    n = lib.sqo_ss_text(ss.ctypes.data_as(C.c_void_p), ss.size, int(bool(rna)), out.ctypes.data_as(C.c_void_p))
    return out[:n].tobytes()


def oracle_extract_read(lib, contig, meth, pos, length, strand, meth_state=None):
    """what gen_read() hands to gen_sig for an accepted read (oracle restatement of src/genread.c / src/seq.h).
    meth_state: None, or a ctypes c_int64 holding the rand_meth stream (advanced in place).
    Returns (read bytes, rand_meth draws taken)."""
    out = C.create_string_buffer(max(length, 1))
    m = None if meth is None else np.ascontiguousarray(meth, dtype=np.uint8)
    n = lib.sqo_extract_read(contig, len(contig), None if m is None else m.ctypes.data_as(C.c_void_p), pos, length,
                             strand.encode() if isinstance(strand, str) else strand,
                             C.byref(meth_state) if meth_state is not None else None, out)
    return out.raw[:length], int(n)


def synthetic_genome(seed=11, n_contigs=3, mean_len=3000, with_meth=True):
    """small contigs with everything the extraction has to get right: N runs, lower case, IUPAC letters, CpGs"""
    rs = np.random.RandomState(seed)
    contigs, meth = [], []
    for c in range(n_contigs):
        ln = int(mean_len * (0.6 + 0.8 * rs.rand()))
        a = np.frombuffer(b"ACGT", dtype=np.uint8)[rs.randint(0, 4, ln)].copy()
        for _ in range(3):  # N runs and single Ns
            p0 = rs.randint(0, ln - 40)
            a[p0:p0 + rs.randint(1, 30)] = ord("N")
        a[rs.randint(0, ln, 12)] = ord("N")
        lo = rs.randint(0, ln - 200)
        a[lo:lo + 150] |= 0x20  # a soft-masked stretch (lower case; an 'n' there is NOT replaced)
        a[rs.randint(0, ln, 6)] = np.frombuffer(b"RYKMSW", dtype=np.uint8)
        contigs.append(a.tobytes())
        meth.append(rs.randint(0, 256, ln).astype(np.uint8))
    return contigs, (meth if with_meth else None)
