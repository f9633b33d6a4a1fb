#!/usr/bin/env python3
"""Build tests/golden/*.npz from the reference's own golden files.

Runs in the BUILD container only (needs /root/reference and oracle/_ref/, see oracle/Makefile):
for every command line of the reference's scripts/test.sh it

  1. runs the UNMODIFIED reference CLI (oracle/_ref/squigulator) with the same arguments plus
     `-q reads.fa`, and insists that the SLOW5 it writes is byte-identical to the shipped
     test/<name>.exp (so each fixture is the reference's golden, not merely a re-run);
  2. records what the hot path saw and produced: the sampled read strings (bases exactly as handed
     to gen_sig, incl. 'M' marks), the effective profile/flags/seed, the pore-model entries those
     reads touch (sparse: rank, level_mean, level_stdv — pulled out of oracle/_ref/libsqref.so,
     i.e. out of the reference's compiled-in tables), per-read offset/median_before/len, a sha256
     of every read's int16 signal, the full int16 signal of the leading reads (up to ~SIG_CAP
     samples), and the per-k-mer dwell arrays where the golden has a PAF.

Nothing under /root/reference is copied: fixtures hold reference OUTPUTS plus the inputs needed
to regenerate them through another implementation.  The GPU box never needs /root/reference.
"""
import ctypes as C
import hashlib
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("SQ_REFERENCE", "/root/reference")
BIN = os.path.join(ROOT, "oracle", "_ref", "squigulator")
LIB = os.path.join(ROOT, "oracle", "_ref", "libsqref.so")
OUT = os.path.join(ROOT, "tests", "golden")
SIG_CAP = 120_000

NCOV = "test/nCoV-2019.reference.fasta"
SEQUIN = "test/rnasequin_sequences_2.4.fa"

# (name, golden .exp, argv as in scripts/test.sh (without -o/-q), optional paf golden)
CASES = [
    ("dna_basic", "test/slow5.exp", [NCOV, "-n", "10", "--seed", "1", "--dwell-std", "1.0", "-r", "20000", "-t1"], None),
    ("rna_basic_prefix", "test/rna_slow5.exp", ["-x", "rna-r9-prom", SEQUIN, "-n", "10", "--seed", "1", "--prefix=yes", "--dwell-std", "3.0", "-t1"], None),
    ("dna_ideal", "test/dna_ideal_slow5.exp", [NCOV, "-n", "2", "--seed", "1", "--ideal", "-r", "20000", "-t1"], None),
    ("dna_ideal_time", "test/dna_ideal_time_slow5.exp", [NCOV, "-n", "2", "--seed", "1", "--ideal-time", "-r", "20000", "-t1"], None),
    ("dna_ideal_amp", "test/dna_ideal_amp_slow5.exp", [NCOV, "-n", "2", "--seed", "1", "--ideal-amp", "-r", "20000", "--dwell-std", "5.0", "-t1"], None),
    ("dna_amp_noise0", "test/dna_ideal_amp_slow5.exp", [NCOV, "-n", "2", "--seed", "1", "--amp-noise", "0.0", "-r", "20000", "--dwell-std", "5.0", "-t1"], None),
    ("dna_prefix", "test/dna_prefix_slow5.exp", [NCOV, "-n", "2", "--seed", "1", "--prefix=yes", "-r", "20000", "--dwell-std", "5.0", "-t1"], None),
    ("rna_prefix_yes", "test/rna_prefixyes_slow5.exp", ["-x", "rna-r9-prom", SEQUIN, "-n", "2", "--seed", "1", "--dwell-std", "3.0", "-t1", "--prefix=yes"], None),
    ("rna_prefix_no", "test/rna_prefixno_slow5.exp", ["-x", "rna-r9-prom", SEQUIN, "-n", "2", "--seed", "1", "--dwell-std", "3.0", "-t1"], None),
    ("dna_full_contig", "test/dna_full_contig.exp", [NCOV, "--seed", "1", "--full-contigs", "--dwell-std", "5.0", "-t1"], None),
    ("dna_r10_paf", "test/dna_r10_paf.exp", ["-x", "dna-r10-prom", "-n", "1", "--seed", "1", "--dwell-std", "4.0", "-t1", NCOV], "test/dna_r10_paf.paf.exp"),
    ("rna_r9_paf", "test/rna_paf.exp", ["-x", "rna-r9-prom", "-n", "1", "--seed", "1", "--dwell-std", "3.0", "-t1", SEQUIN], "test/rna_paf.paf.exp"),
    ("dna_r10_seed2", "test/dna_r10_paf-ref.exp", ["-x", "dna-r10-prom", "-n", "2", "--seed", "2", "--dwell-std", "4.0", "-t1", NCOV], None),
    ("rna004", "test/rna004.slow5.exp", ["-x", "rna004-prom", "-n", "1", "--seed", "1", "--dwell-std", "3.0", "-t1", SEQUIN], None),
    ("dna_r10_amp_noise", "test/dna_r10_amp_noise.exp", ["-x", "dna-r10-prom", "-r", "20000", "-f", "1", "--seed", "2", "--amp-noise", "0.5", "-t1", NCOV], None),
    ("rna004_dwell", "test/rna004_dwell.exp", ["-x", "rna004-min", "-n", "1", "--seed", "1", "--dwell-mean", "30", "--dwell-std", "3.0", "-t1", SEQUIN], None),
    ("dna_r10_bps", "test/bps.exp", ["-x", "dna-r10-prom", "--seed", "1", "--bps", "200", "-t1", "-n", "2", NCOV], None),
    ("cdna", "test/cdna.exp", ["-x", "dna-r10-min", "-n", "1", "--seed", "1", "--dwell-std", "3.0", "-t1", SEQUIN, "--cdna"], None),
    ("trans_count", "test/trans_count.exp", ["-x", "rna004-prom", "-n", "3", "--seed", "3", "--trans-count", "test/sequin_count.tsv", "-t1", SEQUIN], None),
    ("trans_trunc", "test/trans_trunc.exp", ["-x", "rna004-prom", "-n", "1", "--seed", "1", "--trans-trunc", "-t1", SEQUIN], None),  # as in scripts/test.sh:127: "-t1" is swallowed as the option's argument
    ("dev_adc", "test/dev.exp", ["-x", "dna-r10-min", "-n", "1", "--seed", "1", "-t1", SEQUIN, "--digitisation", "4096", "--sample-rate", "10000", "--range", "300", "--offset-mean", "-1000", "--offset-std", "0", "--median-before-mean", "100", "--median-before-std", "0"], None),
    ("r9_meth", "test/r9_mfreq.exp", ["-x", "dna-r9-prom", "--seed", "1", "-t1", "-n", "2", "-r", "29000", NCOV, "--meth-freq", "test/mfreq.tsv"], None),
    # reachable in the reference only with --meth-model (src/sim.c:310-312 gates the built-in table off);
    # the table file is dumped from the reference's own compiled-in R10 CpG table below
    ("r10_meth", "test/r10_methfreq.exp", ["-x", "dna-r10-prom", "--seed", "1", "-t1", "-n", "2", "-r", "29000", NCOV, "--meth-freq", "test/methfreq.tsv", "--meth-model", "@R10CPG@"], None),
]

PROFILE_FIELDS = ["digitisation", "sample_rate", "bps", "range", "offset_mean", "offset_std",
                  "median_before_mean", "median_before_std", "dwell_mean", "dwell_std"]


class Profile(C.Structure):
    _fields_ = [(f, C.c_double) for f in PROFILE_FIELDS]


def load_ref():
    lib = C.CDLL(LIB)
    lib.sqref_profile.argtypes = [C.c_char_p, C.POINTER(Profile), C.POINTER(C.c_uint32)]
    lib.sqref_open.restype = C.c_void_p
    lib.sqref_open.argtypes = [C.POINTER(Profile), C.c_uint32, C.c_int64, C.c_int32, C.c_float, C.c_int,
                               C.c_char_p, C.c_char_p, C.c_int]
    lib.sqref_get_model.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
    lib.sqref_kmer_size.argtypes = [C.c_void_p]
    lib.sqref_kmer_size.restype = C.c_uint32
    lib.sqref_num_kmer.argtypes = [C.c_void_p]
    lib.sqref_num_kmer.restype = C.c_uint32
    lib.sqref_close.argtypes = [C.c_void_p]
    return lib


def effective_config(lib, argv):
    """Mirror of the option handling in the reference's sim_main (src/sim.c:890-1045) for the
    options the hot path consumes."""
    prof_name = "dna-r9-prom"
    if "-x" in argv:
        prof_name = argv[argv.index("-x") + 1]
    p = Profile()
    flags = C.c_uint32(0)
    assert lib.sqref_profile(prof_name.encode(), C.byref(p), C.byref(flags)) == 0
    flags = flags.value
    seed, amp_noise, meth = 0, 1.0, 0
    gvn = set()
    it = iter(range(len(argv)))
    for i in it:
        a = argv[i]

        def val():
            next(it)
            return argv[i + 1]

        if a == "--seed": seed = int(val())
        elif a == "--ideal": flags |= 0x004
        elif a == "--ideal-time": flags |= 0x008
        elif a == "--ideal-amp": flags |= 0x010
        elif a == "--prefix=yes": flags |= 0x020
        elif a == "--dwell-std": p.dwell_std = float(val())
        elif a == "--dwell-mean": p.dwell_mean = float(val()); gvn.add("dwell_mean")
        elif a == "--amp-noise": amp_noise = float(val())
        elif a == "--digitisation": p.digitisation = float(val())
        elif a == "--sample-rate": p.sample_rate = float(val()); gvn.add("sample_rate")
        elif a == "--range": p.range = float(val())
        elif a == "--offset-mean": p.offset_mean = float(val())
        elif a == "--offset-std": p.offset_std = float(val())
        elif a == "--bps": p.bps = float(val()); gvn.add("bps")
        elif a == "--median-before-mean": p.median_before_mean = float(val())
        elif a == "--median-before-std": p.median_before_std = float(val())
        elif a == "--meth-freq": meth = 1; val()
        elif a in ("-x", "-n", "-r", "-f", "--trans-count", "--meth-model", "--trans-trunc"): val()
    if "dwell_mean" in gvn:
        p.bps = float(round(p.sample_rate / p.dwell_mean))
    if "sample_rate" in gvn or "bps" in gvn:
        p.dwell_mean = float(np.floor(p.sample_rate / p.bps + 0.5))
    return p, flags, seed, amp_noise, meth


def rank4(s):
    lut = np.zeros(256, dtype=np.uint32)
    for ch in "CcYB": lut[ord(ch)] = 1
    for ch in "GgSK": lut[ord(ch)] = 2
    for ch in "TtU": lut[ord(ch)] = 3
    return lut[np.frombuffer(s, dtype=np.uint8)]


def rank5(s):
    lut = np.zeros(256, dtype=np.uint32)
    for j, ch in enumerate("ACGMT"): lut[ord(ch)] = j
    return lut[np.frombuffer(s, dtype=np.uint8)]


def kmer_ranks(seq, k, meth):
    d = rank5(seq) if meth else rank4(seq)
    base = 5 if meth else 4
    n = len(seq) - k + 1
    r = np.zeros(n, dtype=np.uint64)
    for j in range(k):
        r = r * base + d[j:j + n]
    return r.astype(np.uint32)


POLYA = b"A" * 158
ADAPTOR_DNA = b"GGCGTCTGCTTGGGTGTTTAACCTTTTTTTTTTAATGTACTTCGTTCAGTTACGTATTGCT"
ADAPTOR_RNA = b"TGATGATGAGGGATAGACGATGGTTGTTTCTGTTGGTGCTGATATTGCTTTTTTTTTTTTTATGATGCAAGATACGCAC"
STALL_DNA = b"TTTTTTTTTTTTTTTTTTAATCAA"
STALL_RNA = b"AAAAAGAAAAAACCCCCCCCCCCCCCCCCC"


def touched_ranks(reads, k, meth, flags):
    segs = []
    for r in reads:
        if flags & 0x020:
            if flags & 0x001:
                segs += [r + POLYA + ADAPTOR_RNA, STALL_RNA]
            else:
                segs += [STALL_DNA + ADAPTOR_DNA + r]
        else:
            segs.append(r)
    segs.append(b"ACGTACGTACGTAAAA")  # short-read hack k-mers
    allr = np.concatenate([kmer_ranks(s, k, meth) for s in segs if len(s) >= k])
    return np.unique(allr)


def parse_slow5(path):
    recs = []
    hdr = None
    with open(path, "rb") as f:
        for line in f:
            if line.startswith(b"#read_id"):
                hdr = line[1:].rstrip(b"\n").split(b"\t")
                continue
            if line[:1] in (b"#", b"@"):
                continue
            cols = line.rstrip(b"\n").split(b"\t")
            d = dict(zip(hdr, cols))
            sig = np.array(d[b"raw_signal"].split(b","), dtype=np.int16)
            assert len(sig) == int(d[b"len_raw_signal"])
            recs.append(dict(read_id=d[b"read_id"].decode(), offset=float(d[b"offset"]),
                             median_before=float(d[b"median_before"]), sig=sig,
                             start_time=int(d[b"start_time"]), read_number=int(d[b"read_number"])))
    return recs


def parse_fasta(path):
    names, seqs = [], []
    with open(path, "rb") as f:
        for line in f:
            line = line.rstrip(b"\n")
            if line.startswith(b">"):
                names.append(line[1:].decode())
                seqs.append(b"")
            else:
                seqs[-1] += line
    return names, seqs


def parse_paf_ss(path, rna):
    out = []
    with open(path) as f:
        for line in f:
            ss = [t for t in line.rstrip("\n").split("\t") if t.startswith("ss:Z:")][0][5:]
            v = np.array([int(x) for x in ss.split(",") if x], dtype=np.int32)
            out.append(v[::-1].copy() if rna else v)  # paf_str prints RNA dwell arrays reversed (src/format.c:70-74)
    return out


def write_ss_text(out_dir):
    """tests/golden/ss_text.json: the `ss:Z:` values of the reference's own PAF goldens, verbatim (src/format.c:69-75),
    keyed by fixture name - pins the dwell-string format (SQG_WANT_SS_TEXT, oracle sqo_ss_text)."""
    import json
    d = {}
    for name, _, _, paf_exp in CASES:
        if paf_exp:
            d[name] = [[t for t in line.rstrip("\n").split("\t") if t.startswith("ss:Z:")][0][5:]
                       for line in open(os.path.join(REF, paf_exp))]
    json.dump(d, open(os.path.join(out_dir, "ss_text.json"), "w"))
    return d


def dump_meth_model(lib, path):
    """Write the reference's compiled-in R10 CpG 9-mer table as an f5c-style model file so the CLI can
    load it through --meth-model (read_model, src/model.c:40-142)."""
    p = Profile()
    fl = C.c_uint32(0)
    lib.sqref_profile(b"dna-r10-prom", C.byref(p), C.byref(fl))
    h = lib.sqref_open(C.byref(p), fl.value, 1, 1, 1.0, 1, None, None, 0)
    n = lib.sqref_num_kmer(h)
    m = np.zeros(2 * n, dtype=np.float32)
    lib.sqref_get_model(h, m.ctypes.data_as(C.POINTER(C.c_float)))
    lib.sqref_close(h)
    alpha = "ACGMT"
    with open(path, "w") as f:
        f.write("#k\t9\nkmer\tlevel_mean\tlevel_stdv\tsd_mean\tsd_stdv\n")
        idx = np.arange(n)
        digs = [(idx // 5 ** (8 - j)) % 5 for j in range(9)]
        for i in range(n):
            km = "".join(alpha[digs[j][i]] for j in range(9))
            f.write(f"{km}\t{float(m[2 * i]):.9g}\t{float(m[2 * i + 1]):.9g}\t0\t0\n")


def main():
    write_ss_text(OUT)
    only = set(sys.argv[1:])
    lib = load_ref()
    os.makedirs(OUT, exist_ok=True)
    tmp = tempfile.mkdtemp(prefix="sqgold")
    meth_model = os.path.join(tmp, "r10cpg.model")
    for name, exp, argv, paf_exp in CASES:
        if only and name not in only:
            continue
        argv = list(argv)
        if "@R10CPG@" in argv:
            if not os.path.exists(meth_model):
                dump_meth_model(lib, meth_model)
            argv[argv.index("@R10CPG@")] = meth_model
        s5, fa, paf = (os.path.join(tmp, name + e) for e in (".slow5", ".fa", ".paf"))
        cmd = [BIN] + argv + ["-o", s5, "-q", fa] + (["-c", paf] if paf_exp else [])
        subprocess.run(cmd, cwd=REF, check=True, stderr=subprocess.DEVNULL)
        assert open(s5, "rb").read() == open(os.path.join(REF, exp), "rb").read(), f"{name}: differs from {exp}"
        if paf_exp:
            assert open(paf, "rb").read() == open(os.path.join(REF, paf_exp), "rb").read()

        p, flags, seed, amp_noise, meth = effective_config(lib, argv)
        h = lib.sqref_open(C.byref(p), flags, seed, 1, amp_noise, meth, None,
                           meth_model.encode() if "--meth-model" in argv else None, 0)
        k, num_kmer = lib.sqref_kmer_size(h), lib.sqref_num_kmer(h)
        model = np.zeros(2 * num_kmer, dtype=np.float32)
        lib.sqref_get_model(h, model.ctypes.data_as(C.POINTER(C.c_float)))
        lib.sqref_close(h)

        recs = parse_slow5(s5)
        names, seqs = parse_fasta(fa)
        assert [r["read_id"] for r in recs] == names
        ranks = touched_ranks(seqs, k, meth, flags)
        ranks = ranks[ranks < num_kmer]

        n_full, acc = 0, 0
        for r in recs:
            if n_full and acc + len(r["sig"]) > SIG_CAP:
                break
            n_full += 1
            acc += len(r["sig"])
        cfg = dict(name=name, golden=exp, argv=argv if "--meth-model" not in argv else argv[:-1] + ["<dump of built-in R10 CpG table>"],
                   profile={f: getattr(p, f) for f in PROFILE_FIELDS}, flags=flags, seed=seed,
                   amp_noise=amp_noise, meth=meth, kmer_size=k, num_kmer=num_kmer, n_reads=len(recs), n_full=n_full)
        arrays = dict(
            cfg=np.frombuffer(json.dumps(cfg).encode(), dtype=np.uint8),
            bases=np.frombuffer(b"".join(seqs), dtype=np.uint8),
            base_off=np.cumsum([0] + [len(s) for s in seqs]).astype(np.int64),
            model_rank=ranks.astype(np.uint32),
            model_mean=model[0::2][ranks], model_stdv=model[1::2][ranks],
            offset=np.array([r["offset"] for r in recs]), median_before=np.array([r["median_before"] for r in recs]),
            sig_len=np.array([len(r["sig"]) for r in recs], dtype=np.int64),
            start_time=np.array([r["start_time"] for r in recs], dtype=np.int64),
            sig_sha256=np.frombuffer(b"".join(hashlib.sha256(r["sig"].astype("<i2").tobytes()).digest() for r in recs), dtype=np.uint8),
            sig_full=np.concatenate([r["sig"] for r in recs[:n_full]]).astype(np.int16),
        )
        if paf_exp:
            ss = parse_paf_ss(paf, bool(flags & 1))
            arrays["ss"] = np.concatenate(ss).astype(np.int32)
            arrays["ss_off"] = np.cumsum([0] + [len(x) for x in ss]).astype(np.int64)
        path = os.path.join(OUT, name + ".npz")
        np.savez_compressed(path, **arrays)
        print(f"{name:20s} reads={len(recs):3d} full={n_full} samples={int(arrays['sig_len'].sum()):8d} "
              f"k={k} ranks={len(ranks):6d} -> {os.path.getsize(path) / 1024:.0f} KB")


if __name__ == "__main__":
    main()
