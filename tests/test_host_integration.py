"""The drop-in proof with the real host (missing 3): the reference's own CLI, built with
integration/sim_hook.patch + integration/sqg_host.c so that process_db() (src/sim.c:622-627) routes every batch through
libsqg.so, must reproduce the reference's golden files (scripts/test.sh) byte for byte when the library runs the
reference's own random streams (SQG_RNG=legacy) - SLOW5 records, FASTA, PAF and SAM lines included.

The binary and the reference's test inputs are staged under oracle/_ref/ by integration/build_patched.sh (run by
__graft_entry__.build() where /root/reference exists; both travel to the GPU box)."""
import os
import shlex
import subprocess

import pytest

from tests import helpers as H

REF = os.path.join(H.ROOT, "oracle", "_ref")
BIN = os.path.join(REF, "squigulator_sqg")
needs_bin = pytest.mark.skipif(not (os.path.exists(BIN) and os.path.exists(os.path.join(REF, "test", "slow5.exp"))),
                               reason="oracle/_ref/squigulator_sqg not built (integration/build_patched.sh)")

# scripts/test.sh of the reference, as data: (arguments, {output file: golden})
NCOV, SEQUIN = "test/nCoV-2019.reference.fasta", "test/rnasequin_sequences_2.4.fa"
CASES = [
    ("basic_dna", f"{NCOV} -o a.slow5 -q a.fasta -n 10 --seed 1 --dwell-std 1.0 -r 20000 -t1", {"a.fasta": "fasta.exp", "a.slow5": "slow5.exp"}),
    ("basic_rna", f"-x rna-r9-prom {SEQUIN} -o a.slow5 -q a.fastq -n 10 --seed 1 --prefix=yes --dwell-std 3.0 -t1", {"a.slow5": "rna_slow5.exp"}),
    ("ideal", f"{NCOV} -o a.slow5 -n 2 --seed 1 --ideal -r 20000 -t1", {"a.slow5": "dna_ideal_slow5.exp"}),
    ("ideal_time", f"{NCOV} -o a.slow5 -n 2 --seed 1 --ideal-time -r 20000 -t1", {"a.slow5": "dna_ideal_time_slow5.exp"}),
    ("ideal_amp", f"{NCOV} -o a.slow5 -n 2 --seed 1 --ideal-amp -r 20000 --dwell-std 5.0 -t1", {"a.slow5": "dna_ideal_amp_slow5.exp"}),
    ("amp_noise_0", f"{NCOV} -o a.slow5 -n 2 --seed 1 --amp-noise 0.0 -r 20000 --dwell-std 5.0 -t1", {"a.slow5": "dna_ideal_amp_slow5.exp"}),
    ("dna_prefix", f"{NCOV} -o a.slow5 -n 2 --seed 1 --prefix=yes -r 20000 --dwell-std 5.0 -t1", {"a.slow5": "dna_prefix_slow5.exp"}),
    ("rna_prefix_yes", f"-x rna-r9-prom {SEQUIN} -o a.slow5 -n 2 --seed 1 --dwell-std 3.0 -t1 --prefix=yes", {"a.slow5": "rna_prefixyes_slow5.exp"}),
    ("rna_prefix_no", f"-x rna-r9-prom {SEQUIN} -o a.slow5 -n 2 --seed 1 --dwell-std 3.0 -t1", {"a.slow5": "rna_prefixno_slow5.exp"}),
    ("full_contigs", f"{NCOV} -o a.slow5 --seed 1 --full-contigs --dwell-std 5.0 -t1", {"a.slow5": "dna_full_contig.exp"}),
    ("r10_paf", f"-x dna-r10-prom -o a.slow5 -n 1 --seed 1 --dwell-std 4.0 -t1 {NCOV} -c a.paf -q a.fa",
     {"a.slow5": "dna_r10_paf.exp", "a.paf": "dna_r10_paf.paf.exp", "a.fa": "dna_r10_paf.fa.exp"}),
    ("rna_paf_sam", f"-x rna-r9-prom -o a.slow5 -n 1 --seed 1 --dwell-std 3.0 -t1 -t1 {SEQUIN} -c a.paf -q a.fa -a a.sam",
     {"a.slow5": "rna_paf.exp", "a.paf": "rna_paf.paf.exp", "a.sam": "rna_paf.sam.exp", "a.fa": "rna_paf.fa.exp"}),
    ("r10_paf_ref", f"-x dna-r10-prom -o a.slow5 -n 2 --seed 2 --dwell-std 4.0 -t1 {NCOV} -c a.paf --paf-ref -a a.sam",
     {"a.slow5": "dna_r10_paf-ref.exp", "a.paf": "dna_r10_paf-ref.paf.exp", "a.sam": "dna_r10_paf-ref.sam.exp"}),
    ("r10_sam_only", f"-x dna-r10-prom -o a.slow5 -n 2 --seed 2 --dwell-std 4.0 -t1 {NCOV} -c a.paf -a a.sam",
     {"a.slow5": "dna_r10_paf-ref.exp", "a.sam": "dna_r10_paf-ref.sam.exp"}),
    ("rna004", f"-x rna004-prom -o a.slow5 -n 1 --seed 1 --dwell-std 3.0 -t1 {SEQUIN}", {"a.slow5": "rna004.slow5.exp"}),
    ("r10_amp_noise", f"-x dna-r10-prom -o a.slow5 -r 20000 -f 1 --seed 2 --amp-noise 0.5 -t1 {NCOV}", {"a.slow5": "dna_r10_amp_noise.exp"}),
    ("rna004_dwell", f"-x rna004-min -o a.slow5 -n 1 --seed 1 --dwell-mean 30 --dwell-std 3.0 -t1 {SEQUIN}", {"a.slow5": "rna004_dwell.exp"}),
    ("bps", f"-x dna-r10-prom -o a.slow5 --seed 1 --bps 200 -t1 -n 2 {NCOV}", {"a.slow5": "bps.exp"}),
    ("cdna", f"-x dna-r10-min -o a.slow5 -n 1 --seed 1 --dwell-std 3.0 -t1 {SEQUIN} --cdna", {"a.slow5": "cdna.exp"}),
    ("trans_count", f"-x rna004-prom -o a.slow5 -n 3 --seed 3 --trans-count test/sequin_count.tsv -t1 {SEQUIN}", {"a.slow5": "trans_count.exp"}),
    ("trans_count_cdna", f"-x dna-r10-min -o a.slow5 -n 3 --seed 3 --trans-count test/sequin_count.tsv -t1 {SEQUIN} --cdna", {"a.slow5": "trans_count_cdna.exp"}),
    ("trans_trunc", f"-x rna004-prom -o a.slow5 -n 1 --seed 1 --trans-trunc -t1 {SEQUIN}", {"a.slow5": "trans_trunc.exp"}),
    ("ont_friendly", f"-x dna-r10-min -o a.slow5 -n 1 --seed 1 -t1 {SEQUIN} --ont-friendly=yes", {"a.slow5": "ont_friendly.exp"}),
    ("dev", f"-x dna-r10-min -o a.slow5 -n 1 --seed 1 -t1 {SEQUIN} --digitisation 4096 --sample-rate 10000 --range 300 "
            "--offset-mean -1000 --offset-std 0 --median-before-mean 100 --median-before-std 0", {"a.slow5": "dev.exp"}),
    ("r9_meth", f"-x dna-r9-prom -o a.slow5 --seed 1 -t1 -n 2 -r 29000 {NCOV} --meth-freq test/mfreq.tsv", {"a.slow5": "r9_mfreq.exp"}),
]


def run_case(args, outs, tmp_path, gpu):
    env = dict(os.environ)
    env.pop("SQG_GPU", None)
    if gpu:
        env.update(SQG_GPU="1", SQG_RNG="legacy")
    # the goldens name inputs as test/...: run from a directory that has that link, outputs next to it
    link = tmp_path / "test"
    if not link.exists():
        os.symlink(os.path.join(REF, "test"), link)
    r = subprocess.run([BIN] + shlex.split(args), cwd=tmp_path, env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    for out, exp in outs.items():
        got, want = (tmp_path / out).read_bytes(), open(os.path.join(REF, "test", exp), "rb").read()
        assert got == want, f"{out} differs from test/{exp} ({len(got)} vs {len(want)} bytes)"


@needs_bin
def test_patched_binary_without_the_switch_is_the_reference(tmp_path):
    """no SQG_GPU in the environment: the hook is inert, the CPU path runs (no GPU needed)"""
    name, args, outs = CASES[0]
    run_case(args, outs, tmp_path, gpu=False)


@needs_bin
@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_reference_goldens_through_the_gpu(case, tmp_path):
    name, args, outs = case
    run_case(args, outs, tmp_path, gpu=True)
