/* oracle/ref_shim.c — TEST INFRASTRUCTURE ONLY.
 *
 * A thin handle API around the UNMODIFIED reference hot path, so that tests and the
 * CPU-baseline leg of bench.py can drive gen_sig() (reference src/gensig.c:346) directly,
 * without FASTA loading, read sampling or SLOW5 encoding around it.
 *
 * This file is ours.  It textually includes the reference's src/sim.c *where it lies*
 * (-I$(REF)/src, see oracle/Makefile) purely to reach three file-static pieces unchanged:
 *   set_profile()  src/sim.c:152   (-x presets)
 *   init_opt()     src/sim.c:197
 *   init_rand()    src/sim.c:215   (per-thread / per-k-mer stream layout)
 * Everything else (gen_sig, set_model, read_model, init_aln ...) is linked from the
 * reference's own objects.  No reference source is copied into this repository.
 */
#define _GNU_SOURCE
#include "sim.c" /* the reference's src/sim.c, found via -I$(REF)/src */

#include <pthread.h>

typedef struct {
    core_t *core;
} sqref_t;

/* fill *p and *flags from an -x preset name; returns 0, or -1 for an unknown name
 * (set_profile() itself exit()s on unknown names, so the names are screened here) */
int sqref_profile(const char *name, profile_t *p, uint32_t *flags) {
    static const char *known[] = {"dna-r9-min", "dna-r9-prom", "rna-r9-min", "rna-r9-prom",
                                  "dna-r10-min", "dna-r10-prom", "rna004-min", "rna004-prom"};
    int ok = 0;
    for (size_t i = 0; i < sizeof(known) / sizeof(known[0]); i++) ok |= (strcmp(name, known[i]) == 0);
    if (!ok) return -1;
    opt_t opt;
    init_opt(&opt);
    enum sq_log_level_opt lvl = get_log_level();
    set_log_level(LOG_OFF);
    *p = set_profile((char *)name, &opt);
    set_log_level(lvl);
    *flags = opt.flag;
    return 0;
}

/* meth: 0 = nucleotide model only, 1 = built-in CpG model for the chemistry in `flags`
 * (R10 is reached by calling set_model(MODEL_ID_DNA_R10_CPG) directly: the gate at
 * src/sim.c:310-312 lives in init_core, which the shim does not use), 2 = meth_model_file. */
void *sqref_open(const profile_t *p, uint32_t flags, int64_t seed, int32_t num_thread, float amp_noise,
                 int meth, const char *model_file, const char *meth_model_file, int verbose) {
    set_log_level(verbose ? LOG_VERB : LOG_OFF);
    core_t *core = (core_t *)calloc(1, sizeof(core_t));
    opt_t opt;
    init_opt(&opt);
    opt.flag = flags;
    opt.seed = seed;
    opt.num_thread = num_thread;
    opt.amp_noise = amp_noise;
    opt.model_file = model_file;
    opt.meth_freq = meth ? "<shim>" : NULL; /* used only as a boolean on the hot path */
    opt.meth_model_file = meth_model_file;
    core->opt = opt;
    core->profile = *p;

    /* same selection logic as init_core, src/sim.c:268-326 */
    core->model = (model_t *)malloc(sizeof(model_t) * MAX_NUM_KMER);
    uint32_t k;
    if (model_file) {
        k = read_model(core->model, model_file, MODEL_TYPE_NUCLEOTIDE);
    } else if (flags & SQ_R10) {
        k = set_model(core->model, (flags & SQ_RNA) ? MODEL_ID_RNA_RNA004_NUCLEOTIDE : MODEL_ID_DNA_R10_NUCLEOTIDE);
    } else {
        k = set_model(core->model, (flags & SQ_RNA) ? MODEL_ID_RNA_R9_NUCLEOTIDE : MODEL_ID_DNA_R9_NUCLEOTIDE);
    }
    core->kmer_size = k;
    core->num_kmer = (uint32_t)(1 << 2 * k);
    if (meth) {
        core->cpgmodel = (model_t *)malloc(sizeof(model_t) * MAX_NUM_KMER_METH);
        uint32_t km;
        if (meth == 2 && meth_model_file) {
            km = read_model(core->cpgmodel, meth_model_file, MODEL_TYPE_METH);
        } else {
            km = set_model(core->cpgmodel, (flags & SQ_R10) ? MODEL_ID_DNA_R10_CPG : MODEL_ID_DNA_R9_CPG);
        }
        if (km != k) return NULL;
        core->num_kmer = (uint32_t)pow(5, k);
    }
    init_rand(core);
    sqref_t *h = (sqref_t *)malloc(sizeof(sqref_t));
    h->core = core;
    return h;
}

uint32_t sqref_kmer_size(void *hv) { return ((sqref_t *)hv)->core->kmer_size; }
uint32_t sqref_num_kmer(void *hv) { return ((sqref_t *)hv)->core->num_kmer; }

/* copy the active (level_mean, level_stdv) table, interleaved, 2*num_kmer floats */
void sqref_get_model(void *hv, float *out) {
    core_t *core = ((sqref_t *)hv)->core;
    model_t *m = core->opt.meth_freq ? core->cpgmodel : core->model;
    for (uint32_t i = 0; i < core->num_kmer; i++) {
        out[2 * i] = m[i].level_mean;
        out[2 * i + 1] = m[i].level_stdv;
    }
}

/* one gen_sig() call on stream set `tid`.  *sig is malloc'd by the reference and owned by the
 * caller (sqref_free_buf).  If ss_out != NULL the per-k-mer dwell array (aln->ss) is returned too. */
int64_t sqref_gen_sig(void *hv, const char *read, int32_t len, int tid, double *offset, double *median_before,
                      int16_t **sig, int32_t **ss_out, int64_t *ss_n) {
    core_t *core = ((sqref_t *)hv)->core;
    int8_t rna = core->opt.flag & SQ_RNA ? 1 : 0;
    aln_t *aln = ss_out ? init_aln() : NULL;
    int64_t n = 0;
    *sig = gen_sig(core, read, len, offset, median_before, &n, rna, tid, aln);
    if (aln) {
        *ss_out = aln->ss;
        *ss_n = aln->ss_n;
        free(aln);
    }
    return n;
}

void sqref_free_buf(void *p) { free(p); }

/* ---- multi-threaded timing loop for the CPU baseline: every thread owns one tid and calls
 * gen_sig() on reads tid, tid+T, tid+2T ... (static round-robin; no SLOW5 encode, no I/O). ---- */
typedef struct {
    sqref_t *h;
    const char *bases;
    const int64_t *off;
    const int32_t *len;
    int64_t n_reads;
    int tid, nthreads;
    int64_t samples;
} sqref_job_t;

static void *sqref_worker(void *a) {
    sqref_job_t *j = (sqref_job_t *)a;
    core_t *core = j->h->core;
    int8_t rna = core->opt.flag & SQ_RNA ? 1 : 0;
    for (int64_t r = j->tid; r < j->n_reads; r += j->nthreads) {
        double o, mb;
        int64_t n = 0;
        char *tmp = strndup(j->bases + j->off[r], j->len[r]); /* gen_sig wants a NUL-terminated read */
        int16_t *s = gen_sig(core, tmp, j->len[r], &o, &mb, &n, rna, j->tid, NULL);
        j->samples += n;
        free(s);
        free(tmp);
    }
    return NULL;
}

/* returns total samples generated; nthreads must be <= the num_thread given to sqref_open */
int64_t sqref_run_batch(void *hv, const char *bases, const int64_t *off, const int32_t *len, int64_t n_reads,
                        int nthreads) {
    sqref_t *h = (sqref_t *)hv;
    if (nthreads > h->core->opt.num_thread) nthreads = h->core->opt.num_thread;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * nthreads);
    sqref_job_t *jobs = (sqref_job_t *)calloc(nthreads, sizeof(sqref_job_t));
    for (int t = 0; t < nthreads; t++) {
        jobs[t] = (sqref_job_t){h, bases, off, len, n_reads, t, nthreads, 0};
        pthread_create(&th[t], NULL, sqref_worker, &jobs[t]);
    }
    int64_t total = 0;
    for (int t = 0; t < nthreads; t++) {
        pthread_join(th[t], NULL);
        total += jobs[t].samples;
    }
    free(th);
    free(jobs);
    return total;
}

void sqref_close(void *hv) {
    sqref_t *h = (sqref_t *)hv;
    core_t *core = h->core;
    for (int i = 0; i < core->opt.num_thread; i++) {
        free_nrng(core->rand_time[i]);
        free_grng(core->rand_rlen[i]);
        free_nrng(core->rand_offset[i]);
        free_nrng(core->rand_median_before[i]);
        for (uint32_t j = 0; j < core->num_kmer; j++) free_nrng(core->kmer_gen[i][j]);
        free(core->kmer_gen[i]);
    }
    free(core->kmer_gen);
    free(core->rand_time);
    free(core->rand_rlen);
    free(core->rand_offset);
    free(core->rand_median_before);
    free(core->ref_pos);
    free(core->rand_strand);
    if (core->rand_meth) free(core->rand_meth);
    free(core->model);
    free(core->cpgmodel);
    free(core);
    free(h);
}

/* the reference's own signal compressor (slow5lib: zig-zag delta + StreamVByte), for pinning the svb-zd restatement.
 * Returns the number of bytes written to out (<= cap), or -1. */
#include <slow5/slow5_press.h>
int64_t sqref_svb_zd(const int16_t *sig, int64_t n, uint8_t *out, int64_t cap) {
    size_t bytes = 0;
    void *p = slow5_ptr_compress_solo(SLOW5_COMPRESS_SVB_ZD, sig, (size_t)n * sizeof(int16_t), &bytes);
    if (!p || (int64_t)bytes > cap) { free(p); return -1; }
    memcpy(out, p, bytes);
    free(p);
    return (int64_t)bytes;
}

/* ---- gen_read() (src/genread.c:357) over an in-memory genome, for pinning the read-extraction restatement.
 * Contig c = seq[off[c] .. off[c+1]); meth (nullable) has the same indexing.  The strings are copied. ---- */
int sqref_set_genome(void *hv, int n, const char *seq, const int64_t *off, const uint8_t *meth) {
    core_t *core = ((sqref_t *)hv)->core;
    ref_t *ref = (ref_t *)calloc(1, sizeof(ref_t));
    ref->num_ref = n;
    ref->ref_names = (char **)calloc(n, sizeof(char *));
    ref->ref_seq = (char **)calloc(n, sizeof(char *));
    ref->ref_lengths = (int32_t *)calloc(n, sizeof(int32_t));
    ref->ref_meth = meth ? (uint8_t **)calloc(n, sizeof(uint8_t *)) : NULL;
    for (int c = 0; c < n; c++) {
        int64_t len = off[c + 1] - off[c];
        char name[32];
        snprintf(name, sizeof name, "ctg%d", c);
        ref->ref_names[c] = strdup(name);
        ref->ref_seq[c] = strndup(seq + off[c], len);
        ref->ref_lengths[c] = (int32_t)len;
        ref->sum += len;
        if (meth) {
            ref->ref_meth[c] = (uint8_t *)malloc(len ? len : 1);
            memcpy(ref->ref_meth[c], meth + off[c], len);
        }
    }
    core->ref = ref;
    return 0;
}

/* one accepted read from thread 0's streams.  Returns rlen (the read is copied to buf, cap >= rlen), or -1. */
int32_t sqref_gen_read(void *hv, int32_t *contig, int32_t *ref_pos, char *strand, char *buf, int32_t cap) {
    core_t *core = ((sqref_t *)hv)->core;
    char *rid = NULL;
    int32_t ref_len = 0, rlen = 0;
    int8_t rna = core->opt.flag & SQ_RNA ? 1 : 0;
    char *seq = gen_read(core, &rid, &ref_len, ref_pos, &rlen, strand, rna, 0);
    if (!seq || rlen > cap) { free(seq); return -1; }
    memcpy(buf, seq, rlen);
    free(seq);
    *contig = -1;
    for (int c = 0; c < core->ref->num_ref; c++)
        if (core->ref->ref_names[c] == rid) *contig = c;
    return rlen;
}

/* ---- BLOW5 records by the reference's own code: slow5_open + set_header_attributes / set_header_aux_fields
 * (src/gensig.c:80-168) + slow5_set_press(NONE, SVB_ZD) + slow5_hdr_write; then per read set_record_primary_fields /
 * set_record_aux_fields (src/gensig.c:171-217) + slow5_encode (slow5lib/src/slow5.c).  For tests/test_records.py. ---- */
void set_header_attributes(slow5_file_t *sp, int8_t rna, int8_t r10, double sample_frequency);
void set_header_aux_fields(slow5_file_t *sp, int8_t ont_friendly);
void set_record_primary_fields(profile_t *profile, slow5_rec_t *slow5_record, char *read_id, double offset, int64_t len_raw_signal, int16_t *raw_signal);
void set_record_aux_fields(slow5_rec_t *slow5_record, slow5_file_t *sp, double median_before, int32_t read_number, uint64_t start_time, int8_t ont_friendly);

void *sqref_rec_open(const char *path, const profile_t *p, uint32_t flags, int ont) {
    slow5_file_t *sp = slow5_open(path, "w");
    if (!sp) return NULL;
    set_header_attributes(sp, flags & SQ_RNA ? 1 : 0, flags & SQ_R10 ? 1 : 0, p->sample_rate);
    set_header_aux_fields(sp, ont ? 1 : 0);
    if (slow5_set_press(sp, SLOW5_COMPRESS_NONE, SLOW5_COMPRESS_SVB_ZD) < 0) return NULL;
    if (slow5_hdr_write(sp) < 0) return NULL;
    return sp;
}

/* the record of one read as slow5_encode makes it; returns its size (or -1), copies at most cap bytes to out */
int64_t sqref_rec_encode(void *spv, const profile_t *p, const char *read_id, double offset, const int16_t *sig, int64_t n,
                         double median_before, int32_t read_number, uint64_t start_time, int ont, uint8_t *out, int64_t cap) {
    slow5_file_t *sp = (slow5_file_t *)spv;
    slow5_rec_t *rec = slow5_rec_init();
    char *id = strdup(read_id);
    int16_t *raw = (int16_t *)malloc(sizeof(int16_t) * (size_t)(n > 0 ? n : 1));
    memcpy(raw, sig, sizeof(int16_t) * (size_t)n);
    profile_t prof = *p;
    set_record_primary_fields(&prof, rec, id, offset, n, raw);
    set_record_aux_fields(rec, sp, median_before, read_number, start_time, ont ? 1 : 0);
    void *mem = NULL;
    size_t bytes = 0;
    if (slow5_encode(&mem, &bytes, rec, sp) < 0) return -1;
    if ((int64_t)bytes <= cap) memcpy(out, mem, bytes);
    free(mem);
    slow5_rec_free(rec);
    return (int64_t)bytes;
}

/* records made elsewhere (the GPU), written as they are */
int sqref_rec_write_bytes(void *spv, const void *mem, int64_t bytes) { return slow5_write_bytes((void *)mem, (size_t)bytes, (slow5_file_t *)spv); }
int sqref_rec_close(void *spv) { return slow5_close((slow5_file_t *)spv); }
