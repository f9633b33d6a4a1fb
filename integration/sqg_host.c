/* sqg_host.c — the reference-side binding of libsqg.so (include/sqg.h): what a squigulator maintainer adds to src/ so
 * that a batch of reads (process_db, src/sim.c:622-627) is generated on the GPU instead of through
 * work_db(core, db, work_per_single_read).  Compiled INSIDE the reference tree (it uses the reference's own headers and
 * helpers: gen_read, set_record_*_fields, paf_str/sam_str, slow5_encode); integration/sim_hook.patch adds the three
 * calls that reach it.  Nothing here generates a sample: the signal comes from sqg_gen_batch().
 *
 * Per batch:
 *   1. gen_read() for every record -> the reads' bases back to back (src/sim.c:542-549, src/genread.c:358): at -t 1
 *      serially on thread 0's streams, so that read ids and coordinates equal the CPU build's; at -t N on the
 *      reference's worker threads with their own streams, as the CPU build does
 *   2. one sqg_gen_batch() call: int16 signals, per-read offset / median_before, dwell arrays (aln->ss) when PAF/SAM
 *      output is on.  first_read_index = core->total_reads.
 *   3. in parallel (the reference's own thread pool, work_db): read id, FASTA / PAF / SAM strings, SLOW5 record
 *      encoding - what the rest of work_per_single_read does (src/sim.c:564-611) - with start_time = the exclusive
 *      prefix sum of the lengths in read order.
 *
 * Environment: SQG_GPU=1 switches the binding on; SQG_RECORD_PRESS=none writes BLOW5 records without zlib; SQG_RNG=legacy selects SQG_RNG_LEGACY (the reference's own minstd
 * streams: byte-identical output to the CPU build at -t1, used by tests/test_host_integration.py), default Philox. */
#include <assert.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "error.h"
#include "format.h"
#include "misc.h"
#include "sq.h"
#include "sqg.h"

char *gen_read(core_t *core, char **ref_id, int32_t *ref_len, int32_t *ref_pos, int32_t *rlen, char *c, int8_t rna, int tid);
void set_record_primary_fields(profile_t *profile, slow5_rec_t *slow5_record, char *read_id, double offset, int64_t len_raw_signal, int16_t *raw_signal);
void set_record_aux_fields(slow5_rec_t *slow5_record, slow5_file_t *sp, double median_before, int32_t read_number, uint64_t start_time, int8_t ont_friendly);
void work_db(core_t *core, db_t *db, void (*func)(core_t *, db_t *, int, int));
void fake_uuid(char *read_id, int64_t num);

typedef struct {          /* what step 1 learns about a read */
    char *rid, *seq;
    int32_t rlen, ref_len, ref_pos_st;
    char strand;
    int owned;            /* seq was malloc'd by gen_read */
} host_read_t;

static sqg_ctx_t *g_ctx = NULL;
static double g_t_init = 0, g_t_reads = 0, g_t_gpu = 0, g_t_records = 0;   /* wall seconds per stage (reported at exit) */
static struct {           /* the batch in flight between steps 1-3 */
    host_read_t *reads;
    sqg_result_t res;
    int64_t *start_time;
    int32_t cap;
    char *bases;
    int64_t *base_off;
    size_t bases_cap;
} g_batch;

int sqg_host_active(void) { return g_ctx != NULL; }

void sqg_host_init(core_t *core) {
    const char *on = getenv("SQG_GPU");
    if (!on || atoi(on) == 0) return;
    sqg_config_t cfg;
    memset(&cfg, 0, sizeof cfg);
    const profile_t *p = &core->profile;
    cfg.profile.digitisation = p->digitisation; cfg.profile.sample_rate = p->sample_rate; cfg.profile.bps = p->bps;
    cfg.profile.range = p->range; cfg.profile.offset_mean = p->offset_mean; cfg.profile.offset_std = p->offset_std;
    cfg.profile.median_before_mean = p->median_before_mean; cfg.profile.median_before_std = p->median_before_std;
    cfg.profile.dwell_mean = p->dwell_mean; cfg.profile.dwell_std = p->dwell_std;
    cfg.flags = core->opt.flag;            /* SQG_* are the SQ_* bits */
    cfg.kmer_size = core->kmer_size;
    cfg.meth = core->opt.meth_freq != NULL;
    cfg.num_kmer = cfg.meth ? (uint32_t)pow(5, core->kmer_size) : (uint32_t)(1u << (2 * core->kmer_size));
    cfg.amp_noise = core->opt.amp_noise;
    cfg.seed = core->opt.seed;
    const char *rng = getenv("SQG_RNG");
    cfg.rng_mode = (rng && strcmp(rng, "legacy") == 0) ? SQG_RNG_LEGACY : SQG_RNG_PHILOX;
    if (cfg.rng_mode == SQG_RNG_LEGACY && core->opt.num_thread != 1)
        WARNING("%s", "SQG_RNG=legacy reproduces the reference's streams of thread 0: equal to the CPU build at -t1 only");
    const model_t *table = cfg.meth ? core->cpgmodel : core->model;   /* model_t == sqg_model_t: two floats */
    const double t0 = realtime();
    int rc = sqg_init(&g_ctx, &cfg, (const sqg_model_t *)table);
    g_t_init = realtime() - t0;
    if (rc != SQG_OK) {
        ERROR("libsqg: %s", sqg_last_error(NULL));
        exit(EXIT_FAILURE);
    }
    VERBOSE("signal generation on the GPU: %s, %s streams", sqg_version(), cfg.rng_mode == SQG_RNG_LEGACY ? "legacy" : "philox");
}

void sqg_host_destroy(void) {
    if (g_ctx) INFO("libsqg stages: init %.2f s, reads (gen_read, serial) %.2f s, signals (sqg_gen_batch) %.2f s, records (encode, %s) %.2f s",
                    g_t_init, g_t_reads, g_t_gpu, "reference thread pool", g_t_records);
    if (g_ctx) sqg_destroy(g_ctx);
    g_ctx = NULL;
    free(g_batch.reads); free(g_batch.start_time); free(g_batch.bases); free(g_batch.base_off);
    memset(&g_batch, 0, sizeof g_batch);
}

/* step 1 on the reference's worker threads (-t N, N > 1): every worker samples with its own streams, exactly as the CPU
 * build does (src/sim.c:548) - and, like there, which read gets which worker is not reproducible from run to run */
static void sample_read(core_t *core, db_t *db, int32_t i, int tid) {
    (void)db;
    host_read_t *r = &g_batch.reads[i];
    const int8_t rna = core->opt.flag & SQ_RNA ? 1 : 0;
    r->ref_len = 0;
    r->seq = gen_read(core, &r->rid, &r->ref_len, &r->ref_pos_st, &r->rlen, &r->strand, rna, tid);
    r->owned = 1;
}

/* before the SLOW5 header is written (src/sim.c:343-352).  SQG_RECORD_PRESS=none: BLOW5 records are not zlib-compressed
 * (the signal inside them still is svb-zd) - zlib at ~13 M samples/s/core is what bounds the end-to-end run once the
 * signal comes from the GPU; slow5tools and every slow5lib reader open such files as they are. */
void sqg_host_configure_output(slow5_file_t *sp) {
    const char *on = getenv("SQG_GPU"), *rp = getenv("SQG_RECORD_PRESS");
    if (!on || atoi(on) == 0 || !rp || strcmp(rp, "none") != 0) return;
    if (sp->format != SLOW5_FORMAT_BINARY) return;
    if (slow5_set_press(sp, SLOW5_COMPRESS_NONE, SLOW5_COMPRESS_SVB_ZD) < 0) {
        ERROR("%s", "slow5_set_press failed");
        exit(EXIT_FAILURE);
    }
}

/* step 3, one record (called by the reference's worker threads) */
static void finish_read(core_t *core, db_t *db, int32_t i, int tid) {
    (void)tid;
    const host_read_t *r = &g_batch.reads[i];
    const sqg_result_t *res = &g_batch.res;
    const int8_t rna = core->opt.flag & SQ_RNA ? 1 : 0;
    const int64_t n = res->len_raw_signal[i];
    const int32_t ref_pos_end = r->ref_pos_st + r->rlen;

    /* slow5lib free()s rec->raw_signal (slow5lib/src/slow5.c:3982, :4118): it gets a heap copy of the pinned result */
    int16_t *raw = (int16_t *)malloc(sizeof(int16_t) * (size_t)(n > 0 ? n : 1));
    MALLOC_CHK(raw);
    memcpy(raw, res->signal + res->sig_off[i], sizeof(int16_t) * (size_t)n);

    char *read_id = (char *)malloc(10000);
    MALLOC_CHK(read_id);
    if (core->opt.flag & SQ_ONT) fake_uuid(read_id, core->total_reads + i + 1);
    else sprintf(read_id, "S1_%ld!%s!%d!%d!%c", (long)(core->total_reads + i + 1), r->rid, r->ref_pos_st, ref_pos_end, r->strand);

    if (core->fp_fasta) {
        db->fasta[i] = (char *)malloc(strlen(read_id) + strlen(r->seq) + 10);
        MALLOC_CHK(db->fasta[i]);
        sprintf(db->fasta[i], ">%s\n%s\n", read_id, r->seq);
    }
    if (core->fp_paf || core->fp_sam) {
        aln_t *aln = init_aln();
        const int64_t nk = res->ss_off[i + 1] - res->ss_off[i];
        aln->ss = (int32_t *)realloc(aln->ss, sizeof(int32_t) * (size_t)(nk > 0 ? nk : 1));
        MALLOC_CHK(aln->ss);
        memcpy(aln->ss, res->ss + res->ss_off[i], sizeof(int32_t) * (size_t)nk);
        aln->ss_n = aln->ss_c = nk;
        aln->sig_start = 0;
        aln->sig_end = n;
        const int64_t n_kmer = r->rlen - core->kmer_size + 1;
        assert(n_kmer > 0);
        aln->read_id = read_id;
        aln->len_raw_signal = n;
        aln->strand = r->strand;
        aln->si_st_ref = rna ? ref_pos_end - core->kmer_size + 1 : r->ref_pos_st;
        aln->si_end_ref = rna ? r->ref_pos_st : ref_pos_end - core->kmer_size + 1;
        if (core->opt.flag & SQ_PAF_REF) {
            aln->tid = r->rid;
            aln->tlen = !(core->opt.flag & SQ_FULL_CONTIG) ? r->ref_len - core->kmer_size + 1 : n_kmer;
            aln->t_st = aln->si_st_ref;
            aln->t_end = aln->si_end_ref;
        } else {
            aln->tid = read_id;
            aln->tlen = n_kmer;
            aln->t_st = rna ? n_kmer : 0;
            aln->t_end = rna ? 0 : n_kmer;
        }
        if (core->fp_paf) db->paf[i] = paf_str(aln);
        if (core->fp_sam) db->sam[i] = sam_str(aln, r->seq, r->rid, r->ref_pos_st);
        free_aln(aln);
    }

    slow5_rec_t *rec = slow5_rec_init();
    if (rec == NULL) {
        ERROR("%s", "Could not allocate space for a slow5 record.");
        exit(EXIT_FAILURE);
    }
    set_record_primary_fields(&core->profile, rec, read_id, res->offset[i], n, raw);
    set_record_aux_fields(rec, core->sp, res->median_before[i], core->total_reads + i, (uint64_t)g_batch.start_time[i],
                          core->opt.flag & SQ_ONT ? 1 : 0);
    if (slow5_encode(&db->mem_records[i], &db->mem_bytes[i], rec, core->sp) < 0) {
        ERROR("%s", "Error encoding record");
        exit(EXIT_FAILURE);
    }
    slow5_rec_free(rec);   /* frees read_id and raw with it */
}

void sqg_host_process_db(core_t *core, db_t *db) {
    const int32_t n = db->n_rec;
    const int8_t rna = core->opt.flag & SQ_RNA ? 1 : 0;
    if (n > g_batch.cap) {
        g_batch.reads = (host_read_t *)realloc(g_batch.reads, sizeof(host_read_t) * (size_t)n);
        g_batch.start_time = (int64_t *)realloc(g_batch.start_time, sizeof(int64_t) * (size_t)n);
        g_batch.base_off = (int64_t *)realloc(g_batch.base_off, sizeof(int64_t) * ((size_t)n + 1));
        MALLOC_CHK(g_batch.reads); MALLOC_CHK(g_batch.start_time); MALLOC_CHK(g_batch.base_off);
        g_batch.cap = n;
    }
    /* 1. the reads, in order, on thread 0's streams */
    double t0 = realtime();
    size_t total = 0;
    const int parallel_sampling = core->opt.num_thread > 1 && !(core->opt.flag & SQ_FULL_CONTIG);
    if (parallel_sampling) work_db(core, db, sample_read);
    for (int32_t i = 0; i < n; i++) {
        host_read_t *r = &g_batch.reads[i];
        if (parallel_sampling) {
            /* sampled above */
        } else if (core->opt.flag & SQ_FULL_CONTIG) {
            r->ref_len = 0;
            r->rid = core->ref->ref_names[core->total_reads + i];
            r->rlen = core->ref->ref_lengths[core->total_reads + i];
            r->seq = core->ref->ref_seq[core->total_reads + i];
            r->strand = '+';
            r->ref_pos_st = 0;
            r->owned = 0;
        } else {
            r->ref_len = 0;
            r->seq = gen_read(core, &r->rid, &r->ref_len, &r->ref_pos_st, &r->rlen, &r->strand, rna, 0);
            r->owned = 1;
        }
        g_batch.base_off[i] = (int64_t)total;
        total += (size_t)r->rlen;
    }
    g_batch.base_off[n] = (int64_t)total;
    if (total + 1 > g_batch.bases_cap) {
        g_batch.bases_cap = total + total / 2 + 1;
        g_batch.bases = (char *)realloc(g_batch.bases, g_batch.bases_cap);
        MALLOC_CHK(g_batch.bases);
    }
    for (int32_t i = 0; i < n; i++) memcpy(g_batch.bases + g_batch.base_off[i], g_batch.reads[i].seq, (size_t)g_batch.reads[i].rlen);

    g_t_reads += realtime() - t0;
    t0 = realtime();
    /* 2. the signals */
    const uint32_t want = (core->fp_paf || core->fp_sam) ? SQG_WANT_SS : 0;
    int rc = sqg_gen_batch(g_ctx, n, g_batch.bases, g_batch.base_off, core->total_reads, want, &g_batch.res);
    if (rc != SQG_OK) {
        ERROR("libsqg: %s", sqg_last_error(g_ctx));
        exit(EXIT_FAILURE);
    }
    g_t_gpu += realtime() - t0;
    t0 = realtime();
    /* start_time = samples before the read, in read order (src/sim.c:602: equal at -t1) */
    for (int32_t i = 0; i < n; i++) {
        g_batch.start_time[i] = core->n_samples;
        core->n_samples += g_batch.res.len_raw_signal[i];
    }
    /* 3. records and text lines, on the reference's own worker threads */
    work_db(core, db, finish_read);
    g_t_records += realtime() - t0;
    for (int32_t i = 0; i < n; i++)
        if (g_batch.reads[i].owned) free(g_batch.reads[i].seq);
}
