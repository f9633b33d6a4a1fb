/* sqg.h — C ABI of libsqg.so: the B200 (sm_100a) signal-generation path of squigulator.
 *
 * Drop-in boundary.  The reference (hasindu2008/squigulator, paths below relative to its root)
 * has no plugin layer; the seam this library replaces is the plain C call
 *
 *     int16_t *gen_sig(core_t*, const char *read, int32_t len, double *offset,
 *                      double *median_before, int64_t *len_raw_signal,
 *                      int8_t rna, int tid, aln_t *aln)            src/gensig.c:346
 *
 * made once per read from work_per_single_read() (src/sim.c:557) under the per-batch fork-join
 * work_db() (src/thread.c:119, called by process_db(), src/sim.c:622).  Everything gen_sig computes
 * (src/gensig.c:226-356 with src/seq.h, src/rand.h and the model tables) runs here as CUDA kernels
 * over a whole batch of reads; read sampling, FASTA/SLOW5 I/O and the CLI stay with the caller.
 *
 * Plain C types only; no CUDA or torch types cross this boundary.  Every call returns SQG_OK (0) or
 * a negative error code — never exit()s, unlike the reference's MALLOC_CHK/ERROR paths.
 * INTEGRATION.md shows the reference-side binding.
 */
#ifndef SQG_H
#define SQG_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SQG_OK 0
#define SQG_ERR_ARG (-1)     /* bad argument / inconsistent configuration */
#define SQG_ERR_CUDA (-2)    /* a CUDA runtime call failed (see sqg_last_error) */
#define SQG_ERR_NOMEM (-3)   /* host or device allocation failed */
#define SQG_ERR_RANGE (-4)   /* a read would exceed UINT32_MAX samples (reference: src/sim.c:559-562) */
#define SQG_ERR_STATE (-5)   /* call not valid in this state (e.g. unknown ticket) */
#define SQG_ERR_NODEVICE (-6) /* no usable CUDA device: this library has NO CPU fallback */

/* Option flags: the reference's SQ_* bits (src/sq.h:33-43), same values, so core->opt.flag can be
 * passed through unchanged.  Only the bits below influence the hot path. */
#define SQG_RNA 0x001u        /* emit the signal 3'->5' (src/gensig.c:348-354) */
#define SQG_IDEAL 0x004u      /* no noise at all; offset/median_before = their means */
#define SQG_IDEAL_TIME 0x008u /* fixed dwell = (int)dwell_mean */
#define SQG_IDEAL_AMP 0x010u  /* level_mean without amplitude noise */
#define SQG_PREFIX 0x020u     /* adaptor/stall (DNA) or polyA/adaptor/stall (RNA), src/genread.c:71-123 */

/* Random-number schemes */
#define SQG_RNG_PHILOX 0 /* counter-based Philox4x32-7; every draw addressed by (read, k-mer | sample);
                            output independent of batching, threads and GPU count */
#define SQG_RNG_LEGACY 1 /* the reference's minstd streams (src/rand.h, src/sim.c:215-258) reproduced by
                            jump-ahead: bit-compatible with `squigulator -t1` for the same read order */

/* == profile_t (src/sq.h:47-58), field for field */
typedef struct {
    double digitisation;
    double sample_rate;
    double bps;
    double range;
    double offset_mean;
    double offset_std;
    double median_before_mean;
    double median_before_std;
    double dwell_mean;
    double dwell_std;
} sqg_profile_t;

/* == model_t (src/sq.h:61-68): one (level_mean, level_stdv) per k-mer rank */
typedef struct {
    float level_mean;
    float level_stdv;
} sqg_model_t;

typedef struct {
    sqg_profile_t profile; /* core->profile */
    uint32_t flags;        /* core->opt.flag (SQG_* bits are read, others ignored) */
    uint32_t kmer_size;    /* core->kmer_size: 1..9 */
    uint32_t num_kmer;     /* core->num_kmer: 4^k, or 5^k with meth */
    int32_t meth;          /* non-zero <=> core->opt.meth_freq != NULL: base-5 ranks, table = core->cpgmodel */
    float amp_noise;       /* core->opt.amp_noise */
    int64_t seed;          /* core->opt.seed */
    int32_t rng_mode;      /* SQG_RNG_* */
    int32_t device;        /* CUDA device ordinal */
    int32_t n_slots;       /* batches in flight for sqg_submit (0 = default 3) */
    int32_t reserved;      /* 0.  (Testing knob, > 0: shrinks the signal kernel's window so that small inputs reach its
                              cut-run and slow-tile paths - low 16 bits: samples per tile, high bits: k-mers.) */
} sqg_config_t;

typedef struct sqg_ctx sqg_ctx_t;

/* ---- life cycle (replaces the model/RNG part of init_core/init_rand, src/sim.c:215-326) ---- */

/* model: num_kmer entries in host memory, rank order (core->model or core->cpgmodel). */
int sqg_init(sqg_ctx_t **ctx, const sqg_config_t *cfg, const sqg_model_t *model);
/* Same, but the table already sits in device memory of cfg->device (e.g. after an NCCL broadcast);
 * d_model is a CUDA device pointer to num_kmer sqg_model_t and is copied, not retained. */
int sqg_init_device_model(sqg_ctx_t **ctx, const sqg_config_t *cfg, const void *d_model);
void sqg_destroy(sqg_ctx_t *ctx);
/* message for the last failure on this context (or of sqg_init* when ctx == NULL) */
const char *sqg_last_error(const sqg_ctx_t *ctx);
const char *sqg_version(void);

/* pinned host memory for inputs (optional; pageable memory works, slower) */
void *sqg_host_alloc(size_t bytes);
void sqg_host_free(void *p);

/* ---- one batch of reads, host buffers in, host buffers out (replaces process_db's fan-out) ---- */

#define SQG_WANT_SS 0x1u  /* also return the per-k-mer dwell array (aln->ss, src/gensig.c:273-281) */
#define SQG_WANT_SVB 0x2u /* return every read's signal as slow5lib's svb-zd stream - exactly the bytes
                             slow5_ptr_compress_solo(SLOW5_COMPRESS_SVB_ZD, raw_signal, ...) gives
                             (slow5lib/src/slow5_press.c:1055-1087) - INSTEAD of raw int16: `signal` is NULL,
                             `svb`/`svb_off` are set; ~1.3 bytes per sample cross PCIe instead of 2 */
#define SQG_WANT_RECORDS 0x10u /* return every read as a finished BLOW5 RECORD - exactly the bytes slow5_rec_to_mem()
                                  (slow5lib/src/slow5.c:3815-4010) makes of it for a file opened with
                                  slow5_set_press(sp, SLOW5_COMPRESS_NONE, SLOW5_COMPRESS_SVB_ZD): record size, read id,
                                  primary fields, svb-zd signal, the auxiliary fields squigulator sets
                                  (src/gensig.c:130-217) - all records back to back in `svb` (svb_off / svb_len per
                                  record), so that the host's part of writing a batch is ONE slow5_write_bytes /
                                  fwrite.  Set by sqg_gen_batch_records / sqg_submit_records, which take the ids. */
#define SQG_WANT_SS_TEXT 0x4u /* return aln->ss already formatted as the PAF/SAM `ss:Z:` value: "d0,d1,...,dn-1," per read
                                 (no terminator), RNA reads last k-mer first, exactly what src/format.c:69-75 appends */

typedef struct {
    int64_t n_reads;
    int64_t total_samples;         /* sum of len_raw_signal */
    const int16_t *signal;         /* pinned host buffer; read i occupies signal[sig_off[i] .. +len_raw_signal[i]) */
    const int64_t *sig_off;        /* n_reads entries (each a multiple of 64 samples) */
    const int64_t *len_raw_signal; /* n_reads */
    const double *offset;          /* n_reads: per-read ADC offset   (src/gensig.c:316) */
    const double *median_before;   /* n_reads                         (src/gensig.c:317) */
    const int32_t *ss;             /* SQG_WANT_SS: dwell per k-mer, read i at ss[ss_off[i] .. ss_off[i+1]) */
    const int64_t *ss_off;         /* n_reads+1 */
    const uint8_t *svb;            /* SQG_WANT_SVB: read i's stream (SQG_WANT_RECORDS: its record) = svb[svb_off[i] .. + svb_len[i]) */
    const int64_t *svb_off;        /* n_reads+1 (streams / records back to back; [n_reads] = bytes copied) */
    const int64_t *svb_len;        /* n_reads */
    const char *ss_text;           /* SQG_WANT_SS_TEXT: read i's dwell string = ss_text[ss_text_off[i] .. ss_text_off[i+1]) */
    const int64_t *ss_text_off;    /* n_reads+1 */
    const char *bases;             /* SQG_WANT_BASES (coordinate batches): read i = bases[bases_off[i] .. bases_off[i+1]) */
    const int64_t *bases_off;      /* n_reads+1 */
    int64_t meth_draws;            /* coordinate batches with meth: rand_meth draws this batch consumed */
} sqg_result_t;

/* bases: the reads' characters back to back (no terminators needed); read i = bases[base_off[i] ..
 * base_off[i+1]).  first_read_index = global number of read 0 (core->total_reads): it is the Philox
 * counter, so a job split over batches/GPUs gives identical output.
 * Synchronous.  *res stays valid until the next sqg_gen_batch/sqg_destroy on this context. */
int sqg_gen_batch(sqg_ctx_t *ctx, int64_t n_reads, const char *bases, const int64_t *base_off,
                  int64_t first_read_index, uint32_t want, sqg_result_t *res);

/* ---- finished BLOW5 records (replaces set_record_*_fields + slow5_encode, src/sim.c:603-611) ----
 * What only the host knows about the records of a batch: the read ids (src/sim.c:564-570), the samples generated before
 * the batch (aux start_time continues from there in read order, src/sim.c:602) and whether --ont-friendly added the
 * end_reason field.  aux read_number of read i is first_read_index + i (src/sim.c:604).  The arrays must stay valid
 * until the batch is done (sqg_wait for submitted batches). */
typedef struct {
    const char *read_ids;   /* ids back to back, no terminators: read i = read_ids[id_off[i] .. id_off[i+1]) */
    const int64_t *id_off;  /* n_reads+1 */
    uint64_t start_time0;   /* core->n_samples before the batch */
    int32_t ont_friendly;   /* non-zero: records carry end_reason = 0 (src/gensig.c:211-217) */
    int32_t reserved;
} sqg_record_info_t;
int sqg_gen_batch_records(sqg_ctx_t *ctx, int64_t n_reads, const char *bases, const int64_t *base_off,
                          int64_t first_read_index, const sqg_record_info_t *rec, uint32_t want, sqg_result_t *res);

/* ---- asynchronous dispatcher: CUDA-stream slots instead of src/thread.c's pthread pool ----
 * sqg_submit returns as soon as the batch is queued on a free slot (it blocks only while all
 * n_slots are in flight); the inputs must stay valid until sqg_wait returns.  One submitter thread;
 * any thread may wait.  Results stay valid until sqg_release.  A ticket holds its slot until sqg_release - also
 * when sqg_wait returned an error (sqg_last_error then carries that job's own message).
 * SQG_RNG_LEGACY consumes the reference's streams in submission order: one slot, and sqg_gen_batch refuses to run
 * while a submitted batch is in flight. */
typedef int64_t sqg_ticket_t;
int sqg_submit(sqg_ctx_t *ctx, int64_t n_reads, const char *bases, const int64_t *base_off,
               int64_t first_read_index, uint32_t want, sqg_ticket_t *ticket);
int sqg_submit_records(sqg_ctx_t *ctx, int64_t n_reads, const char *bases, const int64_t *base_off,
                       int64_t first_read_index, const sqg_record_info_t *rec, uint32_t want, sqg_ticket_t *ticket);
int sqg_wait(sqg_ctx_t *ctx, sqg_ticket_t ticket, sqg_result_t *res);
int sqg_release(sqg_ctx_t *ctx, sqg_ticket_t ticket);

/* ---- device-resident genome: reads named by coordinates, extracted on the GPU ----
 * Replaces, for accepted reads, the sequence work of gen_read() (src/genread.c:357): the copy out of the reference
 * (gen_read_common, src/genread.c:149-153), the replacement of every 'N' by the read's own minstd stream seeded 100
 * (is_bad_read, src/genread.c:132-140), reverse_complement() for '-' reads (src/seq.h:77-112) and the CpG -> 'M'
 * marking of methylate_dna() (src/genread.c:207-241).  Coordinate SAMPLING and the accept/reject decision stay with
 * the caller (they consume the reference's ref_pos/strand/rlen streams in host order); so does the contig-end clip:
 * pos + len must lie inside the contig.
 *
 * seq: all contigs back to back, contig c = seq[contig_off[c] .. contig_off[c+1]) (ASCII, as load_ref() holds them:
 * case, N and IUPAC letters keep the reference's meaning).  meth: NULL, or one uint8 per base with the same indexing
 * (ref->ref_meth, src/ref.c:346); contig_has_meth: NULL (= all, when meth != NULL) or one flag per contig
 * (ref->ref_meth[c] != NULL).  Copied to HBM once; the host arrays are not retained.  Calling it again replaces the genome. */
int sqg_genome_load(sqg_ctx_t *ctx, int32_t n_contigs, const char *seq, const int64_t *contig_off,
                    const uint8_t *meth, const uint8_t *contig_has_meth);

typedef struct {
    int32_t contig; /* index into the loaded genome */
    int32_t len;    /* *rlen of gen_read: bases in the read, >= 0 */
    int64_t pos;    /* *ref_pos: 0-based start on the forward strand */
    int32_t strand; /* '+' or '-' (src/genread.c:196-200) */
    int32_t reserved;
} sqg_coord_t;

#define SQG_WANT_BASES 0x8u /* coordinate batches: also return the extracted reads (for FASTA/FASTQ/SAM records) */

/* As sqg_gen_batch / sqg_submit, the reads given as coordinates.  meth_draw_base = number of values already taken
 * from the reference's rand_meth stream (seed + 6, src/sim.c:230-246; thread 0) before this batch: with cfg->meth the
 * j-th CpG site of the batch (read order, then position order) uses draw meth_draw_base + j, exactly the value
 * `squigulator -t1` would use; res->meth_draws tells the caller how far the batch advanced the stream. */
int sqg_gen_batch_coords(sqg_ctx_t *ctx, int64_t n_reads, const sqg_coord_t *coords, int64_t first_read_index,
                         int64_t meth_draw_base, uint32_t want, sqg_result_t *res);
int sqg_submit_coords(sqg_ctx_t *ctx, int64_t n_reads, const sqg_coord_t *coords, int64_t first_read_index,
                      int64_t meth_draw_base, uint32_t want, sqg_ticket_t *ticket);

/* ---- per-read drop-in with gen_sig's own shape (src/gensig.c:346) ----
 * Returns a malloc()'d buffer the caller free()s (slow5lib free()s rec->raw_signal itself,
 * slow5lib/src/slow5.c:3982), NULL on error.  read_index replaces `tid` (streams are per read,
 * not per thread).  ss/ss_n (nullable) receive a malloc()'d copy of aln->ss. */
int16_t *sqg_gen_sig(sqg_ctx_t *ctx, const char *read, int32_t len, double *offset, double *median_before,
                     int64_t *len_raw_signal, int64_t read_index, int32_t **ss, int64_t *ss_n);

/* ---- device-resident batches: what bench.py times as the kernel-only figure ---- */
typedef struct sqg_dev_batch sqg_dev_batch_t;
/* upload reads once; nothing is generated yet */
int sqg_dev_batch_create(sqg_ctx_t *ctx, int64_t n_reads, const char *bases, const int64_t *base_off,
                         int64_t first_read_index, uint32_t want, sqg_dev_batch_t **batch);
/* run the whole hot path (dwell pass, scans, signal kernel) `steps` times on the batch's stream, inputs and
 * outputs resident in HBM.  Timed with CUDA events on that stream: *ms_total covers all steps,
 * *ms_signal_kernel is the sum of the signal kernel's own launch durations.  Either may be NULL. */
int sqg_dev_batch_run(sqg_ctx_t *ctx, sqg_dev_batch_t *batch, int32_t steps, float *ms_total,
                      float *ms_signal_kernel);
/* after a run: totals and optional copy-out (any pointer may be NULL) */
int sqg_dev_batch_info(sqg_ctx_t *ctx, sqg_dev_batch_t *batch, int64_t *total_samples, int64_t *total_kmers,
                       int64_t *total_bases, int64_t *kernel_launches);
int sqg_dev_batch_fetch(sqg_ctx_t *ctx, sqg_dev_batch_t *batch, sqg_result_t *res);
void sqg_dev_batch_destroy(sqg_ctx_t *ctx, sqg_dev_batch_t *batch);

/* store-only kernel over `bytes` of HBM (the write ceiling the signal kernel is compared with) */
int sqg_bench_store(sqg_ctx_t *ctx, size_t bytes, int32_t steps, float *ms_total);

/* number of kernels launched by this context so far (bench.py reports it as gpu_launches) */
int64_t sqg_launch_count(const sqg_ctx_t *ctx);

#ifdef __cplusplus
}
#endif
#endif /* SQG_H */
