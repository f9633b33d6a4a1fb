/* oracle/sqg_oracle.h — TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C, single-threaded restatement of the reference hot path
 * (/root/reference src/gensig.c:226-356, src/seq.h:14-74, src/rand.h:79-94, src/sim.c:215-258,
 *  src/genread.c:37-281).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this; nothing under squigulator_b200/ links or imports it.
 *
 * Two random-number modes:
 *   SQO_RNG_LEGACY  the reference's own minstd/Lehmer streams + libm Box-Muller, stream-for-stream
 *                   (pinned bit-exactly against the reference's golden .exp files and against
 *                   oracle/_ref/libsqref.so, see tests/test_oracle_golden.py)
 *   SQO_RNG_PHILOX  the counter-based scheme the CUDA path implements:
 *                   Philox4x32-7 keyed by the seed, counters (block, read_lo, read_hi, stream), eight
 *                   10-bit draws per block -> standard normal through the class-stratified quantile
 *                   table (squigulator_b200/data/ztable_v3.bin), samples by fma.rz in binary32.
 *                   The GPU must equal this bit for bit.
 * Also restated here, each pinned to the compiled reference: slow5lib's svb-zd stream, the PAF/SAM
 * ss:Z: text (src/format.c:69-75) and the sequence work of gen_read() (src/genread.c, src/seq.h).
 */
#ifndef SQG_ORACLE_H
#define SQG_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* same bit values as the reference's SQ_* flags (src/sq.h:33-43) */
#define SQO_RNA 0x001
#define SQO_IDEAL 0x004
#define SQO_IDEAL_TIME 0x008
#define SQO_IDEAL_AMP 0x010
#define SQO_PREFIX 0x020

#define SQO_RNG_PHILOX 0
#define SQO_RNG_LEGACY 1

/* == profile_t, src/sq.h:47-58 */
typedef struct {
    double digitisation, sample_rate, bps, range;
    double offset_mean, offset_std, median_before_mean, median_before_std;
    double dwell_mean, dwell_std;
} sqo_profile_t;

typedef struct {
    sqo_profile_t profile;
    uint32_t flags;     /* SQO_* */
    uint32_t kmer_size; /* k */
    uint32_t num_kmer;  /* 4^k, or 5^k when meth != 0 */
    int32_t meth;       /* non-zero: base-5 {A,C,G,M,T} ranks (reference: opt.meth_freq != NULL) */
    float amp_noise;    /* opt.amp_noise, default 1 */
    int64_t seed;
    int32_t rng_mode;   /* SQO_RNG_* */
    int32_t num_thread; /* legacy mode: number of independent stream sets (reference -t) */
} sqo_config_t;

/* model: num_kmer interleaved (level_mean, level_stdv) floats == model_t[] (src/sq.h:61-68).
 * ztable: the bytes of squigulator_b200/data/ztable_v3.bin (Z32[32768] binary32 ++ Z2[8192] binary32);
 * NULL allowed in legacy mode. */
void *sqo_open(const sqo_config_t *cfg, const float *model, const void *ztable);
void sqo_close(void *h);

/* One read.  read_index is the global read number (Philox counter words 1-2; ignored in legacy
 * mode, where order of calls is what matters, as in the reference).  tid selects the legacy
 * stream set.  *sig (and *ss when ss != NULL) are malloc'd; release with sqo_free_buf.
 * Returns len_raw_signal. */
int64_t sqo_gen_sig(void *h, const char *read, int32_t len, int64_t read_index, int tid, double *offset,
                    double *median_before, int16_t **sig, int32_t **ss, int64_t *ss_n);
void sqo_free_buf(void *p);

/* building blocks, exported for known-answer tests */
void sqo_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
uint32_t sqo_kmer_rank(const char *s, uint32_t k);      /* src/seq.h:31-42 */
uint32_t sqo_meth_kmer_rank(const char *s, uint32_t k); /* src/seq.h:62-74 */
double sqo_lehmer_next(int64_t *state);                 /* src/rand.h:79-85 */
double sqo_lehmer_normal(int64_t *state, double m, double s); /* src/rand.h:87-94 */
void sqo_philox4x32(const uint32_t ctr[4], const uint32_t key[2], int rounds, uint32_t out[4]);
float sqo_z32(const void *ztable, uint32_t idx, const uint32_t key[2], uint32_t c0, uint32_t c1, uint32_t c2,
              uint32_t tail_stream);
float sqo_fma_rz(float x, float y, float z);            /* PTX fma.rz.f32, restated */
int64_t sqo_ss_text(const int32_t *ss, int64_t n, int rna, char *out);     /* src/format.c:69-75 */
int64_t sqo_svb_zd_encode(const int16_t *sig, int64_t n, uint8_t *out); /* slow5lib/src/slow5_press.c:1055-1087 */

/* the sequence work gen_read() does for an ACCEPTED read (src/genread.c:149-153 copy, :132-140 N replacement,
 * src/seq.h:77-112 reverse complement, src/genread.c:207-241 CpG marking).  contig/contig_meth: one contig
 * (contig_meth NULL = no methylation data); *meth_state: the rand_meth stream (src/sim.c:253), advanced as the
 * reference would (NULL = no marking).  out: len bytes.  Returns the number of rand_meth draws taken. */
int64_t sqo_extract_read(const char *contig, int64_t contig_len, const uint8_t *contig_meth, int64_t pos, int32_t len,
                         char strand, int64_t *meth_state, char *out);
/* state of a minstd stream after n draws from `seed` (seed*16807^n mod 2^31-1), as a value sqo_lehmer_next continues
 * from; holds for 0 <= seed <= 2^31-1 (a larger seed leaves the modular form for its first few draws) */
int64_t sqo_lehmer_jump(int64_t seed, uint64_t n);

#ifdef __cplusplus
}
#endif
#endif
