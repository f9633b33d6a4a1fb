/* oracle/sqg_oracle.c — TEST INFRASTRUCTURE ONLY (see sqg_oracle.h).
 *
 * CPU restatement of the squigulator signal-generation hot path.  Every function cites the
 * reference file:line it follows (paths relative to /root/reference).  Nothing here is shipped:
 * the product (squigulator_b200/csrc) never links, imports or executes this file.
 *
 * Parity status: PINNED.  Legacy mode reproduces the reference's golden .exp files bit for bit
 * (tests/test_oracle_golden.py, fixtures in tests/golden/ made by scripts/make_golden.py from the
 * reference built unmodified into oracle/_ref/), and equals oracle/_ref/libsqref.so on random
 * inputs (tests/test_oracle_vs_ref.py).  Philox mode shares every non-RNG line with legacy mode.
 *
 * Compile with -ffp-contract=off (oracle/Makefile): no expression below may be fused.
 */
#include "sqg_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ k-mer ranks */

/* src/seq.h:14-28: IUPAC folding; anything unknown (incl. N and NUL) ranks as A */
static uint32_t base_rank4(char b) {
    switch (b) {
        case 'A': case 'a': case 'R': case 'W': case 'M': case 'D': case 'H': case 'V': return 0;
        case 'C': case 'c': case 'Y': case 'B': return 1;
        case 'G': case 'g': case 'S': case 'K': return 2;
        case 'T': case 't': case 'U': return 3;
        default: return 0;
    }
}

/* src/seq.h:45-60: upper-case A,C,G,M,T only */
static uint32_t base_rank5(char b) {
    switch (b) {
        case 'A': return 0;
        case 'C': return 1;
        case 'G': return 2;
        case 'M': return 3;
        case 'T': return 4;
        default: return 0;
    }
}

/* src/seq.h:31-42: first base is the most significant 2-bit digit */
uint32_t sqo_kmer_rank(const char *s, uint32_t k) {
    uint32_t r = 0;
    for (uint32_t i = 0; i < k; i++) r = (r << 2) | base_rank4(s[i]);
    return r;
}

/* src/seq.h:62-74: first base is the most significant base-5 digit */
uint32_t sqo_meth_kmer_rank(const char *s, uint32_t k) {
    uint32_t r = 0;
    for (uint32_t i = 0; i < k; i++) r = r * 5 + base_rank5(s[i]);
    return r;
}

/* ------------------------------------------------------------------ legacy RNG (src/rand.h) */

/* src/rand.h:79-85: Schrage minstd step; the UN-normalised value is what is stored back */
double sqo_lehmer_next(int64_t *state) {
    int64_t x = *state;
    int64_t hi = x / 127773, lo = x % 127773;
    int64_t nx = 16807 * lo - 2836 * hi;
    *state = nx;
    if (nx <= 0) nx += 2147483647;
    return (double)nx / 2147483647;
}

/* src/rand.h:87-94: Box-Muller, cosine branch, truncated pi literal, exactly two uniforms */
double sqo_lehmer_normal(int64_t *state, double m, double s) {
    double u = sqo_lehmer_next(state);
    double t = 2.0 * 3.14159265 * sqo_lehmer_next(state);
    double x = sqrt(-2.0 * log(u)) * cos(t);
    return x * s + m;
}

/* ------------------------------------------------------------------ Philox4x32-R */

/* Salmon, Moraes, Dror, Shaw 2011.  `rounds` = 10 is the standard generator (Random123 known answers);
 * the signal path uses SQO_PHILOX_ROUNDS = 7, the smallest round count the authors report as
 * Crush-resistant, as the product does (squigulator_b200/csrc/sqg_device.cuh: PHILOX_ROUNDS). */
void sqo_philox4x32(const uint32_t ctr[4], const uint32_t key[2], int rounds, uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < rounds; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

void sqo_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    sqo_philox4x32(ctr, key, 10, out);
}

#define SQO_PHILOX_ROUNDS 7
static void philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    sqo_philox4x32(ctr, key, SQO_PHILOX_ROUNDS, out);
}

/* counter word 3 = stream tag */
enum { ST_AMP = 0, ST_DWELL = 1, ST_READ = 2, ST_AMP_TAIL = 3, ST_DWELL_TAIL = 4, ST_READ_TAIL = 5 };

#define Z32_N 32768
#define Z_TAIL_IDX 0x3FE0u /* (m = 511) << 5 | (class 0): the outermost half-normal cell */
#define Z2_SUB 8192

/* Table normal (DESIGN.md "Z32").  idx = (r << 5) | c: r = ten random bits (bit 9 = sign, bits 0-8 = slot m),
 * c = the CLASS = low five bits of the draw's Philox block number.  The entry is a 4-byte float whose word
 * index has c in its low five bits = the shared-memory bank, so the 32 lanes of a warp (32 consecutive
 * blocks) never collide; a class owns 512 half-normal cells spread evenly over the quantile range and is
 * scaled to unit variance (scripts/make_ztable.py).  The outermost cell (class 0, m = 511) takes 13 more
 * bits from a dedicated Philox block addressed by (c0,c1,c2,tail_stream).
 * zt = the bytes of squigulator_b200/data/ztable_v3.bin: Z32[32768] binary32, then Z2[8192] binary32. */
float sqo_z32(const void *zt, uint32_t idx, const uint32_t key[2], uint32_t c0, uint32_t c1, uint32_t c2,
              uint32_t tail_stream) {
    const float *z32 = (const float *)zt;
    const float *z2 = z32 + Z32_N;
    idx &= Z32_N - 1;
    if ((idx & 0x3FFFu) == Z_TAIL_IDX) {
        uint32_t ctr[4] = {c0, c1, c2, tail_stream}, w[4];
        philox(ctr, key, w);
        float z = z2[w[0] & (Z2_SUB - 1)];
        return (idx & 0x4000u) ? -z : z;
    }
    return z32[idx];
}

/* the j-th 10-bit draw (j in 0..7) of a Philox block: even draws are bits 7..16 of word j/2, odd draws bits
 * 7..16 of the same word rotated by 16 (i.e. bits 23..31 and 0).  On the GPU `word & 0x1FF80` is directly
 * the byte offset of the class-0 table entry. */
static uint32_t draw10(const uint32_t w[4], uint32_t j) {
    uint32_t x = w[j >> 1];
    if (j & 1) x = (x >> 16) | (x << 16);
    return (x >> 7) & 0x3FFu;
}
static uint32_t zindex(const uint32_t w[4], uint32_t j, uint32_t block) { return (draw10(w, j) << 5) | (block & 31u); }

/* ------------------------------------------------------------------ handle */

typedef struct {
    int64_t dwell, offset, median; /* Lehmer states: seeds S+2, S+4, S+5 (src/sim.c:238-242) */
    int64_t *kmer;                 /* one state per k-mer rank, seed S+j (src/sim.c:247-250) */
} legacy_streams_t;

typedef struct {
    sqo_config_t cfg;
    float *mean, *sd_eff; /* level_mean; level_stdv*amp_noise as a FLOAT product (src/sim.c:249) */
    const void *zt;
    uint32_t key[2];
    legacy_streams_t *ls;
} oracle_t;

void *sqo_open(const sqo_config_t *cfg, const float *model, const void *ztable) {
    oracle_t *o = (oracle_t *)calloc(1, sizeof(oracle_t));
    o->cfg = *cfg;
    uint32_t n = cfg->num_kmer;
    o->mean = (float *)malloc(sizeof(float) * n);
    o->sd_eff = (float *)malloc(sizeof(float) * n);
    for (uint32_t j = 0; j < n; j++) {
        o->mean[j] = model[2 * j];
        o->sd_eff[j] = model[2 * j + 1] * cfg->amp_noise;
    }
    o->zt = ztable;
    o->key[0] = (uint32_t)((uint64_t)cfg->seed & 0xFFFFFFFFu);
    o->key[1] = (uint32_t)((uint64_t)cfg->seed >> 32);
    if (cfg->rng_mode == SQO_RNG_LEGACY) {
        int t = cfg->num_thread > 0 ? cfg->num_thread : 1;
        o->cfg.num_thread = t;
        o->ls = (legacy_streams_t *)calloc(t, sizeof(legacy_streams_t));
        int64_t s = cfg->seed; /* src/sim.c:236-257: thread i starts at seed + i*(num_kmer+10) */
        for (int i = 0; i < t; i++) {
            o->ls[i].dwell = s + 2;
            o->ls[i].offset = s + 4;
            o->ls[i].median = s + 5;
            o->ls[i].kmer = (int64_t *)malloc(sizeof(int64_t) * n);
            for (uint32_t j = 0; j < n; j++) o->ls[i].kmer[j] = s + j;
            s += (int64_t)n + 10;
        }
    } else if (!ztable) {
        free(o->mean); free(o->sd_eff); free(o);
        return NULL;
    }
    return o;
}

void sqo_close(void *hv) {
    oracle_t *o = (oracle_t *)hv;
    if (!o) return;
    if (o->ls) {
        for (int i = 0; i < o->cfg.num_thread; i++) free(o->ls[i].kmer);
        free(o->ls);
    }
    free(o->mean);
    free(o->sd_eff);
    free(o);
}

void sqo_free_buf(void *p) { free(p); }

/* ------------------------------------------------------------------ prefix constants */
/* src/genread.c:37-39 (polyA is 158 x 'A'), :88 (RNA stall), :113 (DNA stall) */
static const char ADAPTOR_DNA[] = "GGCGTCTGCTTGGGTGTTTAACCTTTTTTTTTTAATGTACTTCGTTCAGTTACGTATTGCT";
static const char ADAPTOR_RNA[] =
    "TGATGATGAGGGATAGACGATGGTTGTTTCTGTTGGTGCTGATATTGCTTTTTTTTTTTTTATGATGCAAGATACGCAC";
static const char STALL_DNA[] = "TTTTTTTTTTTTTTTTTTAATCAA";
static const char STALL_RNA[] = "AAAAAGAAAAAACCCCCCCCCCCCCCCCCC";
#define POLYA_LEN 158

/* ------------------------------------------------------------------ the path */

typedef struct {
    const char *s;
    int32_t len;
} seg_t;

/* double -> int16 exactly as the reference's `raw_signal[n] = <double>` store compiles on x86-64
 * (src/gensig.c:270): truncate toward zero to int32, keep the low 16 bits (no clamp) */
static int16_t to_i16_d(double v) {
    int32_t i;
    if (v >= 2147483648.0) i = INT32_MAX;       /* out-of-int32 inputs never occur for sane */
    else if (v < -2147483648.0) i = INT32_MIN;  /* profiles; saturate like cvt.rzi.s32.f64   */
    else i = (int32_t)v;
    return (int16_t)(uint16_t)((uint32_t)i & 0xFFFFu);
}
static int16_t to_i16_f(float v) {
    int32_t i;
    if (v >= 2147483648.0f) i = INT32_MAX;
    else if (v < -2147483648.0f) i = INT32_MIN;
    else i = (int32_t)v;
    return (int16_t)(uint16_t)((uint32_t)i & 0xFFFFu);
}

/* single-precision fused multiply-add rounded toward zero (PTX fma.rz.f32).  The exact value of x*y + z is
 * formed in double with round-to-odd (x*y is exact in double; the addition is done toward zero and made
 * sticky), then rounded toward zero to float: 53 bits with a sticky last bit round correctly to 24. */
#include <fenv.h>
static float fma_rz(float x, float y, float z);
float sqo_fma_rz(float x, float y, float z) { return fma_rz(x, y, z); }
static float fma_rz(float x, float y, float z) {
    const int mode = fegetround();
    fesetround(FE_TOWARDZERO);
    volatile double p = (double)x * (double)y; /* exact: 24 x 24 bits */
    volatile double zz = (double)z;
    feclearexcept(FE_INEXACT);
    volatile double s = p + zz;                /* toward zero */
    if (fetestexcept(FE_INEXACT)) {            /* sticky bit: round to odd */
        union { double d; uint64_t u; } c;
        c.d = s;
        c.u |= 1u;
        s = c.d;
    }
    volatile float r = (float)s;               /* toward zero again */
    fesetround(mode);
    return r;
}

int64_t sqo_gen_sig(void *hv, const char *read, int32_t len, int64_t read_index, int tid, double *offset_out,
                    double *median_out, int16_t **sig_out, int32_t **ss_out, int64_t *ss_n_out) {
    oracle_t *o = (oracle_t *)hv;
    const sqo_config_t *c = &o->cfg;
    const sqo_profile_t *p = &c->profile;
    const uint32_t k = c->kmer_size;
    const int legacy = c->rng_mode == SQO_RNG_LEGACY;
    const int ideal = (c->flags & SQO_IDEAL) != 0;
    const int fixed_dwell = ideal || (c->flags & SQO_IDEAL_TIME);
    const int fixed_amp = ideal || (c->flags & SQO_IDEAL_AMP);
    const int rna = (c->flags & SQO_RNA) != 0;
    const int prefix = (c->flags & SQO_PREFIX) != 0;
    legacy_streams_t *ls = legacy ? &o->ls[tid] : NULL;
    const uint32_t r_lo = (uint32_t)((uint64_t)read_index & 0xFFFFFFFFu);
    const uint32_t r_hi = (uint32_t)((uint64_t)read_index >> 32);

    /* --- per-read ADC offset and median_before: src/gensig.c:312-318 --- */
    double offset, median;
    if (ideal) {
        offset = p->offset_mean;
        median = p->median_before_mean;
    } else if (legacy) {
        offset = sqo_lehmer_normal(&ls->offset, p->offset_mean, p->offset_std);
        median = sqo_lehmer_normal(&ls->median, p->median_before_mean, p->median_before_std);
    } else {
        /* one Philox block per read: draws 0-3 make the offset deviate, draws 4-7 the median_before deviate, each
         * a unit-norm mix of four table normals (weights cos/sin products of 35, 40, 55 degrees) so that per-read
         * values have ~2^40 atoms; the draws' classes walk with the read index so that reads use all 32 */
        static const double W4[4] = {0.6275068715971331, 0.43938504177070503, 0.3686878264946124, 0.5265407845183632};
        uint32_t ctr[4] = {0, r_lo, r_hi, ST_READ}, w[4];
        philox(ctr, o->key, w);
        double z[2];
        for (int d = 0; d < 2; d++) {
            double t[4];
            for (uint32_t u = 0; u < 4; u++) {
                uint32_t j = 4 * (uint32_t)d + u;
                float zz = sqo_z32(o->zt, zindex(w, j, 8 * r_lo + j), o->key, j, r_lo, r_hi, ST_READ_TAIL);
                t[u] = (double)zz * W4[u];
            }
            double ab = t[0] + t[1], cd = t[2] + t[3];
            z[d] = ab + cd;
        }
        double t0 = z[0] * p->offset_std;
        offset = t0 + p->offset_mean;
        double t1 = z[1] * p->median_before_std;
        median = t1 + p->median_before_mean;
    }
    *offset_out = offset;
    *median_out = median;

    /* --- sequence segments: src/genread.c:95-123 (attach_prefix) + :88-89 (RNA stall) --- */
    char *joined = NULL;
    seg_t seg[2];
    int nseg = 1;
    if (prefix) {
        if (rna) {
            int32_t al = (int32_t)strlen(ADAPTOR_RNA);
            joined = (char *)malloc((size_t)len + POLYA_LEN + al + 1);
            memcpy(joined, read, (size_t)len);
            memset(joined + len, 'A', POLYA_LEN);
            memcpy(joined + len + POLYA_LEN, ADAPTOR_RNA, (size_t)al);
            len += POLYA_LEN + al;
            seg[1].s = STALL_RNA;
            seg[1].len = (int32_t)strlen(STALL_RNA);
            nseg = 2;
        } else {
            int32_t sl = (int32_t)strlen(STALL_DNA), al = (int32_t)strlen(ADAPTOR_DNA);
            joined = (char *)malloc((size_t)len + sl + al + 1);
            memcpy(joined, STALL_DNA, (size_t)sl);
            memcpy(joined + sl, ADAPTOR_DNA, (size_t)al);
            memcpy(joined + sl + al, read, (size_t)len);
            len += sl + al;
        }
        joined[len] = '\0';
        read = joined;
    }
    seg[0].s = read;
    seg[0].len = len;

    /* --- k-mer list (rank per k-mer), src/gensig.c:240-253 --- */
    int64_t nk_seg[2] = {0, 0}, nk = 0;
    for (int g = 0; g < nseg; g++) {
        nk_seg[g] = seg[g].len < (int32_t)k ? 5 : (int64_t)seg[g].len - k + 1; /* :242-245 "a hack" */
        nk += nk_seg[g];
    }
    uint32_t *rank = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)nk);
    int32_t *sps = (int32_t *)malloc(sizeof(int32_t) * (size_t)nk);
    {
        int64_t q = 0;
        for (int g = 0; g < nseg; g++) {
            /* short read: 5 k-mers of "ACGTACGTACGT"; for k=9 the last window runs onto the
             * terminating NUL, which ranks 0 like any unknown base */
            static const char HACK[16] = "ACGTACGTACGT\0\0\0";
            const char *s = seg[g].len < (int32_t)k ? HACK : seg[g].s;
            for (int64_t i = 0; i < nk_seg[g]; i++, q++)
                rank[q] = c->meth ? sqo_meth_kmer_rank(s + i, k) : sqo_kmer_rank(s + i, k);
        }
    }

    /* --- dwell per k-mer, src/gensig.c:254-257 --- */
    int64_t total = 0;
    for (int64_t i = 0; i < nk; i++) {
        int d = (int)p->dwell_mean;
        if (!fixed_dwell) {
            if (legacy) {
                d = (int)round(sqo_lehmer_normal(&ls->dwell, p->dwell_mean, p->dwell_std));
            } else if (p->dwell_std == 0.0) {
                d = (int)round(p->dwell_mean); /* round(N(mean, 0)) is a constant (rna004 presets) */
            } else {
                /* draw index: position in the read's k-mer list; a 2nd segment (RNA stall) starts at
                 * the next multiple of 8 so that every block of 8 draws lies inside one segment */
                int64_t di = i < nk_seg[0] ? i : ((nk_seg[0] + 7) & ~(int64_t)7) + (i - nk_seg[0]);
                uint32_t ctr[4] = {(uint32_t)(di >> 3), r_lo, r_hi, ST_DWELL}, w[4];
                philox(ctr, o->key, w);
                float z = sqo_z32(o->zt, zindex(w, (uint32_t)(di & 7), (uint32_t)(di >> 3)), o->key, (uint32_t)di, r_lo,
                                  r_hi, ST_DWELL_TAIL);
                /* Philox mode: single-precision FMA, round to nearest (ties to even; the reference's
                 * round() differs only on exact .5 ties, which table normals do not produce) */
                d = (int)lrintf(fmaf(z, (float)p->dwell_std, (float)p->dwell_mean));
            }
            if (d < 1) d = -d + 1;
        }
        sps[i] = d;
        total += d;
    }
    /* NB legacy mode: the reference draws a k-mer's dwell immediately before that k-mer's samples,
     * but the dwell stream is a separate state from every per-k-mer stream (k-mer 2 shares its SEED
     * with it, src/sim.c:240,249, not its state), so drawing all dwells first is exact. */

    /* --- samples, src/gensig.c:259-272 --- */
    int16_t *raw = (int16_t *)malloc(sizeof(int16_t) * (size_t)(total > 0 ? total : 1));
    const double scale = p->digitisation / p->range; /* Philox mode only */
    int64_t n = 0;
    for (int64_t i = 0; i < nk; i++) {
        const uint32_t r = rank[i];
        const float mean = o->mean[r];
        if (fixed_amp) {
            int16_t v = to_i16_d((double)mean * p->digitisation / p->range - offset);
            for (int j = 0; j < sps[i]; j++) raw[n++] = v;
        } else if (legacy) {
            for (int j = 0; j < sps[i]; j++) {
                float s = (float)sqo_lehmer_normal(&ls->kmer[r], (double)mean, (double)o->sd_eff[r]);
                raw[n++] = to_i16_d((double)s * p->digitisation / p->range - offset);
            }
        } else {
            /* Philox mode: single precision.  A' = (stdv*amp_noise)*scale and M = mean*scale are rounded once per product
             * (they depend on the k-mer only: the GPU tabulates them), c_r = 32768 - (float)offset once per read, and
             * Bm = M + c_r once per k-mer - for every sane profile Bm lies in [32768, 65536), where floats are spaced 2^-8
             * apart, so Bq = Bm - 32768 is B' on the 2^-8 grid.  The sample is the single-precision FMA z*A' + Bq ROUNDED
             * TOWARD ZERO, truncated toward zero like the reference's store (src/gensig.c:270) and wrapped to 16 bits.  The
             * GPU gets the same integer for 0 <= value < 32768 from the mantissa of fma.rz(z, A', Bm) without a conversion,
             * and falls back to this very expression otherwise. */
            const float scale_f = (float)scale, off_f = (float)offset;
            const float A = o->sd_eff[r] * scale_f;
            volatile float M = mean * scale_f;
            volatile float c_r = 32768.0f - off_f;
            volatile float Bm = M + c_r;
            const float Bq = Bm - 32768.0f;
            for (int j = 0; j < sps[i]; j++, n++) {
                /* Philox draws are addressed by the position q in the EMITTED signal (after the RNA reversal of
                 * src/gensig.c:348-354).  A block holds twelve 10-bit draws - three fields per word: bits 7..16 (F0), bits
                 * 17..26 (F1), bits 27..31,0..4 (F2) - and the chunks of 8 samples are taken in units of 96: chunk
                 * Cq = 96 u + 32 g + l uses blocks A = 64 u + 2 l and B = A + 1:
                 *   g = 0: sample e -> word e/2 of A, F0 (e even) / F1 (e odd);   g = 2: the same of B;
                 *   g = 1: F2 of word e of A (e < 4) or of word e - 4 of B.
                 * The draw's table class is l XOR five hash bits of (unit, group, read): a bijection of the 32 chunks of a
                 * group, different from group to group, so no sample position is tied to one class. */
                uint32_t q = (uint32_t)(rna ? total - 1 - n : n);
                uint32_t Cq = q >> 3, e = q & 7;
                uint32_t u = Cq / 96, r = Cq % 96, g = r >> 5, l = r & 31;
                uint32_t blk = 64 * u + 2 * l + (g == 2 || (g == 1 && e >= 4) ? 1 : 0);
                uint32_t ctr[4] = {blk, r_lo, r_hi, ST_AMP}, w[4];
                philox(ctr, o->key, w);
                uint32_t d10;
                if (g == 1) {
                    uint32_t x = w[e & 3];
                    d10 = ((x >> 27) | (x << 5)) & 0x3FFu; /* bits 27..31 (low part of the draw), 0..4 (high part) */
                } else {
                    uint32_t x = w[e >> 1];
                    d10 = (e & 1) ? (x >> 17) & 0x3FFu : (x >> 7) & 0x3FFu;
                }
                uint32_t h = u * 0x9E3779B1u + r_lo * 0x85EBCA6Bu; /* one xorshift-multiply round per unit and read ... */
                h ^= h >> 15;
                h = ((h * 0x2C1B3C6Du) >> (27 - 5 * g)) & 31u;           /* ... group g takes bits 27-5g .. 31-5g */
                float z = sqo_z32(o->zt, (d10 << 5) | (l ^ h), o->key, q, r_lo, r_hi, ST_AMP_TAIL);
                raw[n] = to_i16_f(fma_rz(z, A, Bq));
            }
        }
    }

    /* --- RNA adaptor level shift, src/genread.c:80-86 (stall k-mers were generated above as
     * segment 1, which is what the 2nd gen_sig_core_seq call at :89 appends) --- */
    if (prefix && rna) {
        int64_t n0 = 0;
        for (int64_t i = 0; i < nk_seg[0]; i++) n0 += sps[i];
        int64_t st = n0 - (int64_t)strlen(ADAPTOR_RNA) * (int)p->dwell_mean;
        if (st < 0) st = 0; /* the reference would write out of bounds here */
        int16_t off = (int16_t)(30 * p->digitisation / p->range);
        for (int64_t i = st; i < n0; i++) raw[i] = (int16_t)(raw[i] - off);
    }

    /* --- RNA: 3'->5' emission, src/gensig.c:348-354 --- */
    if (rna) {
        for (int64_t i = 0; i < total / 2; i++) {
            int16_t t = raw[i];
            raw[i] = raw[total - 1 - i];
            raw[total - 1 - i] = t;
        }
    }

    *sig_out = raw;
    if (ss_out) {
        *ss_out = sps; /* aln->ss, src/gensig.c:273-281 */
        *ss_n_out = nk;
    } else {
        free(sps);
    }
    free(rank);
    free(joined);
    return total;
}

/* ------------------------------------------------------------------ svb-zd */

/* slow5lib's signal compression for BLOW5 records, restated: ptr_compress_svb_zd
 * (slow5lib/src/slow5_press.c:1055-1087) = zig-zag delta with prev = 0
 * (thirdparty/streamvbyte/src/streamvbyte_zigzag.c:15-27) then ptr_compress_svb (:1029-1052): a uint32
 * sample count followed by the StreamVByte stream - ceil(n/4) key bytes holding 2 bits per value
 * (bytes - 1, first value in the low bits), then the values little-endian in 1-4 bytes each
 * (streamvbyte_encode.c:31-79).  out needs 4 + (n+3)/4 + 4n bytes.  Returns the bytes written. */
int64_t sqo_svb_zd_encode(const int16_t *sig, int64_t n, uint8_t *out) {
    uint32_t length = (uint32_t)n;
    memcpy(out, &length, 4);
    uint8_t *key = out + 4;
    uint8_t *data = key + (n + 3) / 4;
    int32_t prev = 0;
    for (int64_t i = 0; i < n; i++) {
        int32_t d = (int32_t)sig[i] - prev;
        prev = sig[i];
        uint32_t v = ((uint32_t)d + (uint32_t)d) ^ (uint32_t)(d >> 31);
        uint32_t code = v < (1u << 8) ? 0 : v < (1u << 16) ? 1 : v < (1u << 24) ? 2 : 3;
        if ((i & 3) == 0) key[i >> 2] = 0;
        key[i >> 2] |= (uint8_t)(code << (2 * (i & 3)));
        for (uint32_t b = 0; b <= code; b++) *data++ = (uint8_t)(v >> (8 * b));
    }
    return (int64_t)(data - out);
}

/* ------------------------------------------------------------------ ss:Z: text */

/* The dwell string of a PAF/SAM record, src/format.c:69-75 (paf_str) and :114-118 (sam_str): "%d," per k-mer,
 * last k-mer first for RNA (t_st > t_end).  out needs 12*n bytes; no terminator is written.  Returns the length. */
#include <stdio.h>
int64_t sqo_ss_text(const int32_t *ss, int64_t n, int rna, char *out) {
    char *p = out;
    for (int64_t i = 0; i < n; i++) {
        int64_t idx = rna ? n - i - 1 : i;
        p += sprintf(p, "%d,", ss[idx]);
    }
    return (int64_t)(p - out);
}

/* ------------------------------------------------------------------ read extraction */

static uint64_t mulmod31(uint64_t a, uint64_t b) { return (a * b) % 2147483647ULL; }

int64_t sqo_lehmer_jump(int64_t seed, uint64_t n) {
    /* one Schrage step maps any state x to 16807*x mod m (up to the sign convention of the stored value), so n steps
     * give seed*16807^n mod m; 0 stands for m */
    uint64_t r = (uint64_t)(seed % 2147483647LL + 2147483647LL) % 2147483647ULL, a = 16807, e = n, pw = 1;
    while (e) {
        if (e & 1) pw = mulmod31(pw, a);
        a = mulmod31(a, a);
        e >>= 1;
    }
    r = mulmod31(r, pw);
    return (int64_t)(r ? r : (n ? 2147483647ULL : 0));
}

int64_t sqo_extract_read(const char *contig, int64_t contig_len, const uint8_t *contig_meth, int64_t pos, int32_t len,
                         char strand, int64_t *meth_state, char *out) {
    /* src/genread.c:149-153 (the caller has clipped len at the contig end) + :132-140 */
    int64_t nstate = 100;
    for (int32_t i = 0; i < len; i++) {
        char b = contig[pos + i];
        if (b == 'N') {
            int n = (int)round(sqo_lehmer_next(&nstate) * 3);
            b = n == 0 ? 'A' : n == 1 ? 'C' : n == 2 ? 'G' : 'T';
        }
        out[i] = b;
    }
    if (strand == '-') { /* src/seq.h:77-112 */
        for (int32_t i = 0, j = len - 1; i <= j; i++, j--) {
            char x = out[i], y = out[j], cx, cy;
            cx = (x == 'A' || x == 'a') ? 'T' : (x == 'C' || x == 'c') ? 'G' : (x == 'G' || x == 'g') ? 'C'
                 : (x == 'T' || x == 't') ? 'A' : 'T';
            cy = (y == 'A' || y == 'a') ? 'T' : (y == 'C' || y == 'c') ? 'G' : (y == 'G' || y == 'g') ? 'C'
                 : (y == 'T' || y == 't') ? 'A' : 'T';
            out[i] = cy;
            out[j] = cx;
        }
    }
    int64_t draws = 0;
    if (meth_state && contig_meth) { /* src/genread.c:207-241 */
        for (int32_t i = 0; i < len; i++) {
            if (pos + i + 1 < contig_len && i + 1 < len && contig[pos + i] == 'C' && contig[pos + i + 1] == 'G') {
                int methr = (int)(sqo_lehmer_next(meth_state) * 254);
                draws++;
                if (methr <= contig_meth[pos + i]) out[strand == '-' ? len - i - 2 : i] = 'M';
            }
        }
    }
    return draws;
}
