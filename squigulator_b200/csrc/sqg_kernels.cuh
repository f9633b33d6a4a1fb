// sqg_kernels.cuh — the CUDA kernels of the signal-generation path (sm_100a, hand-written).
//
// Work decomposition (DESIGN.md "Kernels"):
//   a read is 1-2 SEGMENTS (the 2nd only for the RNA stall of --prefix, src/genread.c:88-89);
//   a segment is cut into TILES of T consecutive k-mers; a tile is what one CTA turns into samples.
//
//   K1 dwell_sum_kernel    per tile: draw the T dwells (Philox), sum them              -> tile_sum
//   K2 read_plan_kernel    per read: exclusive scan of its tile sums, per-read draws    -> tile_base, siglen, offset, median_before
//   K3 read_offsets_kernel one CTA : exclusive scan of the 64-sample-aligned read lengths -> sigoff, totals
//   K4 signal_kernel       per tile: encode k-mers, look up (mean,stdv), re-draw dwells, block scan,
//                          then emit every int16 sample with 128-bit stores             -> signal  (the hot kernel)
//
// Reference statements: src/gensig.c:226-288 (gen_sig_core_seq), :293-343 (gen_sig_core), :346-356 (gen_sig).
#pragma once
#include "sqg_device.cuh"

namespace sqg {

struct SegDesc {
    int64_t off_a;  // piece a: bases[off_a .. off_a+len_a)
    int64_t off_b;  // piece b (prefix/suffix constant or the read): logical positions >= len_a
    int32_t len_a;
    int32_t nk;      // k-mers in this segment
    int32_t read;    // local read index
    int32_t k0;      // k-mers of this read before the segment (ss offset)
    int32_t k0_rng;  // same, rounded up to 8 (Philox dwell-block alignment)
    int32_t tile0;   // first tile of the segment
};

struct ReadDesc {
    int32_t seg0;
    int32_t nseg;
    int64_t ss_off;     // start of this read in ss[]
    int32_t shift_len;  // RNA --prefix: trailing samples of segment 0 to lower (src/genread.c:80-86)
    int32_t pad;
};

struct GenParams {
    // inputs
    const uint8_t *bases;
    const SegDesc *segs;
    const ReadDesc *reads;
    const float2 *model;  // (level_mean, level_stdv) by rank
    const float *z1;      // Z1[32768]
    const float *z2;      // Z2[16*1024]
    // plan (written by K1-K3, read by K4)
    int32_t *tile_seg;
    uint32_t *tile_sum;
    uint32_t *tile_base;
    uint32_t *read_siglen;
    uint32_t *read_n0;
    int64_t *read_sigoff;
    double *read_offset;
    double *read_median;
    int64_t *meta;  // [0] arena samples needed, [1] sum of siglen, [2] error flag
    // outputs
    int16_t *sig;
    int32_t *ss;
    // geometry
    int32_t n_reads, n_segs, n_tiles;
    int32_t T;         // k-mers per tile (multiple of 8)
    int32_t k;         // k-mer size
    uint32_t kmask;    // base 4: 4^k-1;  base 5: 5^(k-1)
    uint32_t num_kmer;
    int32_t model_in_smem;
    // profile (src/sq.h:47-58) and options
    double digitisation, range, scale;  // scale = digitisation/range
    double offset_mean, offset_std, median_mean, median_std;
    double dwell_mean, dwell_std;
    int32_t sps_fixed;  // (int)dwell_mean
    int32_t ideal;      // SQ_IDEAL: per-read draws replaced by the means
    float amp_noise;
    uint32_t key0, key1;
    int64_t first_read;
    int32_t want_ss;
    int32_t shift_val;  // (int16)(30*digitisation/range)
};

constexpr int K1_THREADS = 256;
constexpr int K4_THREADS = 512;
constexpr int MAX_T = 2048;
constexpr int MAP_CAP = 8192;  // 8-sample chunks per tile the chunk->k-mer map can hold

// ------------------------------------------------------------------------------------------------
// small helpers

__device__ __forceinline__ RngKey make_key(const GenParams &p, int read_local) {
    const uint64_t r = (uint64_t)(p.first_read + read_local);
    return RngKey{p.key0, p.key1, (uint32_t)r, (uint32_t)(r >> 32)};
}

// the 8 dwells of k-mer block `blk` (k-mers 8*blk .. 8*blk+7 of the read's draw index space)
__device__ __forceinline__ void draw_dwell8(const GenParams &p, const float *__restrict__ z1, RngKey key, uint32_t blk,
                                            int d[8]) {
    const uint4 w = philox4x32_10(blk, key.r_lo, key.r_hi, ST_DWELL, key.k0, key.k1);
    const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const uint32_t h = (j & 1) ? (ww[j >> 1] >> 16) : (ww[j >> 1] & 0xFFFFu);
        const float z = z16(z1, p.z2, h, blk * 8 + j, key, ST_DWELL_TAIL);
        d[j] = dwell_from_z(z, p.dwell_mean, p.dwell_std);
    }
}

__device__ __forceinline__ int find_seg(const SegDesc *__restrict__ segs, int n_segs, int tile) {
    int lo = 0, hi = n_segs - 1;  // largest s with segs[s].tile0 <= tile
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (__ldg(&segs[mid].tile0) <= tile) lo = mid; else hi = mid - 1;
    }
    return lo;
}

// ------------------------------------------------------------------------------------------------
// K1: per-tile sum of dwells (random-dwell modes only)
__global__ void __launch_bounds__(K1_THREADS) dwell_sum_kernel(const __grid_constant__ GenParams p) {
    __shared__ int s_seg;
    __shared__ uint32_t s_part[K1_THREADS / 32];
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        if (threadIdx.x == 0) s_seg = find_seg(p.segs, p.n_segs, tile);
        __syncthreads();
        const int si = s_seg;
        const SegDesc seg = p.segs[si];
        const int kstart = (tile - seg.tile0) * p.T;
        const int nk_tile = min(p.T, seg.nk - kstart);
        const RngKey key = make_key(p, seg.read);
        uint32_t sum = 0;
        for (int g = threadIdx.x; g * 8 < nk_tile; g += K1_THREADS) {
            int d[8];
            draw_dwell8(p, p.z1, key, (uint32_t)((seg.k0_rng + kstart) >> 3) + g, d);
#pragma unroll
            for (int j = 0; j < 8; j++)
                if (g * 8 + j < nk_tile) sum += (uint32_t)d[j];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = sum;
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t t = 0;
#pragma unroll
            for (int w = 0; w < K1_THREADS / 32; w++) t += s_part[w];
            p.tile_sum[tile] = t;
            p.tile_seg[tile] = si;
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// K2: per read — scan its tiles, draw offset / median_before (src/gensig.c:312-318)
template <bool RAND_DWELL>
__global__ void __launch_bounds__(256) read_plan_kernel(const __grid_constant__ GenParams p) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= p.n_reads) return;
    const ReadDesc rd = p.reads[r];
    uint64_t total = 0;
    uint32_t n0 = 0;
    for (int s = rd.seg0; s < rd.seg0 + rd.nseg; s++) {
        const SegDesc seg = p.segs[s];
        const int ntile = (seg.nk + p.T - 1) / p.T;
        for (int t = 0; t < ntile; t++) {
            const int tile = seg.tile0 + t;
            uint32_t sum;
            if (RAND_DWELL) {
                sum = p.tile_sum[tile];
            } else {
                sum = (uint32_t)min(p.T, seg.nk - t * p.T) * (uint32_t)p.sps_fixed;
                p.tile_sum[tile] = sum;
                p.tile_seg[tile] = s;
            }
            p.tile_base[tile] = (uint32_t)total;
            total += sum;
        }
        if (s == rd.seg0) n0 = (uint32_t)total;
    }
    if (total >= 0xFFFFFFFFull) {  // src/sim.c:559-562
        atomicExch((unsigned long long *)&p.meta[2], 1ull);
        total = 0;
    }
    p.read_siglen[r] = (uint32_t)total;
    p.read_n0[r] = n0;

    double off = p.offset_mean, med = p.median_mean;
    if (!p.ideal) {
        const RngKey key = make_key(p, r);
        const uint4 w = philox4x32_10(0u, key.r_lo, key.r_hi, ST_READ, key.k0, key.k1);
        const uint32_t ww[2] = {w.x, w.y};
        double z[2];
#pragma unroll
        for (int d = 0; d < 2; d++) {
            const float za = z16(p.z1, p.z2, ww[d] & 0xFFFFu, 2 * d, key, ST_READ_TAIL);
            const float zb = z16(p.z1, p.z2, ww[d] >> 16, 2 * d + 1, key, ST_READ_TAIL);
            z[d] = __dadd_rn(__dmul_rn((double)za, 0.8), __dmul_rn((double)zb, 0.6));
        }
        off = __dadd_rn(__dmul_rn(z[0], p.offset_std), p.offset_mean);
        med = __dadd_rn(__dmul_rn(z[1], p.median_std), p.median_mean);
    }
    p.read_offset[r] = off;
    p.read_median[r] = med;
}

// ------------------------------------------------------------------------------------------------
// K3: one CTA — exclusive scan of the 64-sample-aligned read lengths
__global__ void __launch_bounds__(1024) read_offsets_kernel(const __grid_constant__ GenParams p) {
    __shared__ uint64_t s_warp[32];
    __shared__ uint64_t s_total;
    const int tid = threadIdx.x;
    const int per = (p.n_reads + 1023) / 1024;
    const int lo = min(p.n_reads, tid * per), hi = min(p.n_reads, lo + per);
    uint64_t part = 0, raw = 0;
    for (int r = lo; r < hi; r++) {
        const uint64_t l = p.read_siglen[r];
        part += (l + 63) & ~63ull;
        raw += l;
    }
    uint64_t inc = part;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint64_t v = __shfl_up_sync(0xffffffffu, inc, o);
        if ((tid & 31) >= o) inc += v;
    }
    if ((tid & 31) == 31) s_warp[tid >> 5] = inc;
    __syncthreads();
    if (tid < 32) {
        uint64_t w = s_warp[tid], winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint64_t v = __shfl_up_sync(0xffffffffu, winc, o);
            if (tid >= o) winc += v;
        }
        s_warp[tid] = winc - w;
        if (tid == 31) s_total = winc;
    }
    __syncthreads();
    uint64_t base = s_warp[tid >> 5] + inc - part;
    for (int r = lo; r < hi; r++) {
        p.read_sigoff[r] = (int64_t)base;
        base += ((uint64_t)p.read_siglen[r] + 63) & ~63ull;
    }
    // sum of raw lengths: warp + atomics
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) raw += __shfl_xor_sync(0xffffffffu, raw, o);
    if (tid == 0) p.meta[1] = 0;
    __syncthreads();
    if ((tid & 31) == 0) atomicAdd((unsigned long long *)&p.meta[1], (unsigned long long)raw);
    if (tid == 0) p.meta[0] = (int64_t)s_total;
}

// ------------------------------------------------------------------------------------------------
// K4: the signal kernel.

struct __align__(16) TileSmem {
    uint32_t off[MAX_T + 8];      // tile-relative first sample of each k-mer; off[nk] = S
    float2 par[MAX_T];            // NOISY: (A', B') ; else (unused, int16 value bits)
    uint16_t map[MAP_CAP];        // chunk -> first k-mer (random dwell only)
    uint8_t digit[MAX_T + 16];    // base digits of the tile's window
    uint8_t lut[256];
    uint32_t warp_sum[K4_THREADS / 32];
    uint32_t S;
    alignas(8) unsigned long long mbar;
};

// TMA bulk copy global -> shared (1-D), completion on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_load_1d(void *smem_dst, const void *gsrc, uint32_t bytes, unsigned long long *mbar) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    const uint32_t b = (uint32_t)__cvta_generic_to_shared(mbar);
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(d),
                 "l"(gsrc), "r"(bytes), "r"(b)
                 : "memory");
}
__device__ __forceinline__ void mbar_init(unsigned long long *mbar, uint32_t count) {
    const uint32_t b = (uint32_t)__cvta_generic_to_shared(mbar);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(b), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *mbar, uint32_t bytes) {
    const uint32_t b = (uint32_t)__cvta_generic_to_shared(mbar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(b), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *mbar, uint32_t phase) {
    const uint32_t b = (uint32_t)__cvta_generic_to_shared(mbar);
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE_%=;\n"
        "bra WAIT_LOOP_%=;\n"
        "WAIT_DONE_%=:\n"
        "}\n" ::"r"(b),
        "r"(phase)
        : "memory");
}

// dynamic shared memory: [TileSmem][z1: 128 KB if USE_Z][model: num_kmer*8 B if model_in_smem]
template <bool NOISY, bool RAND_DWELL, bool METH, bool REV>
__global__ void __launch_bounds__(K4_THREADS, 1) signal_kernel(const __grid_constant__ GenParams p) {
    constexpr bool USE_Z = NOISY || RAND_DWELL;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    TileSmem &ts = *reinterpret_cast<TileSmem *>(smem_raw);
    float *z1s = reinterpret_cast<float *>(smem_raw + ((sizeof(TileSmem) + 127) & ~127u));
    float2 *models = reinterpret_cast<float2 *>(reinterpret_cast<unsigned char *>(z1s) + (USE_Z ? Z1_N * 4 : 0));
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;

    // ---- one-time staging: quantile table and (small) pore model by TMA bulk copies ----
    if (tid == 0) mbar_init(&ts.mbar, 1);
    for (int i = tid; i < 256; i += K4_THREADS) ts.lut[i] = base_code(i);
    __syncthreads();
    if (tid == 0) {
        uint32_t bytes = 0;
        if (USE_Z) bytes += Z1_N * 4;
        if (p.model_in_smem) bytes += p.num_kmer * 8;
        if (bytes) {
            mbar_expect_tx(&ts.mbar, bytes);
            if (USE_Z) {
                // <= 64 KB per bulk copy keeps each request modest
                tma_load_1d(z1s, p.z1, Z1_N * 2, &ts.mbar);
                tma_load_1d(z1s + Z1_N / 2, p.z1 + Z1_N / 2, Z1_N * 2, &ts.mbar);
            }
            if (p.model_in_smem) tma_load_1d(models, p.model, p.num_kmer * 8, &ts.mbar);
        }
    }
    if (USE_Z || p.model_in_smem) mbar_wait(&ts.mbar, 0);
    const float2 *__restrict__ model = p.model_in_smem ? models : p.model;

    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        const int si = p.tile_seg[tile];
        const SegDesc seg = p.segs[si];
        const int kstart = (tile - seg.tile0) * p.T;
        const int nk_tile = min(p.T, seg.nk - kstart);
        const RngKey key = make_key(p, seg.read);
        const uint32_t B = p.tile_base[tile];     // first logical sample of the tile within the read
        const uint32_t L = p.read_siglen[seg.read];
        const double offset = p.read_offset[seg.read];
        int16_t *__restrict__ out = p.sig + p.read_sigoff[seg.read];

        // ---- phase A0: digits of the tile's base window ----
        const int nb = nk_tile + p.k - 1;
        for (int i = tid; i < nb; i += K4_THREADS) {
            const int pos = kstart + i;
            const uint8_t c = pos < seg.len_a ? p.bases[seg.off_a + pos] : p.bases[seg.off_b + (pos - seg.len_a)];
            const uint8_t code = ts.lut[c];
            ts.digit[i] = METH ? (code >> 4) : (code & 3);
        }
        __syncthreads();

        // ---- phase A1: 8 k-mers per thread: rank, model lookup, dwell, parameters ----
        const int g = tid;  // group index; groups beyond the tile idle (T/8 <= 256 < K4_THREADS)
        const bool active = g * 8 < nk_tile;
        int d[8];
        uint32_t local = 0;
        float2 pr[8];
        if (active) {
            if (RAND_DWELL) {
                draw_dwell8(p, z1s, key, (uint32_t)((seg.k0_rng + kstart) >> 3) + g, d);
            } else {
#pragma unroll
                for (int j = 0; j < 8; j++) d[j] = p.sps_fixed;
            }
            uint32_t rank = 0;
            const uint8_t *dg = ts.digit + g * 8;
            for (int i = 0; i < p.k - 1; i++) rank = METH ? rank * 5 + dg[i] : (rank << 2) | dg[i];
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const bool valid = g * 8 + j < nk_tile;
                if (!valid) d[j] = 0;
                const uint32_t dn = valid ? dg[p.k - 1 + j] : 0;
                // src/seq.h:31-42 / :62-74, rolling: drop the leading digit, append the new one
                rank = METH ? (rank % p.kmask) * 5 + dn : ((rank << 2) | dn) & p.kmask;
                const float2 mv = model[valid ? rank : 0];
                if (NOISY) {
                    const float sd = __fmul_rn(mv.y, p.amp_noise);  // float product, src/sim.c:249
                    const double a = __dmul_rn((double)sd, p.scale);
                    const double b = __dsub_rn(__dmul_rn((double)mv.x, p.scale), offset);
                    pr[j] = make_float2((float)a, (float)b);
                } else {
                    // src/gensig.c:266,270: (double)level_mean*digitisation/range - offset, truncated
                    const double v = __dsub_rn(__ddiv_rn(__dmul_rn((double)mv.x, p.digitisation), p.range), offset);
                    pr[j] = make_float2(0.f, __uint_as_float(to_i16_bits(v)));
                }
                local += (uint32_t)d[j];
            }
        }
        // block-wide exclusive scan of `local`
        uint32_t inc = local;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += v;
        }
        if (lane == 31) ts.warp_sum[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            const uint32_t w = lane < K4_THREADS / 32 ? ts.warp_sum[lane] : 0;
            uint32_t winc = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t v = __shfl_up_sync(0xffffffffu, winc, o);
                if (lane >= o) winc += v;
            }
            if (lane < K4_THREADS / 32) ts.warp_sum[lane] = winc - w;
            if (lane == 31) ts.S = winc;
        }
        __syncthreads();
        const uint32_t S = ts.S;
        // phase of the 8-sample chunks relative to the tile: chunk w covers tile samples [8w-ph, 8w-ph+8)
        const uint32_t ph = REV ? ((B - L) & 7u) : (B & 7u);
        if (active) {
            uint32_t o = ts.warp_sum[warp] + inc - local;
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const int m = g * 8 + j;
                if (m < nk_tile) {
                    ts.off[m] = o;
                    ts.par[m] = pr[j];
                    if (RAND_DWELL) {
                        // chunks whose first (clipped) sample falls inside this k-mer
                        uint32_t w0 = m == 0 ? 0u : (o + ph + 7) >> 3;
                        const uint32_t w1 = (o + (uint32_t)d[j] + ph + 7) >> 3;
                        for (; w0 < w1; w0++) ts.map[w0] = (uint16_t)m;
                    }
                    if (p.want_ss) p.ss[p.reads[seg.read].ss_off + seg.k0 + kstart + m] = d[j];
                    o += (uint32_t)d[j];
                }
            }
            if ((g + 1) * 8 >= nk_tile) ts.off[nk_tile] = S;
        }
        __syncthreads();

        // ---- phase B: samples.  One thread = one 16-byte chunk of the emitted signal. ----
        const uint32_t nW = (S + ph + 7) >> 3;
        for (uint32_t w = tid; w < nW; w += K4_THREADS) {
            const int s0 = (int)(8 * w) - (int)ph;  // first tile sample of the chunk (may be < 0)
            const uint32_t q0 = REV ? (L - B - (uint32_t)(s0 + 8)) : (B + (uint32_t)s0);  // emitted position, multiple of 8
            int k;
            uint32_t nxt;
            if (RAND_DWELL) {
                k = ts.map[w];
                nxt = ts.off[k + 1];
            } else {
                const int sc = max(s0, 0);
                k = sc / p.sps_fixed;
                nxt = (uint32_t)(k + 1) * (uint32_t)p.sps_fixed;
            }
            float2 ab = ts.par[k];
            uint4 r0 = make_uint4(0, 0, 0, 0);
            if (NOISY) r0 = philox4x32_10(q0 >> 3, key.r_lo, key.r_hi, ST_AMP, key.k0, key.k1);
            const uint32_t rw[4] = {r0.x, r0.y, r0.z, r0.w};
            uint32_t v[8];
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const int s = s0 + j;
                const int e = REV ? 7 - j : j;  // slot in the emitted chunk
                uint32_t bits = 0;
                if (s >= 0 && (uint32_t)s < S) {
                    if ((uint32_t)s >= nxt) {
                        k++;
                        nxt = RAND_DWELL ? ts.off[k + 1] : nxt + (uint32_t)p.sps_fixed;
                        ab = ts.par[k];
                    }
                    if (NOISY) {
                        const uint32_t h = (e & 1) ? (rw[e >> 1] >> 16) : (rw[e >> 1] & 0xFFFFu);
                        const float z = z16(z1s, p.z2, h, q0 + e, key, ST_AMP_TAIL);
                        bits = to_i16_bits(fmaf(z, ab.x, ab.y));
                    } else {
                        bits = __float_as_uint(ab.y);
                    }
                }
                v[e] = bits;
            }
            if (s0 >= 0 && (uint32_t)(s0 + 8) <= S) {
                const uint4 pk = make_uint4(v[0] | (v[1] << 16), v[2] | (v[3] << 16), v[4] | (v[5] << 16), v[6] | (v[7] << 16));
                __stcs(reinterpret_cast<uint4 *>(out + q0), pk);
            } else {
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const int s = s0 + j;
                    const int e = REV ? 7 - j : j;
                    if (s >= 0 && (uint32_t)s < S) out[q0 + e] = (int16_t)v[e];
                }
            }
        }
        __syncthreads();  // tile state is reused by the next tile
    }
}

// RNA --prefix: lower the adaptor region (src/genread.c:80-86).  Emitted positions [L-n0, L-n0+shift_len).
__global__ void __launch_bounds__(256) prefix_shift_kernel(const __grid_constant__ GenParams p) {
    const int r = blockIdx.x;
    const ReadDesc rd = p.reads[r];
    if (rd.shift_len <= 0) return;
    const uint32_t L = p.read_siglen[r], n0 = p.read_n0[r];
    const uint32_t len = min((uint32_t)rd.shift_len, n0);
    int16_t *out = p.sig + p.read_sigoff[r] + (L - n0);
    for (uint32_t i = threadIdx.x; i < len; i += blockDim.x) out[i] = (int16_t)(out[i] - (int16_t)p.shift_val);
}

// store-only kernel: the HBM write ceiling next to which the signal kernel is read
__global__ void __launch_bounds__(512) store_only_kernel(uint4 *dst, size_t n16) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    for (size_t i = t; i < n16; i += stride) __stcs(dst + i, make_uint4(t, t + 1, t + 2, (uint32_t)i));
}

}  // namespace sqg
