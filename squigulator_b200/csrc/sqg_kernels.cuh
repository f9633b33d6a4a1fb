// sqg_kernels.cuh — the CUDA kernels of the signal-generation path (sm_100a, hand-written).
//
// Work decomposition (DESIGN.md "Kernels"):
//   a read is 1-2 SEGMENTS (the 2nd only for the RNA stall of --prefix, src/genread.c:88-89);
//   a segment is cut into TILES of T consecutive k-mers; a tile is what one thread group turns into samples.
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
    const __half *z16;    // Z16[65536]
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
    float dwell_mean, dwell_std;
    int32_t sps_fixed;    // fixed-dwell modes: samples per k-mer
    uint32_t sps_magic;   // floor(2^32/sps_fixed)+1: n/sps_fixed == umulhi(n, magic) for n*sps_fixed < 2^32
    int32_t ideal;        // SQ_IDEAL: per-read draws replaced by the means
    float amp_noise;
    uint32_t key0, key1;
    int64_t first_read;
    int32_t want_ss;
    int32_t shift_val;  // (int16)(30*digitisation/range)
};

constexpr int K1_THREADS = 128;
constexpr int GROUPS = 2;    // independent thread groups per CTA: own tile state, own named barrier
constexpr int GT = 384;      // threads per group
constexpr int K4_THREADS = GROUPS * GT;
constexpr int MAX_T = 1024;  // k-mers per tile
constexpr int MAP_CAP = 4096;  // 8-sample chunks per tile (random dwell)

// ------------------------------------------------------------------------------------------------
// small helpers

__device__ __forceinline__ RngKey make_key(const GenParams &p, int read_local) {
    const uint64_t r = (uint64_t)(p.first_read + read_local);
    return RngKey{p.key0, p.key1, (uint32_t)r, (uint32_t)(r >> 32)};
}

__device__ __forceinline__ uint32_t halfword(const uint4 &w, int j) {  // j in 0..7, compile-time after unrolling
    const uint32_t x = (j >> 1) == 0 ? w.x : (j >> 1) == 1 ? w.y : (j >> 1) == 2 ? w.z : w.w;
    return (j & 1) ? (x >> 16) : (x & 0xFFFFu);
}

// n / sps_fixed for tile-local sample numbers (exact: see GenParams::sps_magic)
__device__ __forceinline__ uint32_t div_sps(const GenParams &p, uint32_t n) {
    return p.sps_fixed == 1 ? n : __umulhi(n, p.sps_magic);
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
// K1: per-tile sum of dwells (random-dwell modes only).  One CTA per tile, one thread per Philox block of 8 k-mers.
__global__ void __launch_bounds__(K1_THREADS) dwell_sum_kernel(const __grid_constant__ GenParams p) {
    __shared__ int s_seg;
    __shared__ uint32_t s_part[K1_THREADS / 32];
    const int tile = blockIdx.x;
    if (threadIdx.x == 0) s_seg = find_seg(p.segs, p.n_segs, tile);
    __syncthreads();
    const int si = s_seg;
    const SegDesc seg = p.segs[si];
    const int kstart = (tile - seg.tile0) * p.T;
    const int nk_tile = min(p.T, seg.nk - kstart);
    const RngKey key = make_key(p, seg.read);
    uint32_t sum = 0;
    const int g = threadIdx.x;
    if (g * 8 < nk_tile) {
        const uint32_t blk = (uint32_t)((seg.k0_rng + kstart) >> 3) + g;
        const uint4 w = philox4x32_10(blk, key.r_lo, key.r_hi, ST_DWELL, key.k0, key.k1);
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const float z = z16(p.z16, p.z2, halfword(w, j), blk * 8 + j, key, ST_DWELL_TAIL);
            const int d = dwell_from_z(z, p.dwell_mean, p.dwell_std);
            if (g * 8 + j < nk_tile) sum += (uint32_t)d;
        }
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
            const float za = z16(p.z16, p.z2, ww[d] & 0xFFFFu, 2 * d, key, ST_READ_TAIL);
            const float zb = z16(p.z16, p.z2, ww[d] >> 16, 2 * d + 1, key, ST_READ_TAIL);
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
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) raw += __shfl_xor_sync(0xffffffffu, raw, o);
    if (tid == 0) p.meta[1] = 0;
    __syncthreads();
    if ((tid & 31) == 0) atomicAdd((unsigned long long *)&p.meta[1], (unsigned long long)raw);
    if (tid == 0) p.meta[0] = (int64_t)s_total;
}

// ------------------------------------------------------------------------------------------------
// K4: the signal kernel.

struct __align__(16) TileState {
    uint32_t off[MAX_T + 8];       // tile-relative first sample of each k-mer; off[nk] = S       (random dwell)
    float2 par[MAX_T];             // NOISY: (A', B'); else (0, int16 value bits)
    uint16_t map[MAP_CAP];         // chunk -> k-mer of its first sample                           (random dwell)
    uint32_t bmap[MAP_CAP / 4];    // byte w: bit b set <=> a k-mer starts at sample b of chunk w   (random dwell)
    uint8_t digit[MAX_T + 16];     // base digits of the tile's window
    uint32_t warp_sum[GT / 32];
    uint32_t S;
    uint32_t pad[3];
};

struct __align__(16) CtaShared {
    TileState ts[GROUPS];
    uint4 lut[256];     // boundary mask -> byte offsets (8 * k-mers passed) of the 8 samples, 16 bit each
    uint8_t code[256];  // base -> digits
    unsigned long long mbar;
    uint32_t pad[2];
};

// ---- raw shared-memory access by 32-bit shared address (lets ptxas use LDS [R+UR+imm]) ----
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float lds_half(uint32_t addr) {
    unsigned short h;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(h) : "r"(addr));
    return __half2float(__ushort_as_half(h));
}
__device__ __forceinline__ float2 lds_f2(uint32_t addr) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ void group_sync(int group) {
    asm volatile("bar.sync %0, %1;" ::"r"(group + 1), "n"(GT) : "memory");
}

// TMA bulk copy global -> shared (1-D), completion on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_load_1d(void *smem_dst, const void *gsrc, uint32_t bytes, unsigned long long *mbar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(mbar))
                 : "memory");
}
__device__ __forceinline__ void mbar_init(unsigned long long *mbar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(mbar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *mbar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(mbar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *mbar, uint32_t phase) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE_%=;\n"
        "bra WAIT_LOOP_%=;\n"
        "WAIT_DONE_%=:\n"
        "}\n" ::"r"(smem_u32(mbar)),
        "r"(phase)
        : "memory");
}

// Everything phase B needs to know about the tile
struct TileCtx {
    uint32_t S;    // samples in the tile
    uint32_t ph;   // chunk w covers tile samples [8w-ph, 8w-ph+8)
    uint32_t B;    // first logical sample of the tile within the read
    uint32_t L;    // samples in the read
    int16_t *out;  // start of the read in the signal arena
    RngKey key;
};

// Generic (slow) chunk: partial chunks at the tile edges.  Walks the k-mers sample by sample.
template <bool NOISY, bool RAND_DWELL, bool REV>
__device__ __noinline__ void slow_chunk(const GenParams &p, const TileState &ts, const __half *__restrict__ z16s,
                                        const TileCtx c, uint32_t w) {
    const int s0 = (int)(8 * w) - (int)c.ph;
    const uint32_t q0 = REV ? (c.L - c.B - (uint32_t)(s0 + 8)) : (c.B + (uint32_t)s0);
    const int sc = max(s0, 0);
    int k;
    uint32_t nxt;
    if (RAND_DWELL) {
        k = ts.map[w];
        nxt = ts.off[k + 1];
    } else {
        k = (int)div_sps(p, (uint32_t)sc);
        nxt = (uint32_t)(k + 1) * (uint32_t)p.sps_fixed;
    }
    uint4 rw = make_uint4(0, 0, 0, 0);
    if (NOISY) rw = philox4x32_10(q0 >> 3, c.key.r_lo, c.key.r_hi, ST_AMP, c.key.k0, c.key.k1);
    for (int j = 0; j < 8; j++) {
        const int s = s0 + j;
        const int e = REV ? 7 - j : j;
        if (s < 0 || (uint32_t)s >= c.S) continue;
        while ((uint32_t)s >= nxt) {
            k++;
            nxt = RAND_DWELL ? ts.off[k + 1] : nxt + (uint32_t)p.sps_fixed;
        }
        const float2 ab = ts.par[k];
        uint32_t bits;
        if (NOISY) {
            const uint32_t x = (e >> 1) == 0 ? rw.x : (e >> 1) == 1 ? rw.y : (e >> 1) == 2 ? rw.z : rw.w;
            const uint32_t h = (e & 1) ? (x >> 16) : (x & 0xFFFFu);
            const float z = z16(z16s, p.z2, h, q0 + e, c.key, ST_AMP_TAIL);
            bits = to_i16_bits(fmaf(z, ab.x, ab.y));
        } else {
            bits = __float_as_uint(ab.y);
        }
        c.out[q0 + e] = (int16_t)bits;
    }
}

// dynamic shared memory: [CtaShared][Z16: 128 KB if USE_Z][model: num_kmer*8 B if model_in_smem]
template <bool NOISY, bool RAND_DWELL, bool METH, bool REV>
__global__ void __launch_bounds__(K4_THREADS, 1) signal_kernel(const __grid_constant__ GenParams p) {
    constexpr bool USE_Z = NOISY || RAND_DWELL;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    CtaShared &cs = *reinterpret_cast<CtaShared *>(smem_raw);
    __half *z16s = reinterpret_cast<__half *>(smem_raw + ((sizeof(CtaShared) + 127) & ~127u));
    float2 *models = reinterpret_cast<float2 *>(reinterpret_cast<unsigned char *>(z16s) + (USE_Z ? Z16_N * 2 : 0));
    const int tid = threadIdx.x;
    const int group = tid / GT;
    const int gtid = tid - group * GT;
    const int lane = tid & 31, gwarp = gtid >> 5;
    TileState &ts = cs.ts[group];

    // ---- one-time staging: quantile table and (small) pore model by TMA bulk copies ----
    if (tid == 0) mbar_init(&cs.mbar, 1);
    for (int i = tid; i < 256; i += K4_THREADS) {
        cs.code[i] = base_code(i);
        uint32_t f[8], cnt = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            if (j >= 1 && ((i >> j) & 1)) cnt++;
            f[j] = cnt * 8;
        }
        cs.lut[i] = make_uint4(f[0] | (f[1] << 16), f[2] | (f[3] << 16), f[4] | (f[5] << 16), f[6] | (f[7] << 16));
    }
    __syncthreads();
    if (tid == 0) {
        uint32_t bytes = 0;
        if (USE_Z) bytes += Z16_N * 2;
        if (p.model_in_smem) bytes += p.num_kmer * 8;
        if (bytes) {
            mbar_expect_tx(&cs.mbar, bytes);
            if (USE_Z) {
                tma_load_1d(z16s, p.z16, Z16_N, &cs.mbar);  // two 64 KB bulk copies
                tma_load_1d(z16s + Z16_N / 2, p.z16 + Z16_N / 2, Z16_N, &cs.mbar);
            }
            if (p.model_in_smem) tma_load_1d(models, p.model, p.num_kmer * 8, &cs.mbar);
        }
    }
    if (USE_Z || p.model_in_smem) mbar_wait(&cs.mbar, 0);
    const float2 *__restrict__ model = p.model_in_smem ? models : p.model;
    const uint32_t zbase = smem_u32(z16s);
    const uint32_t par_addr = smem_u32(ts.par);  // < 64 KB: the tile states lead the shared-memory layout
    if (par_addr + (MAX_T + 8) * 8 >= 0x10000u) __trap();

    for (int tile = blockIdx.x * GROUPS + group; tile < p.n_tiles; tile += gridDim.x * GROUPS) {
        const int si = p.tile_seg[tile];
        const SegDesc seg = p.segs[si];
        const int kstart = (tile - seg.tile0) * p.T;
        const int nk_tile = min(p.T, seg.nk - kstart);
        TileCtx c;
        c.key = make_key(p, seg.read);
        c.B = p.tile_base[tile];  // first logical sample of the tile within the read
        c.L = p.read_siglen[seg.read];
        c.out = p.sig + p.read_sigoff[seg.read];
        // chunk w covers tile samples [8w-ph, 8w-ph+8): chunks are aligned in the EMITTED signal
        c.ph = REV ? ((c.B - c.L) & 7u) : (c.B & 7u);
        const double offset = p.read_offset[seg.read];

        // ---- phase A0: digits of the tile's base window; clear the boundary bitmap ----
        const int nb = nk_tile + p.k - 1;
        for (int i = gtid; i < nb; i += GT) {
            const int pos = kstart + i;
            const uint8_t ch = pos < seg.len_a ? p.bases[seg.off_a + pos] : p.bases[seg.off_b + (pos - seg.len_a)];
            const uint8_t code = cs.code[ch];
            ts.digit[i] = METH ? (code >> 4) : (code & 3);
        }
        if (RAND_DWELL)
            for (int i = gtid; i < MAP_CAP / 16; i += GT) reinterpret_cast<uint4 *>(ts.bmap)[i] = make_uint4(0, 0, 0, 0);
        group_sync(group);

        // ---- phase A1: 4 k-mers per thread: rank, model lookup, dwell, parameters ----
        const int m0 = gtid * 4;  // first k-mer of this thread within the tile (T/4 <= 256 <= GT)
        const bool active = m0 < nk_tile;
        int d[4] = {0, 0, 0, 0};
        uint32_t local = 0;
        float2 pr[4];
        if (active) {
            if (RAND_DWELL) {
                const uint32_t kidx = (uint32_t)(seg.k0_rng + kstart + m0);  // draw index of the first k-mer (multiple of 4)
                const uint4 w = philox4x32_10(kidx >> 3, c.key.r_lo, c.key.r_hi, ST_DWELL, c.key.k0, c.key.k1);
                const uint32_t wa = (kidx & 4) ? w.z : w.x, wb = (kidx & 4) ? w.w : w.y;
                const uint32_t hh[4] = {wa & 0xFFFFu, wa >> 16, wb & 0xFFFFu, wb >> 16};
#pragma unroll
                for (int j = 0; j < 4; j++)
                    d[j] = dwell_from_z(z16(z16s, p.z2, hh[j], kidx + j, c.key, ST_DWELL_TAIL), p.dwell_mean, p.dwell_std);
            } else {
#pragma unroll
                for (int j = 0; j < 4; j++) d[j] = p.sps_fixed;
            }
            uint32_t rank = 0;
            const uint8_t *dg = ts.digit + m0;
            for (int i = 0; i < p.k - 1; i++) rank = METH ? rank * 5 + dg[i] : (rank << 2) | dg[i];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const bool valid = m0 + j < nk_tile;
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
        uint32_t o = 0, S;
        if (RAND_DWELL) {
            // group-wide exclusive scan of `local`
            uint32_t inc = local;
#pragma unroll
            for (int sh = 1; sh < 32; sh <<= 1) {
                const uint32_t v = __shfl_up_sync(0xffffffffu, inc, sh);
                if (lane >= sh) inc += v;
            }
            if (lane == 31) ts.warp_sum[gwarp] = inc;
            group_sync(group);
            if (gwarp == 0) {
                const uint32_t ws = lane < GT / 32 ? ts.warp_sum[lane] : 0;
                uint32_t winc = ws;
#pragma unroll
                for (int sh = 1; sh < 16; sh <<= 1) {
                    const uint32_t v = __shfl_up_sync(0xffffffffu, winc, sh);
                    if (lane >= sh) winc += v;
                }
                if (lane < GT / 32) ts.warp_sum[lane] = winc - ws;
                if (lane == GT / 32 - 1) ts.S = winc;
            }
            group_sync(group);
            S = ts.S;
            o = ts.warp_sum[gwarp] + inc - local;
        } else {
            S = (uint32_t)nk_tile * (uint32_t)p.sps_fixed;
            o = (uint32_t)m0 * (uint32_t)p.sps_fixed;
        }
        c.S = S;
        if (active) {
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int m = m0 + j;
                if (m < nk_tile) {
                    ts.par[m] = pr[j];
                    if (RAND_DWELL) {
                        ts.off[m] = o;
                        // chunks whose first (clipped) sample falls inside this k-mer
                        uint32_t w0 = m == 0 ? 0u : (o + c.ph + 7) >> 3;
                        const uint32_t w1 = (o + (uint32_t)d[j] + c.ph + 7) >> 3;
                        for (; w0 < w1; w0++) ts.map[w0] = (uint16_t)m;
                        const uint32_t pos = o + c.ph;  // boundary bit: this k-mer starts at bit (pos&7) of chunk (pos>>3)
                        atomicOr(&ts.bmap[pos >> 5], 1u << (pos & 31));
                    }
                    if (p.want_ss) p.ss[p.reads[seg.read].ss_off + seg.k0 + kstart + m] = d[j];
                    o += (uint32_t)d[j];
                }
            }
            if (RAND_DWELL && m0 + 4 >= nk_tile) ts.off[nk_tile] = S;
        }
        group_sync(group);

        // ---- phase B: samples.  One thread = one 16-byte chunk of the emitted signal. ----
        const uint32_t nW = (S + c.ph + 7) >> 3;
        for (uint32_t w = gtid; w < nW; w += GT) {
            const int s0 = (int)(8 * w) - (int)c.ph;  // first tile sample of the chunk (may be < 0)
            if (s0 < 0 || (uint32_t)(s0 + 8) > S) {
                slow_chunk<NOISY, RAND_DWELL, REV>(p, ts, z16s, c, w);
                continue;
            }
            const uint32_t q0 = REV ? (c.L - c.B - (uint32_t)(s0 + 8)) : (c.B + (uint32_t)s0);  // emitted position, multiple of 8
            // k-mer of the first sample and the boundary mask of the chunk
            uint32_t k0, bm;
            if (RAND_DWELL) {
                k0 = ts.map[w];
                bm = reinterpret_cast<const uint8_t *>(ts.bmap)[w] & 0xFEu;
            } else {
                k0 = div_sps(p, (uint32_t)s0);
                bm = 0;
                for (uint32_t b = (k0 + 1) * (uint32_t)p.sps_fixed - (uint32_t)s0; b < 8; b += (uint32_t)p.sps_fixed) bm |= 1u << b;
            }
            const uint4 lu = cs.lut[bm];
            const uint32_t rep = (k0 * 8 + par_addr) * 0x00010001u;
            const uint32_t pa[4] = {lu.x + rep, lu.y + rep, lu.z + rep, lu.w + rep};
            uint32_t v[8];
            if (NOISY) {
                const uint4 r4 = philox4x32_10(q0 >> 3, c.key.r_lo, c.key.r_hi, ST_AMP, c.key.k0, c.key.k1);
                const uint32_t rw[4] = {r4.x, r4.y, r4.z, r4.w};
                float zmax = 0.f;
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const int e = REV ? 7 - j : j;  // slot in the emitted chunk = which 16-bit draw
                    const float2 ab = lds_f2((j & 1) ? (pa[j >> 1] >> 16) : (pa[j >> 1] & 0xFFFFu));
                    const uint32_t h = (e & 1) ? (rw[e >> 1] >> 16) : (rw[e >> 1] & 0xFFFFu);
                    const float z = lds_half(zbase + 2u * h);  // extract + one multiply-add for the address
                    zmax = fmaxf(zmax, fabsf(z));
                    v[e] = (uint32_t)__float2int_rz(fmaf(z, ab.x, ab.y));
                }
                if (__builtin_expect(zmax >= Z_TAIL_THR, 0)) {
                    // rare: some draw fell into one of the 16 outermost cells -> refine it (10 more bits)
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        const int e = REV ? 7 - j : j;
                        const uint32_t h = (e & 1) ? (rw[e >> 1] >> 16) : (rw[e >> 1] & 0xFFFFu);
                        if ((h & 0x7FFFu) >= Z_TAIL_FIRST) {
                            const float2 ab = lds_f2((j & 1) ? (pa[j >> 1] >> 16) : (pa[j >> 1] & 0xFFFFu));
                            const float z = z16_tail(p.z2, h, q0 + e, c.key, ST_AMP_TAIL);
                            v[e] = (uint32_t)__float2int_rz(fmaf(z, ab.x, ab.y));
                        }
                    }
                }
            } else {
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const int e = REV ? 7 - j : j;
                    const float2 ab = lds_f2((j & 1) ? (pa[j >> 1] >> 16) : (pa[j >> 1] & 0xFFFFu));
                    v[e] = __float_as_uint(ab.y);
                }
            }
            // low 16 bits of each int32 (the reference's wrap, src/gensig.c:270), packed little-endian
            const uint4 pk = make_uint4(__byte_perm(v[0], v[1], 0x5410), __byte_perm(v[2], v[3], 0x5410),
                                        __byte_perm(v[4], v[5], 0x5410), __byte_perm(v[6], v[7], 0x5410));
            __stcs(reinterpret_cast<uint4 *>(c.out + q0), pk);
        }
        group_sync(group);  // the tile state is reused by this group's next tile
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
