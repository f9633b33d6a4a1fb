// sqg_kernels.cuh — the CUDA kernels of the signal-generation path (sm_100a, hand-written).
//
// Work decomposition (DESIGN.md "Kernels"):
//   a read is 1-2 SEGMENTS (the 2nd only for the RNA stall of --prefix, src/genread.c:88-89);
//   a segment is cut into TILES of T consecutive k-mers; a tile is what one thread group turns into samples.
//
//   K0 tile_desc_kernel    per segment (warp): one 80-byte descriptor per tile                   -> tiles
//   K1 dwell_kernel        per tile (warp)  : draw the T dwells (Philox + table normals)         -> dwells (u16), tile_sum, ss
//   K2 read_plan_kernel    per read (warp)  : exclusive scan of its tile sums, per-read draws    -> tiles.B/S/L/offset, siglen, offset, median_before
//   K3 read_offsets_kernel one CTA          : exclusive scan of the 64-sample-aligned lengths    -> sigoff, totals
//   K4 signal_kernel       per tile (warp)  : encode k-mers, gather (mean,stdv), chunk map, then emit every int16
//                          sample with 128-bit stores                                            -> signal  (the hot kernel)
//
// Reference statements: src/gensig.c:226-288 (gen_sig_core_seq), :293-343 (gen_sig_core), :346-356 (gen_sig).
#pragma once
#include <type_traits>

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

// Everything the dwell pass and the signal kernel need to know about a tile, in one 80-byte record (written by K0,
// completed by K2) so that a warp reaches its bases, its dwells and its place in the output with independent loads.
struct __align__(16) TileDesc {
    int64_t a_off;   // window byte i < a_rem is bases[a_off + i]   (piece a of the segment)
    int64_t b_off;   // window byte i >= a_rem is bases[b_off + i]  (piece b)
    int32_t a_rem;   // may be <= 0 or beyond the window
    int32_t nk;      // k-mers in the tile
    int32_t read;    // local read index
    uint32_t kidx0;  // dwell draw index of the tile's first k-mer (multiple of 8)
    int64_t ss_pos;  // where the tile's dwells go in ss[]
    uint32_t B;      // first logical sample of the tile within the read (K2)
    uint32_t S;      // samples in the tile (K2)
    double offset;   // the read's ADC offset (K2)
    uint32_t L;      // samples in the read (K2)
    uint32_t pad;
    uint32_t r_lo, r_hi;  // global read index = Philox counter words 1, 2 (K0)
    uint32_t pad2[2];
};
static_assert(sizeof(TileDesc) == 80, "TileDesc is loaded as five 16-byte words");

struct GenParams {
    // inputs
    const uint8_t *bases;
    const SegDesc *segs;
    const ReadDesc *reads;
    const float2 *model;  // (level_mean, level_stdv) by rank
    const float4 *pair_model;  // by (k+1)-mer rank: the parameters of its two k-mers (first k bases, last k bases); base-4 models only
    const float4 *quad_model;  // k <= 6: by (k+3)-mer rank, 32-byte entries: the parameters of its four k-mers (else NULL)
    const float *z32;     // Z32[32768]
    const float *z2;      // Z2[8192]
    // plan (written by K0-K3, read by K4)
    TileDesc *tiles;
    uint32_t *tile_sum;
    uint4 *dwells;        // per tile: TK dwells as uint16 (32 x uint4), written by K1 (random-dwell modes)
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
    int32_t wide;      // 1: the model/profile does not guarantee 16384 <= sample + 32768 < 131072 -> exact path for every sample
    // profile (src/sq.h:47-58) and options
    double digitisation, range, scale;  // scale = digitisation/range
    double offset_mean, offset_std, median_mean, median_std;
    float dwell_mean, dwell_std;
    int32_t sps_fixed;    // fixed-dwell modes: samples per k-mer
    uint32_t sps_magic;   // floor(2^32/sps_fixed)+1: n/sps_fixed == umulhi(n, magic) for n*sps_fixed < 2^32
    int32_t ideal;        // SQ_IDEAL: per-read draws replaced by the means
    float amp_noise;
    uint32_t key0, key1;
    uint32_t rk[2 * PHILOX_ROUNDS];  // the Philox round keys (k0 + r*W0, k1 + r*W1), precomputed
    int64_t first_read;
    int32_t want_ss;
    int32_t shift_val;  // (int16)(30*digitisation/range)
};

constexpr int TK = 256;          // k-mers per tile at most (32 lanes x 8); also the row length of GenParams::dwells

// ------------------------------------------------------------------------------------------------
// small helpers

__device__ __forceinline__ RngKey make_key(const GenParams &p, int read_local) {
    const uint64_t r = (uint64_t)(p.first_read + read_local);
    return RngKey{p.key0, p.key1, (uint32_t)r, (uint32_t)(r >> 32)};
}

// n / sps_fixed for tile-local sample numbers (exact: see GenParams::sps_magic)
__device__ __forceinline__ uint32_t div_sps(const GenParams &p, uint32_t n) {
    return p.sps_fixed == 1 ? n : __umulhi(n, p.sps_magic);
}

// ------------------------------------------------------------------------------------------------
// init: the pore model indexed by (k+1)-mer.  Two consecutive k-mers of a read overlap in k-1 bases, so one 16-byte
// entry addressed by the (k+1)-mer they span holds the parameters of both: the signal kernel's model gathers - one
// 32-byte sector request each, the scarcest resource of its k-mer phase - are halved.  4^(k+1) x 16 B (16 MB for 9-mers).
__global__ void __launch_bounds__(256) pair_model_kernel(const float2 *__restrict__ model, float4 *__restrict__ pair, uint32_t n_pair, uint32_t kmask) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pair) return;
    const float2 a = model[i >> 2], b = model[i & kmask];
    pair[i] = make_float4(a.x, a.y, b.x, b.y);
}

// k <= 6: the same idea one step further - a (k+3)-mer spans four consecutive k-mers, 4 x 8 bytes = exactly one 32-byte
// sector, fetched with one 256-bit load (LDG.E.256); 4^(k+3) x 32 B = 8 MB for 6-mers.
__global__ void __launch_bounds__(256) quad_model_kernel(const float2 *__restrict__ model, float4 *__restrict__ quad, uint32_t n_quad, uint32_t kmask) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_quad) return;
    const float2 a = model[(i >> 6) & kmask], b = model[(i >> 4) & kmask], c = model[(i >> 2) & kmask], d = model[i & kmask];
    quad[2 * (size_t)i] = make_float4(a.x, a.y, b.x, b.y);
    quad[2 * (size_t)i + 1] = make_float4(c.x, c.y, d.x, d.y);
}

// ------------------------------------------------------------------------------------------------
// K0: tile descriptors (one warp per segment, one lane per tile)
__global__ void __launch_bounds__(256) tile_desc_kernel(const __grid_constant__ GenParams p) {
    const int si = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (si >= p.n_segs) return;
    const SegDesc seg = p.segs[si];
    const int64_t ss0 = p.reads[seg.read].ss_off + seg.k0;
    const RngKey key = make_key(p, seg.read);
    const int nt = (seg.nk + p.T - 1) / p.T;
    for (int t = lane; t < nt; t += 32) {
        const int kstart = t * p.T;
        TileDesc d;
        d.a_off = seg.off_a + kstart;
        d.b_off = seg.off_b + kstart - seg.len_a;
        d.a_rem = seg.len_a - kstart;
        d.nk = min(p.T, seg.nk - kstart);
        d.read = seg.read;
        d.kidx0 = (uint32_t)(seg.k0_rng + kstart);
        d.ss_pos = ss0 + kstart;
        d.B = 0; d.S = 0; d.offset = 0.0; d.L = 0; d.pad = 0;
        d.r_lo = key.r_lo; d.r_hi = key.r_hi;
        d.pad2[0] = d.pad2[1] = 0;
        p.tiles[seg.tile0 + t] = d;
    }
}

// K1: the dwells (random-dwell modes only; src/gensig.c:255-256).  Persistent CTAs (one per SM, 32 warps) with the
// quantile table staged in shared memory by TMA; a warp takes a tile at a time, one lane per Philox block of 8 k-mers:
// 8 table normals -> 8 dwells, stored as one 16-byte row piece of uint16 (k-mers beyond the tile get 0), the tile's
// sample count, and - when asked for - the reference's aln->ss (src/gensig.c:273-281).
constexpr int K1_THREADS = 1024;
constexpr uint32_t K1_SMEM = Z32_BYTES + 16;
__device__ __forceinline__ uint32_t smem_u32(const void *p);
__device__ __forceinline__ void tma_load_1d(void *smem_dst, const void *gsrc, uint32_t bytes, unsigned long long *mbar);
__device__ __forceinline__ void mbar_init(unsigned long long *mbar, uint32_t count);
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *mbar, uint32_t bytes);
__device__ __forceinline__ void mbar_wait(unsigned long long *mbar, uint32_t parity);

__global__ void __launch_bounds__(K1_THREADS, 1) dwell_kernel(const __grid_constant__ GenParams p) {
    extern __shared__ __align__(128) unsigned char smem1[];
    unsigned long long *bar = reinterpret_cast<unsigned long long *>(smem1 + Z32_BYTES);
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        mbar_expect_tx(bar, Z32_BYTES);
        tma_load_1d(smem1, p.z32, Z32_BYTES / 2, bar);
        tma_load_1d(smem1 + Z32_BYTES / 2, reinterpret_cast<const unsigned char *>(p.z32) + Z32_BYTES / 2, Z32_BYTES / 2, bar);
    }
    __syncthreads();
    mbar_wait(bar, 0);
    const int lane = threadIdx.x & 31;
    const int nw = K1_THREADS / 32;
    const int tile0 = blockIdx.x * nw + (threadIdx.x >> 5), tstride = gridDim.x * nw;
    if (tile0 >= p.n_tiles) return;
    // the descriptor words of the NEXT tile are loaded one iteration ahead (the kernel is otherwise bound by this latency)
    const uint4 *q0 = reinterpret_cast<const uint4 *>(p.tiles + tile0);
    uint4 nb = __ldg(q0 + 1), nc = __ldg(q0 + 2), ne = __ldg(q0 + 4);
    for (int tile = tile0; tile < p.n_tiles; tile += tstride) {
        const uint4 b = nb, c = nc, e = ne;
        {
            const uint4 *qn = reinterpret_cast<const uint4 *>(p.tiles + min(tile + tstride, p.n_tiles - 1));
            nb = __ldg(qn + 1); nc = __ldg(qn + 2); ne = __ldg(qn + 4);
        }
        const int nk_tile = (int)b.y;
        const uint32_t kidx0 = b.w;
        const int64_t ss_pos = (int64_t)(((uint64_t)c.y << 32) | c.x);
        const RngKey key{p.key0, p.key1, e.x, e.y};
        uint32_t sum = 0;
        uint32_t d[8];
#pragma unroll
        for (int j = 0; j < 8; j++) d[j] = 0;
        if (lane * 8 < nk_tile) {
            const uint32_t blk = (kidx0 >> 3) + lane;
            const uint4 w = philox4x32_rk(blk, key.r_lo, key.r_hi, ST_DWELL, p.rk);
            const int64_t ss0 = ss_pos + lane * 8;
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const uint32_t off = z_offset(draw_word(w, j), (blk & 31u) << 2);
                float z = *reinterpret_cast<const float *>(smem1 + off);
                if (__builtin_expect(z_is_tail(off), 0)) z = z_tail(p.z2, off, blk * 8 + j, key, ST_DWELL_TAIL);
                if (lane * 8 + j < nk_tile) {
                    d[j] = (uint32_t)dwell_from_z(z, p.dwell_mean, p.dwell_std);
                    if (p.want_ss) p.ss[ss0 + j] = (int32_t)d[j];
                }
                sum += d[j];
            }
        }
        p.dwells[(size_t)tile * (TK / 8) + lane] = make_uint4(d[0] | (d[1] << 16), d[2] | (d[3] << 16), d[4] | (d[5] << 16), d[6] | (d[7] << 16));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        if (lane == 0) p.tile_sum[tile] = sum;
    }
}

// fixed-dwell modes with aln->ss requested: every k-mer has sps_fixed samples
__global__ void __launch_bounds__(256) fixed_ss_kernel(const __grid_constant__ GenParams p, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p.ss[i] = p.sps_fixed;
}

// ------------------------------------------------------------------------------------------------
// K2: per read (one warp, one lane per tile) — scan its tiles, draw offset / median_before (src/gensig.c:312-318),
// complete the tile descriptors
template <bool RAND_DWELL>
__global__ void __launch_bounds__(256) read_plan_kernel(const __grid_constant__ GenParams p) {
    const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= p.n_reads) return;
    const ReadDesc rd = p.reads[r];
    uint64_t total = 0;
    uint32_t n0 = 0;
    for (int s = rd.seg0; s < rd.seg0 + rd.nseg; s++) {
        const SegDesc seg = p.segs[s];
        const int ntile = (seg.nk + p.T - 1) / p.T;
        for (int t0 = 0; t0 < ntile; t0 += 32) {
            const int t = t0 + lane;
            const int tile = seg.tile0 + t;
            uint32_t sum = 0;
            if (t < ntile) {
                if (RAND_DWELL) {
                    sum = p.tile_sum[tile];
                } else {
                    sum = (uint32_t)min(p.T, seg.nk - t * p.T) * (uint32_t)p.sps_fixed;
                    p.tile_sum[tile] = sum;
                }
            }
            uint64_t inc = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint64_t v = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += v;
            }
            if (t < ntile) {
                p.tiles[tile].B = (uint32_t)(total + inc - sum);
                p.tiles[tile].S = sum;
            }
            total += __shfl_sync(0xffffffffu, inc, 31);
        }
        if (s == rd.seg0) n0 = (uint32_t)total;
    }
    if (total >= 0xFFFFFFFFull) {  // src/sim.c:559-562
        if (lane == 0) atomicExch((unsigned long long *)&p.meta[2], 1ull);
        total = 0;
    }
    double off = p.offset_mean, med = p.median_mean;
    if (!p.ideal) {
        // one Philox block per read: draws 0-3 -> offset, 4-7 -> median_before, each a unit-norm mix of four table
        // normals (weights: cos/sin products of 35, 40, 55 degrees); classes walk with the read index
        const RngKey key = make_key(p, r);
        const uint4 w = philox4x32(0u, key.r_lo, key.r_hi, ST_READ, key.k0, key.k1);
        const double W4[4] = {0.6275068715971331, 0.43938504177070503, 0.3686878264946124, 0.5265407845183632};
        double z[2];
#pragma unroll
        for (int d = 0; d < 2; d++) {
            double t[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int j = 4 * d + u;
                const float zz = z_global(p.z32, p.z2, z_offset(draw_word(w, j), ((8u * key.r_lo + j) & 31u) << 2), j, key, ST_READ_TAIL);
                t[u] = __dmul_rn((double)zz, W4[u]);
            }
            z[d] = __dadd_rn(__dadd_rn(t[0], t[1]), __dadd_rn(t[2], t[3]));
        }
        off = __dadd_rn(__dmul_rn(z[0], p.offset_std), p.offset_mean);
        med = __dadd_rn(__dmul_rn(z[1], p.median_std), p.median_mean);
    }
    if (lane == 0) {
        p.read_siglen[r] = (uint32_t)total;
        p.read_n0[r] = n0;
        p.read_offset[r] = off;
        p.read_median[r] = med;
    }
    for (int s = rd.seg0; s < rd.seg0 + rd.nseg; s++) {
        const SegDesc seg = p.segs[s];
        const int ntile = (seg.nk + p.T - 1) / p.T;
        for (int t = lane; t < ntile; t += 32) {
            TileDesc *td = p.tiles + seg.tile0 + t;
            td->offset = off;
            td->L = (uint32_t)total;
        }
    }
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
//
// Every WARP is autonomous: it owns a private tile buffer in shared memory and walks its own sequence of tiles,
// alternating a k-mer phase (A: latency-bound — descriptor, base window, dwells, model gathers, warp scan) and a sample
// phase (B: issue-bound — Philox, table normals, FFMA, 128-bit stores).  There are no CTA-wide barriers and no
// producer/consumer hand-offs after the prologue; the 16 resident warps of an SM are at different points of their
// tiles, so phase-A latency of some warps is covered by phase-B work of the others.  The CTA shares only read-only
// tables: the 128 KB float quantile table staged by TMA bulk copies, the boundary LUT and the base-code table.
//   phase A, lane = 8 consecutive k-mers: their dwells (one 16-byte load of what K1 drew), warp scan, chunk->k-mer map +
//            boundary bitmap; digits -> ranks -> (mean,stdv) gathers -> (A', B'+32768) into par[]
//   phase B, lane = one 16-byte chunk (8 samples) of the emitted signal per iteration: k-mer of the first sample from
//            the map, the parameters of that k-mer and the next two, the chunk's boundary byte -> predicates, one
//            Philox block -> 8 table normals (bank-stratified lookups) -> fma.rz (+ predicated fma.rz for samples past
//            a boundary) -> PRMT of the mantissas -> one 128-bit store
//
// Shared memory (byte offsets into the dynamic array, all compile-time so that they fold into LDS/STS immediates):
//   [par: one TK x float2 array per warp]  [Z32: 128 KB]  [code: 256 B]  [mbar]  [per warp: map MAPC u8, bmap MAPC u8, digits]

#ifndef SQG_K4_WARPS
#define SQG_K4_WARPS 16
#endif
constexpr int K4_WARPS = SQG_K4_WARPS;
constexpr int K4_THREADS = K4_WARPS * 32;  // register budget: 65536 / 512 = 128
constexpr int MAPC = 1344;   // 8-sample chunks per tile: >= (TK*max_dwell + 14)/8 (dna-r10: 256 k-mers x 37)
constexpr int DIG_BYTES = TK + 32;   // digits of the tile's base window; the same size holds the raw window (16-byte granules)
// per-warp buffer
constexpr uint32_t W_MAP = 0;                        // per 32 samples (4 chunks) an 8-byte entry: {bit s = a k-mer (other than the
                                                     // tile's first) starts at sample s, number of such starts before the entry}
constexpr uint32_t W_DIG = W_MAP + 2 * MAPC;         // base digits
constexpr uint32_t W_RAW = W_DIG + DIG_BYTES;        // prefetched base window (ASCII), 16-byte granules
constexpr uint32_t W_DWELL = W_RAW + DIG_BYTES;      // prefetched dwells of the tile: TK x uint16
constexpr uint32_t W_DESC = W_DWELL + TK * 2;        // two TileDesc slots
constexpr uint32_t W_SIGOFF = W_DESC + 2 * 80;       // two int64 slots
constexpr uint32_t WARP_BYTES = W_SIGOFF + 16;
constexpr uint32_t SM_PAR = 0;
constexpr uint32_t PAR_BYTES = TK * 8 + 48;   // + rows the sample loop may load (never use) past the tile's last k-mer; 16-byte multiple
constexpr uint32_t SM_Z = (SM_PAR + K4_WARPS * PAR_BYTES + 127) & ~127u;
constexpr uint32_t SM_CODE = SM_Z + Z32_BYTES;
constexpr uint32_t SM_MBAR = SM_CODE + 256;
constexpr uint32_t SM_WARP = SM_MBAR + 16;
constexpr uint32_t SM_TOTAL = SM_WARP + K4_WARPS * WARP_BYTES;
static_assert(SM_Z % 128 == 0 && SM_WARP % 16 == 0 && WARP_BYTES % 16 == 0 && MAPC % 16 == 0 && DIG_BYTES % 16 == 0, "alignment");
static_assert(SM_TOTAL <= 232448, "227 KB of shared memory per CTA");

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier / TMA bulk copy for the one-time table staging (SASS: SYNCS, UBLKCP) ----
__device__ __forceinline__ void tma_load_1d(void *smem_dst, const void *gsrc, uint32_t bytes, unsigned long long *mbar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(mbar))
                 : "memory");
}
__device__ __forceinline__ void mbar_init(unsigned long long *mbar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(mbar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *mbar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(mbar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *mbar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE_%=;\n"
        "bra WAIT_LOOP_%=;\n"
        "WAIT_DONE_%=:\n"
        "}\n" ::"r"(smem_u32(mbar)),
        "r"(parity)
        : "memory");
}
// ---- per-lane asynchronous global -> shared copies (SASS: LDGSTS) for the next tile's inputs ----
__device__ __forceinline__ void cp_async16(uint32_t saddr, const void *g) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(saddr), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async8(uint32_t saddr, const void *g) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(saddr), "l"(g) : "memory");
}
// shared-window address of a pointer, computed behind an opaque asm: the compiler must not tie the (vector-register)
// addresses of the asynchronous copies to the base of the ordinary shared-memory accesses, which it keeps in a uniform
// register ([R + UR + imm] addressing in the sample loop)
__device__ __forceinline__ uint32_t opaque_smem_addr(const void *sptr) {
    uint32_t a;
    asm volatile("{ .reg .u64 t; cvta.to.shared.u64 t, %1; cvt.u32.u64 %0, t; }" : "=r"(a) : "l"(sptr));
    return a;
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }
#ifndef SQG_ST_POLICY
#define SQG_ST_POLICY ".cs"   // streaming (evict-first) stores: the signal is written once and never read back by the kernel
#endif
__device__ __forceinline__ void st_cs_v4(void *gptr, uint4 v) {
    asm volatile("st.global" SQG_ST_POLICY ".v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(gptr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// ---- shared-memory loads of the sample loop, by absolute shared-window address with the constant part as an
// immediate.  The dynamic array starts right after the driver's reserved kilobyte (cudaDevAttrReservedSharedMemoryPerBlock;
// this kernel has no static shared memory), so `offset + SMEM_ORIGIN + constant` needs no base register: one LOP3 makes
// the table offset and the load takes it as is.  The kernel prologue checks the origin and refuses to run otherwise. ----
constexpr uint32_t SMEM_ORIGIN = 0x400;
template <uint32_t IMM>
__device__ __forceinline__ float lds_f32(uint32_t off) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(v) : "r"(off), "n"(SMEM_ORIGIN + IMM) : "memory");
    return v;
}
template <uint32_t IMM>
__device__ __forceinline__ float2 lds_f2(uint32_t off) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2+%3];" : "=f"(v.x), "=f"(v.y) : "r"(off), "n"(SMEM_ORIGIN + IMM) : "memory");
    return v;
}
template <uint32_t IMM>
__device__ __forceinline__ uint2 lds_u2(uint32_t off) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2+%3];" : "=r"(v.x), "=r"(v.y) : "r"(off), "n"(SMEM_ORIGIN + IMM) : "memory");
    return v;
}
// k-mer of a chunk's first sample and the chunk's boundary mask from its map entry (chunk w = byte w & 3 of the entry)
__device__ __forceinline__ void entry_kmers(uint2 ent, uint32_t sh /* 8 * (w & 3) */, uint32_t &k0, uint32_t &m1) {
    k0 = ent.y + __popc(ent.x & ((2u << sh) - 1u));   // starts at or before the chunk's first sample
    m1 = (ent.x >> sh) & 0xFEu;                        // (a start on the chunk's first sample is not a boundary to cross)
}

// What phase B needs to know about the tile (registers, warp-uniform)
struct TileCtx {
    uint32_t S;      // samples in the tile
    uint32_t ph;     // chunk w covers tile samples [8w-ph, 8w-ph+8)
    uint32_t C0;     // emitted chunk (= Philox block) of the tile's chunk 0; chunk w is C0+w, or C0-w when reversed
    uint32_t r_lo, r_hi;  // global read index (Philox counter words 1,2)
    int16_t *out;    // start of the read in the signal arena
};

// ---- phase B ---------------------------------------------------------------------------------------------------

// k-mer of a chunk's first sample and the boundary mask (bit j, 1..7: a k-mer starts at tile-order slot j)
template <bool RAND_DWELL>
__device__ __forceinline__ void chunk_kmers(const GenParams &p, const unsigned char *smem, uint32_t map_off, const TileCtx &t, uint32_t w,
                                            uint32_t &k0, uint32_t &m1) {
    if (RAND_DWELL) {
        entry_kmers(*reinterpret_cast<const uint2 *>(smem + map_off + W_MAP + 8 * (w >> 2)), 8 * (w & 3), k0, m1);
    } else {
        const int s0 = (int)(8 * w) - (int)t.ph;
        k0 = div_sps(p, (uint32_t)max(s0, 0));
        m1 = 0;
        for (int b = (int)((k0 + 1) * (uint32_t)p.sps_fixed) - s0; b < 8; b += p.sps_fixed) m1 |= 1u << b;
    }
}

// The exact path of one chunk, start to finish (rare: a clipped chunk at an end of the tile, a chunk with a flagged
// sample - tail cell of the table, negative value, value beyond int16 - or with four or more k-mers, and every chunk in
// wide mode): samples are trunc(fma.rz(z, A', Bq)) with the tail cells refined and any number of boundaries; only the
// tile's own samples are stored (the neighbouring tile computes the same Philox block and stores the rest).
// par[] holds B'+32768 (or Bq itself in wide mode).
template <bool NOISY, bool RAND_DWELL, bool REV>
__device__ __noinline__ void exact_chunk(const GenParams &p, const unsigned char *smem, TileCtx t, uint32_t par_off, uint32_t map_off,
                                         uint32_t w) {
    uint32_t k0, m1;
    chunk_kmers<RAND_DWELL>(p, smem, map_off, t, w, k0, m1);
    const uint32_t par0 = k0 * 8 + par_off;
    const uint32_t Cq = REV ? t.C0 - w : t.C0 + w;
    const uint32_t class4 = (Cq & 31u) << 2;
    const RngKey key{p.key0, p.key1, t.r_lo, t.r_hi};
    uint4 r4 = make_uint4(0, 0, 0, 0);
    if (NOISY) r4 = philox4x32_rk(Cq, t.r_lo, t.r_hi, ST_AMP, p.rk);
    const float sub = p.wide ? 0.f : SAMPLE_MAGIC;
    int16_t *dst = t.out + (size_t)Cq * 8;
#pragma unroll
    for (int e = 0; e < 8; e++) {
        const int j = REV ? 7 - e : e;
        if (8 * w + j - t.ph < t.S) {
            const float2 ab = *reinterpret_cast<const float2 *>(smem + par0 + 8 * __popc(m1 & ((2u << j) - 1u)));
            uint32_t v;
            if (NOISY) {
                const uint32_t off = z_offset(draw_word(r4, e), class4);
                float z = *reinterpret_cast<const float *>(smem + SM_Z + off);
                if (z_is_tail(off)) z = z_tail(p.z2, off, Cq * 8 + e, key, ST_AMP_TAIL);
                v = sample_exact(z, ab.x, __fsub_rn(ab.y, sub));
            } else {
                v = __float_as_uint(ab.y);
            }
            dst[e] = (int16_t)v;
        }
    }
}

// One chunk = 8 consecutive samples = at most 3 k-mers on the fast path: the parameters of k-mers k0, k0+1, k0+2 are
// loaded once (three 8-byte loads) and every sample picks its own by PREDICATE - the k-mer boundaries inside the
// chunk arrive as a bit mask, `mask-1` has its bits clear exactly from the first boundary upwards, and one R2P moves
// seven of those bits into predicate registers - so a sample costs one FFMA plus at most two predicated ones on the
// FMA pipe, and no shared-memory traffic of its own.  The loop body has no rare path inside: a lane only notes which
// of its chunks needs the exact path (flagged sample, 4+ k-mers) and redoes it after the loop.
template <bool NOISY, bool RAND_DWELL, bool REV>
__device__ __forceinline__ void emit_tile(const GenParams &p, const unsigned char *smem, const TileCtx t, int lane, uint32_t par_off,
                                          uint32_t map_off) {
    const uint32_t nW = (t.S + t.ph + 7) >> 3;      // chunks touched by the tile (the first and last may be clipped)
    const uint32_t w_lo = (t.ph + 7) >> 3;          // chunks [w_lo, w_hi) lie wholly inside the tile
    const uint32_t w_hi = (t.S + t.ph) >> 3;
    if (NOISY && p.wide) {
        for (uint32_t w = lane; w < nW; w += 32) exact_chunk<NOISY, RAND_DWELL, REV>(p, smem, t, par_off, map_off, w);
        return;
    }
    const uint32_t class4 = ((REV ? t.C0 - (uint32_t)lane : t.C0 + (uint32_t)lane) & 31u) << 2;  // same for all chunks of a lane
    uint32_t n_redo = 0, w_redo = 0;
    // Software pipeline, one chunk deep: the Philox block and the map bytes of the lane's NEXT chunk are produced while
    // the table lookups and FFMAs of the current one are in flight (two independent dependency chains per warp).
    uint4 r4n = make_uint4(0, 0, 0, 0);
    uint2 entn = make_uint2(0, 0);
    const uint32_t ent_sh = 8 * (lane & 3);                       // w = lane (mod 32): the chunk's byte within its map entry
    const uint32_t ent_lane = map_off + 8 * ((uint32_t)lane >> 2);
    {
        const uint32_t w = lane;
        if (NOISY) r4n = philox4x32_rk(REV ? t.C0 - w : t.C0 + w, t.r_lo, t.r_hi, ST_AMP, p.rk);
        if (RAND_DWELL) entn = lds_u2<W_MAP>(ent_lane);
    }
    // One iteration = 32 consecutive chunks (wb is warp-uniform).  EDGE iterations hold a clipped chunk or run past the
    // tile's end; all others - the bulk - carry no bounds logic at all.
    auto iteration = [&](const uint32_t wb, auto edge_tag) {
        constexpr bool EDGE = decltype(edge_tag)::value;
        const uint32_t w = wb + lane;
        if (EDGE && w >= nW) return;
        const uint4 r4 = r4n;
        uint32_t k0 = 0, m1 = 0;
        if (RAND_DWELL) entry_kmers(entn, ent_sh, k0, m1);
        {
            const uint32_t wn = w + 32;  // (past the tile's end for the last one: computed, never used)
#ifdef SQG_KO_PHILOX
            if (NOISY) r4n = make_uint4(wn * 0x9E3779B9u, wn * 0x85EBCA6Bu, wn * 0xC2B2AE35u, wn * 0x27D4EB2Fu);
#else
            if (NOISY) r4n = philox4x32_rk(REV ? t.C0 - wn : t.C0 + wn, t.r_lo, t.r_hi, ST_AMP, p.rk);
#endif
            if (RAND_DWELL) entn = lds_u2<W_MAP>(ent_lane + 2 * wb + 64);   // entry of chunk wn = 8 * (wn >> 2)
        }
        if (!RAND_DWELL) chunk_kmers<RAND_DWELL>(p, smem, map_off, t, w, k0, m1);
        const uint32_t par0 = k0 * 8 + par_off;
#ifdef SQG_KO_PAR
        const float2 q0 = lds_f2<0>(par_off + 8 * (w & 1)), q1 = q0, q2 = q0;
#else
        const float2 q0 = lds_f2<0>(par0), q1 = lds_f2<8>(par0), q2 = lds_f2<16>(par0);
#endif
        const uint32_t t1 = m1 - 1u;        // bit j clear  <=>  slot j lies at or after the 1st boundary
        const uint32_t m2 = m1 & t1;        // boundaries after the first
        const uint32_t t2 = m2 - 1u;        // bit j clear  <=>  slot j lies at or after the 2nd boundary
        const uint32_t m3 = m2 & t2;        // non-zero: a 3rd boundary -> exact path
        const uint32_t Cq = REV ? t.C0 - w : t.C0 + w;
        uint4 pk;
        uint32_t bad;
        if (NOISY) {
            float zz[8], v[8];
#pragma unroll
            for (int e = 0; e < 8; e++) {   // e = slot in the emitted chunk = which draw; j = slot in tile order
#ifdef SQG_KO_Z
                zz[e] = __uint_as_float((z_offset(draw_word(r4, e), class4) & 0x7FFFFFu) | 0x3F000000u);
#else
                zz[e] = lds_f32<SM_Z>(z_offset(draw_word(r4, e), class4));
#endif
                v[e] = fma_rz(zz[e], q0.x, q0.y);
            }
#pragma unroll
            for (int e = 0; e < 8; e++) {   // level by level, so that one R2P per level sets the predicates
                const int j = REV ? 7 - e : e;
                if (j >= 1 && !(t1 & (1u << j))) v[e] = fma_rz(zz[e], q1.x, q1.y);
            }
#pragma unroll
            for (int e = 0; e < 8; e++) {
                const int j = REV ? 7 - e : e;
                if (j >= 2 && !(t2 & (1u << j))) v[e] = fma_rz(zz[e], q2.x, q2.y);
            }
            uint32_t u[8];
#pragma unroll
            for (int e = 0; e < 8; e++) u[e] = __float_as_uint(v[e]);
            // bits 8..23 of each float, packed little-endian
            pk = make_uint4(__byte_perm(u[0], u[1], 0x6521), __byte_perm(u[2], u[3], 0x6521),
                            __byte_perm(u[4], u[5], 0x6521), __byte_perm(u[6], u[7], 0x6521));
            bad = ((pk.x | pk.y | pk.z | pk.w) & 0x80008000u) | m3;
        } else {
            uint32_t v[8];
#pragma unroll
            for (int e = 0; e < 8; e++) {
                const int j = REV ? 7 - e : e;
                float x = q0.y;
                if (j >= 1 && !(t1 & (1u << j))) x = q1.y;
                if (j >= 2 && !(t2 & (1u << j))) x = q2.y;
                v[e] = __float_as_uint(x);
            }
            // low 16 bits of each int32 (the reference's wrap, src/gensig.c:270), packed little-endian
            pk = make_uint4(__byte_perm(v[0], v[1], 0x5410), __byte_perm(v[2], v[3], 0x5410),
                            __byte_perm(v[4], v[5], 0x5410), __byte_perm(v[6], v[7], 0x5410));
            bad = m3;
        }
        int16_t *dst = t.out + (size_t)Cq * 8;
#ifdef SQG_KO_STORE
        if (pk.x == 0x12345678u && pk.y == 0x9abcdef0u) st_cs_v4(dst, pk);
        if (bad != 0) { n_redo++; w_redo = w; }
        return;
#endif
        if (!EDGE || w - w_lo < w_hi - w_lo) {
            st_cs_v4(dst, pk);
        } else {
            // clipped chunk at an end of the tile (at most two per tile): store only the tile's own samples; the
            // neighbouring tile computes the same Philox block and stores the rest
            const uint32_t pw[4] = {pk.x, pk.y, pk.z, pk.w};
#pragma unroll
            for (int e = 0; e < 8; e++) {
                const int j = REV ? 7 - e : e;
                if (8 * w + j - t.ph < t.S) dst[e] = (int16_t)((e & 1) ? (pw[e >> 1] >> 16) : pw[e >> 1]);
            }
        }
        if (bad != 0) { n_redo++; w_redo = w; }
    };
    const uint32_t it_mid0 = (w_lo + 31) >> 5, it_mid1 = w_hi >> 5, it_end = (nW + 31) >> 5;   // clean iterations [it_mid0, it_mid1)
    uint32_t it = 0;
    for (; it < min(it_mid0, it_end); it++) iteration(32 * it, std::true_type{});
    for (; it < it_mid1; it++) iteration(32 * it, std::false_type{});
    for (; it < it_end; it++) iteration(32 * it, std::true_type{});
    if (__builtin_expect(n_redo != 0, 0)) {
        if (n_redo == 1) {
            exact_chunk<NOISY, RAND_DWELL, REV>(p, smem, t, par_off, map_off, w_redo);
        } else {
            for (uint32_t w = lane; w < nW; w += 32) exact_chunk<NOISY, RAND_DWELL, REV>(p, smem, t, par_off, map_off, w);
        }
    }
}

// ---- phase A ---------------------------------------------------------------------------------------------------

__device__ __forceinline__ TileDesc read_tile_desc(const unsigned char *smem, uint32_t off) {
    const uint4 *q = reinterpret_cast<const uint4 *>(smem + off);
    const uint4 a = q[0], b = q[1], c = q[2], d = q[3], e = q[4];
    TileDesc t;
    t.a_off = (int64_t)(((uint64_t)a.y << 32) | a.x);
    t.b_off = (int64_t)(((uint64_t)a.w << 32) | a.z);
    t.a_rem = (int32_t)b.x; t.nk = (int32_t)b.y; t.read = (int32_t)b.z; t.kidx0 = b.w;
    t.ss_pos = (int64_t)(((uint64_t)c.y << 32) | c.x);
    t.B = c.z; t.S = c.w;
    t.offset = __longlong_as_double((long long)(((uint64_t)d.y << 32) | d.x));
    t.L = d.z; t.pad = 0;
    t.r_lo = e.x; t.r_hi = e.y; t.pad2[0] = t.pad2[1] = 0;
    return t;
}

constexpr int WIN_LOADS = (TK + 8 + 31) / 32;  // bytes per lane of a tile's base window (k <= 9)

// a tile whose base window straddles the two pieces of its segment (only around a --prefix junction)
__device__ __forceinline__ bool tile_is_junction(const GenParams &p, int32_t a_rem, int32_t nk) {
    return a_rem > 0 && a_rem < nk + p.k - 1;
}

// Asynchronous fetch of a tile's inputs into the warp's buffer: the base window as 16-byte granules (the aligned
// superset of the window), its dwells, its read's arena offset.  `desc_off` = the tile's descriptor, already in smem.
template <bool RAND_DWELL>
__device__ __forceinline__ void fetch_tile_inputs(const GenParams &p, const unsigned char *smem, uint32_t wbase, uint32_t map_off,
                                                  uint32_t desc_off, uint32_t sigoff_off, int tile, int lane) {
    const uint4 a = *reinterpret_cast<const uint4 *>(smem + desc_off);
    const uint4 b = *reinterpret_cast<const uint4 *>(smem + desc_off + 16);
    const int32_t a_rem = (int32_t)b.x, nk = (int32_t)b.y, read = (int32_t)b.z;
    if (!tile_is_junction(p, a_rem, nk)) {
        const int64_t off = a_rem > 0 ? (int64_t)(((uint64_t)a.y << 32) | a.x) : (int64_t)(((uint64_t)a.w << 32) | a.z);
        const uint8_t *g = p.bases + off;
        const uint32_t shift = (uint32_t)(reinterpret_cast<uintptr_t>(g) & 15u);
        const uint32_t nbytes = shift + (uint32_t)(nk + p.k - 1);
        if ((uint32_t)lane * 16 < nbytes) cp_async16(wbase + W_RAW + lane * 16, g - shift + lane * 16);
    }
    if (RAND_DWELL) cp_async16(wbase + W_DWELL + lane * 16, p.dwells + (size_t)tile * (TK / 8) + lane);
    if (lane == 0) cp_async8(wbase + (sigoff_off - map_off), p.read_sigoff + read);
}
__device__ __forceinline__ void fetch_tile_desc(const GenParams &p, uint32_t wbase, uint32_t desc_rel, int tile, int lane) {
    if (lane < 5) cp_async16(wbase + desc_rel + lane * 16, reinterpret_cast<const uint4 *>(p.tiles + tile) + lane);
}

// Phase A of one tile.  Its descriptor, base window, dwells and arena offset are already in the warp's buffer.
template <bool NOISY, bool RAND_DWELL, bool METH, bool REV, bool QUAD>
__device__ __forceinline__ TileCtx prepare_tile(const GenParams &p, int lane, unsigned char *smem, uint32_t par_off, uint32_t map_off,
                                                uint32_t desc_off, uint32_t sigoff_off) {
    const TileDesc td = read_tile_desc(smem, desc_off);
    const int nk_tile = td.nk;
    const int nb = nk_tile + p.k - 1;
    const uint32_t dig_off = map_off + W_DIG;
    const uint32_t B = td.B, L = td.L, S = td.S;
    const uint32_t ph = REV ? ((B - L) & 7u) : (B & 7u);
    const int m0 = lane * 8;

    // (1) bases -> digits.  Fast path (base-4 models, window in one piece): every lane takes the 16 raw bytes of its own
    // 8 k-mers straight from the prefetched window (three aligned 8-byte loads + a funnel shift by the window's
    // misalignment) and turns A/C/G/T of either case into digits arithmetically, ((c>>1) ^ (c>>2)) & 3, four bytes at
    // a time; a PRMT maps the digits back to letters to check that every byte really was one of those eight.  Any
    // other byte in the tile (IUPAC codes, U, N: src/seq.h:14-28 folds them) sends the whole warp through the
    // 256-entry code table, which is also the path of base-5 (CpG) models (src/seq.h:45-60) and of prefix junctions.
    uint32_t dg[4] = {0, 0, 0, 0};   // the lane's 16 digits, one per byte
    bool table_path = METH || tile_is_junction(p, td.a_rem, nk_tile);
    if (!table_path) {
        const uint32_t shift = (uint32_t)(reinterpret_cast<uintptr_t>(p.bases + (td.a_rem > 0 ? td.a_off : td.b_off)) & 15u);
        const uint32_t s0 = shift + 8u * (uint32_t)lane;     // lane's first byte within the raw buffer
        const uint2 *rw = reinterpret_cast<const uint2 *>(smem + map_off + W_RAW + (s0 & ~7u));
        const uint2 w0 = rw[0], w1 = rw[1], w2 = rw[2];
        const bool hi = (s0 & 4u) != 0;                       // (warp-uniform: shift & 4)
        const uint32_t q0 = hi ? w0.y : w0.x, q1 = hi ? w1.x : w0.y, q2 = hi ? w1.y : w1.x, q3 = hi ? w2.x : w1.y, q4 = hi ? w2.y : w2.x;
        const uint32_t fs = 8u * (s0 & 3u);
        const uint32_t x[4] = {__funnelshift_r(q0, q1, fs), __funnelshift_r(q1, q2, fs), __funnelshift_r(q2, q3, fs), __funnelshift_r(q3, q4, fs)};
        uint32_t bad = 0;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const uint32_t c = ((x[i] >> 1) ^ (x[i] >> 2)) & 0x03030303u;
            const uint32_t t = (c | (c >> 4)) & 0x00FF00FFu;
            const uint32_t sel = (t | (t >> 8)) & 0xFFFFu;                    // the four digits as PRMT selectors
            bad |= __byte_perm(0x54474341u /* "ACGT" */, 0u, sel) ^ (x[i] & 0xDFDFDFDFu);
            dg[i] = c;
        }
        // bytes past the window's end are whatever the 16-byte granules held: harmless as digits, but they must not
        // force the table path, so only the lane's bytes inside the window count
        const int inside = nb - 8 * lane;
        if (inside < 16) {
            if (inside <= 0) bad = 0;
            else {
                // re-derive per word: keep the flags of the first `inside` bytes
                uint32_t keep = 0;
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const int nbytes = min(max(inside - 4 * i, 0), 4);
                    const uint32_t m = nbytes >= 4 ? 0xFFFFFFFFu : ((1u << (8 * nbytes)) - 1u);
                    const uint32_t c = dg[i];
                    const uint32_t t = (c | (c >> 4)) & 0x00FF00FFu;
                    const uint32_t sel = (t | (t >> 8)) & 0xFFFFu;
                    keep |= (__byte_perm(0x54474341u, 0u, sel) ^ (x[i] & 0xDFDFDFDFu)) & m;
                }
                bad = keep;
            }
        }
        table_path = __any_sync(0xffffffffu, bad != 0);
    }
    if (table_path) {
        if (!tile_is_junction(p, td.a_rem, nk_tile)) {
            const uint32_t shift = (uint32_t)(reinterpret_cast<uintptr_t>(p.bases + (td.a_rem > 0 ? td.a_off : td.b_off)) & 15u);
            const uint32_t raw_off = map_off + W_RAW + shift + lane;
#pragma unroll 1
            for (int u = 0; u < WIN_LOADS; u++) {
                const int i = lane + 32 * u;
                if (i < nb) {
                    const uint32_t c = smem[SM_CODE + smem[raw_off + 32 * u]];
                    smem[dig_off + i] = (unsigned char)(METH ? (c >> 4) : (c & 3u));
                }
            }
        } else {
#pragma unroll 1
            for (int i = lane; i < nb; i += 32) {
                const uint32_t c = smem[SM_CODE + __ldg(p.bases + (i < td.a_rem ? td.a_off : td.b_off) + i)];
                smem[dig_off + i] = (unsigned char)(METH ? (c >> 4) : (c & 3u));
            }
        }
        __syncwarp();
        const uint2 dwa = *reinterpret_cast<const uint2 *>(smem + dig_off + m0);
        const uint2 dwb = *reinterpret_cast<const uint2 *>(smem + dig_off + m0 + 8);
        dg[0] = dwa.x; dg[1] = dwa.y; dg[2] = dwb.x; dg[3] = dwb.y;
        __syncwarp();   // (the digit buffer is rewritten by this warp's next tile)
    }

    // (1b) base-4 models: the four (k+1)-mer gathers of this lane are issued now, so that their L2 latency runs under
    // the shared-memory work of step (2).  16 two-bit digits packed first-digit-most-significant
    // (((w & 0x03030303) * 0x40100401) >> 24 packs 4 bytes); k-mers 2j, 2j+1 of the lane = the two k-mers of the
    // (k+1)-mer at digit 2j: ONE 16-byte gather for both.  The four pairs are visited in ROTATED order
    // jj(j) = (j + lane/2) & 3 so that the 16-byte parameter stores of a quarter-warp fall into 8 different bank groups.
    float4 mv4[4];
    const int rot4 = lane >> 1;
    if (!METH) {
        const uint32_t P = ((((dg[0] & 0x03030303u) * 0x40100401u) >> 24) << 24) | ((((dg[1] & 0x03030303u) * 0x40100401u) >> 24) << 16) |
                           ((((dg[2] & 0x03030303u) * 0x40100401u) >> 24) << 8) | (((dg[3] & 0x03030303u) * 0x40100401u) >> 24);
        const int sh0 = 30 - 2 * p.k;                 // 32 - 2(k+1)
        const uint32_t pmask = (p.kmask << 2) | 3u;   // 4^(k+1) - 1
        if (QUAD) {
            // k <= 6: two 256-bit gathers, each the four k-mers of one (k+3)-mer = one whole sector
            const int shq = 26 - 2 * p.k;             // 32 - 2(k+3)
            const uint32_t qmask = (p.kmask << 6) | 63u;
#pragma unroll
            for (int h = 0; h < 2; h++) {
                uint32_t r = (P >> (shq - 8 * h)) & qmask;
                if (nk_tile < TK && m0 + 4 * h >= nk_tile) r = 0;
                const float4 *src = p.quad_model + 2 * (size_t)r;
                asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                             : "=f"(mv4[2 * h].x), "=f"(mv4[2 * h].y), "=f"(mv4[2 * h].z), "=f"(mv4[2 * h].w), "=f"(mv4[2 * h + 1].x),
                               "=f"(mv4[2 * h + 1].y), "=f"(mv4[2 * h + 1].z), "=f"(mv4[2 * h + 1].w)
                             : "l"(src));
            }
        } else
#pragma unroll
        for (int j = 0; j < 4; j++) {
            uint32_t r = (P >> (sh0 - 4 * ((j + rot4) & 3))) & pmask;
            if (nk_tile < TK && m0 + 2 * ((j + rot4) & 3) >= nk_tile) r = 0;
#ifdef SQG_KO_GATHER
            mv4[j] = __ldg(&p.pair_model[(r & 0x3F) * 0 + ((td.kidx0 >> 3) & 0xFF) * 128 + j * 32 + lane]);   // coalesced (timing only)
#else
            mv4[j] = __ldg(&p.pair_model[r]);
#endif
        }
    }

    // (2) warp scan of the dwells, then the chunk -> k-mer map and the boundary bitmap
    if (RAND_DWELL) {
        const uint4 dq = *reinterpret_cast<const uint4 *>(smem + map_off + W_DWELL + lane * 16);
        // clear the map entries this tile touches (+ the one-ahead read of the sample loop)
        const uint32_t n_ent = min((((S + ph + 7) >> 3) + 3) / 4 + 9u, (uint32_t)(MAPC / 4));
        for (uint32_t e = lane * 2; e < n_ent; e += 64) *reinterpret_cast<uint4 *>(smem + map_off + W_MAP + 8 * e) = make_uint4(0, 0, 0, 0);
        const uint32_t t4 = dq.x + dq.y + dq.z + dq.w;  // packed halves: no carry, every dwell < 2^14
        const uint32_t local = (t4 & 0xFFFFu) + (t4 >> 16);
        uint32_t inc = local;
#pragma unroll
        for (int sh = 1; sh < 32; sh <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, inc, sh);
            if (lane >= sh) inc += v;
        }
        __syncwarp();
#ifndef SQG_KO_MAP
        // (2a) one bit per k-mer start (the tile's first k-mer excepted: chunks before any bit belong to it)
        const uint32_t dw[4] = {dq.x, dq.y, dq.z, dq.w};
        uint32_t pos = inc - local + ph;  // frame position of the k-mer's first sample: bit (pos & 31) of entry pos >> 5
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const uint32_t dj = (j & 1) ? (dw[j >> 1] >> 16) : (dw[j >> 1] & 0xFFFFu);
            if (dj != 0 && m0 + j != 0)
                atomicOr(reinterpret_cast<uint32_t *>(smem + map_off + W_MAP + ((pos >> 5) << 3)), 1u << (pos & 31u));
            pos += dj;
        }
        __syncwarp();
        // (2b) per entry: the number of starts before it.  Per round of 128 entries a lane takes entries 2l, 2l+1 and
        // 64+2l, 64+2l+1 (two conflict-free 16-byte accesses); the two halves are scanned together, their counts packed
        // in the halves of one register (every count < 2^16).
        uint32_t carry = 0;
        for (uint32_t e0 = 0; e0 < n_ent; e0 += 128) {
            uint4 *ep = reinterpret_cast<uint4 *>(smem + map_off + W_MAP + 8 * e0 + 16 * lane);
            const bool in1 = e0 + 2 * lane < (uint32_t)(MAPC / 4), in2 = e0 + 64 + 2 * lane < (uint32_t)(MAPC / 4);
            uint4 a = make_uint4(0, 0, 0, 0), b = make_uint4(0, 0, 0, 0);
            if (in1) asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w) : "r"(smem_u32(ep)) : "memory");
            if (in2) asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "r"(smem_u32(ep + 32)) : "memory");
            const uint32_t c0 = __popc(a.x), c1 = __popc(a.z), c2 = __popc(b.x), c3 = __popc(b.z);
            const uint32_t mine = (c0 + c1) | ((c2 + c3) << 16);
            uint32_t run = mine;
#pragma unroll
            for (int sh = 1; sh < 32; sh <<= 1) {
                const uint32_t v = __shfl_up_sync(0xffffffffu, run, sh);
                if (lane >= sh) run += v;
            }
            const uint32_t tot = __shfl_sync(0xffffffffu, run, 31);
            const uint32_t before1 = carry + (run & 0xFFFFu) - (c0 + c1);
            const uint32_t before2 = carry + (tot & 0xFFFFu) + (run >> 16) - (c2 + c3);
            if (in1) ep[0] = make_uint4(a.x, before1, a.z, before1 + c0);
            if (in2) ep[32] = make_uint4(b.x, before2, b.z, before2 + c2);
            carry += (tot & 0xFFFFu) + (tot >> 16);
        }
#endif
    } else {
        __syncwarp();
    }

    // (3) the parameters of this lane's 8 k-mers from (level_mean, level_stdv); base-5 (CpG) models: ranks
    // (src/seq.h:62-74) and single gathers here
    const float scale_f = (float)p.scale, off_f = (float)td.offset;
    auto make_par = [&](float mean, float stdv) -> float2 {
        if (NOISY) {
            // single precision, rounded once each: A' = (stdv*amp_noise)*scale, B' = fma(mean, scale, -offset), then
            // B' + 32768 (the sample arithmetic's magic offset; wide mode keeps Bq = (B' + 32768) - 32768 itself)
            const float Bm = __fadd_rn(fmaf(mean, scale_f, -off_f), SAMPLE_MAGIC);
            return make_float2(__fmul_rn(__fmul_rn(stdv, p.amp_noise), scale_f), p.wide ? __fsub_rn(Bm, SAMPLE_MAGIC) : Bm);
        } else {
            // src/gensig.c:266,270: (double)level_mean*digitisation/range - offset, truncated
            const double v = __dsub_rn(__ddiv_rn(__dmul_rn((double)mean, p.digitisation), p.range), td.offset);
            return make_float2(0.f, __uint_as_float(to_i16_bits(v)));
        }
    };
    {
        if (!METH) {
            if (m0 < nk_tile) {
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const float2 pa = make_par(mv4[j].x, mv4[j].y), pb = make_par(mv4[j].z, mv4[j].w);
                    const int piece = QUAD ? j : ((j + rot4) & 3);   // (the 256-bit gathers arrive in k-mer order)
                    *reinterpret_cast<float4 *>(smem + par_off + 8 * (m0 + 2 * piece)) = make_float4(pa.x, pa.y, pb.x, pb.y);
                }
            }
        } else {
            const uint32_t dw[4] = {dg[0], dg[1], dg[2], dg[3]};
            const int km1 = p.k - 1;
            uint32_t rank = 0, ranks[8];
#pragma unroll
            for (int i = 0; i < 8; i++)
                if (i < km1) rank = rank * 5 + ((dw[i >> 2] >> (8 * (i & 3))) & 0xFFu);
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const int bi = km1 + j;
                const uint32_t word = bi < 4 ? dw[0] : bi < 8 ? dw[1] : bi < 12 ? dw[2] : dw[3];
                rank = (rank % p.kmask) * 5 + ((word >> (8 * (bi & 3))) & 0xFFu);
                ranks[j] = (m0 + j < nk_tile) ? rank : 0u;
            }
            float2 mv[8];
#pragma unroll
            for (int j = 0; j < 8; j++) mv[j] = __ldg(&p.model[ranks[j]]);
            if (m0 < nk_tile) {
#pragma unroll
                for (int j = 0; j < 8; j++) *reinterpret_cast<float2 *>(smem + par_off + 8 * (m0 + j)) = make_par(mv[j].x, mv[j].y);
            }
        }
    }
    TileCtx h;
    h.S = S; h.ph = ph;
    h.C0 = REV ? ((L - B + ph) >> 3) - 1u : (B >> 3);
    h.r_lo = td.r_lo; h.r_hi = td.r_hi;
    h.out = p.sig + *reinterpret_cast<const int64_t *>(smem + sigoff_off);
    __syncwarp();
    return h;
}

template <bool NOISY, bool RAND_DWELL, bool METH, bool REV, bool QUAD>
__global__ void __launch_bounds__(K4_THREADS, 1) signal_kernel(const __grid_constant__ GenParams p) {
    constexpr bool USE_Z = NOISY;
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    unsigned long long *stage_bar = reinterpret_cast<unsigned long long *>(smem + SM_MBAR);
    const uint32_t par_off = SM_PAR + (uint32_t)warp * PAR_BYTES;
    const uint32_t map_off = SM_WARP + (uint32_t)warp * WARP_BYTES;
    const uint32_t wbase = opaque_smem_addr(smem + map_off);  // this warp's buffer, for the asynchronous copies

    // ---- prologue: tables; this warp's first descriptor ----
    const int gwarp = blockIdx.x * K4_WARPS + warp;
    const int stride = gridDim.x * K4_WARPS;
    const bool has_work = gwarp < p.n_tiles;
    if (has_work) {
        fetch_tile_desc(p, wbase, W_DESC, gwarp, lane);
        cp_async_commit();
    }
    if (smem_u32(smem) != SMEM_ORIGIN) __trap();  // lds_f32 & co. address shared memory absolutely (the host checks this too)
    if (tid == 0) mbar_init(stage_bar, 1);
    for (int i = tid; i < 256; i += K4_THREADS) smem[SM_CODE + i] = base_code(i);
    __syncthreads();
    if (USE_Z) {
        if (tid == 0) {
            mbar_expect_tx(stage_bar, Z32_BYTES);
            tma_load_1d(smem + SM_Z, p.z32, Z32_BYTES / 2, stage_bar);  // two 64 KB bulk copies
            tma_load_1d(smem + SM_Z + Z32_BYTES / 2, reinterpret_cast<const unsigned char *>(p.z32) + Z32_BYTES / 2, Z32_BYTES / 2, stage_bar);
        }
        mbar_wait(stage_bar, 0);
    }
    if (!has_work) return;
    cp_async_wait_all();
    __syncwarp();
    fetch_tile_inputs<RAND_DWELL>(p, smem, wbase, map_off, map_off + W_DESC, map_off + W_SIGOFF, gwarp, lane);
    fetch_tile_desc(p, wbase, W_DESC + 80, min(gwarp + stride, p.n_tiles - 1), lane);
    cp_async_commit();

    // ---- main loop: this warp's tiles; the inputs of tile t+1 and the descriptor of tile t+2 fly during phase B of t ----
    uint32_t slot = 0;
    for (int tile = gwarp; tile < p.n_tiles; tile += stride) {
        cp_async_wait_all();
        __syncwarp();
        const uint32_t desc_off = map_off + W_DESC + slot * 80, sigoff_off = map_off + W_SIGOFF + slot * 8;
        const TileCtx h = prepare_tile<NOISY, RAND_DWELL, METH, REV, QUAD && !METH>(p, lane, smem, par_off, map_off, desc_off, sigoff_off);
        const int next = tile + stride;
        if (next < p.n_tiles) {
            fetch_tile_inputs<RAND_DWELL>(p, smem, wbase, map_off, map_off + W_DESC + (slot ^ 1) * 80, map_off + W_SIGOFF + (slot ^ 1) * 8, next, lane);
            fetch_tile_desc(p, wbase, W_DESC + slot * 80, min(next + stride, p.n_tiles - 1), lane);
            cp_async_commit();
        }
#ifndef SQG_KO_PHASEB
        emit_tile<NOISY, RAND_DWELL, REV>(p, smem, h, lane, par_off, map_off);
#else
        if (h.S == 0x7fffffffu) emit_tile<NOISY, RAND_DWELL, REV>(p, smem, h, lane, par_off, map_off);
#endif
        __syncwarp();  // the tile buffer is rewritten by the next prepare_tile
        slot ^= 1;
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

// A few 64-bit values from device memory straight into PINNED HOST memory (unified addressing: the host pointer is valid
// on the device).  Used for the per-batch totals the host has to see before it can size buffers: a cudaMemcpyAsync of
// 8 bytes would queue on the D2H copy engine behind another slot's gigabyte of signal and stall this slot's kernels.
__global__ void publish_kernel(const int64_t *a, int na, const int64_t *b, int nb, int64_t *host_dst) {
    const int t = threadIdx.x;
    if (t < na) host_dst[t] = a[t];
    else if (t < na + nb) host_dst[t] = b[t - na];
}

// store-only kernel: the HBM write ceiling next to which the signal kernel is read
__global__ void __launch_bounds__(512) store_only_kernel(uint4 *dst, size_t n16) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    for (size_t i = t; i < n16; i += stride) __stcs(dst + i, make_uint4(t, t + 1, t + 2, (uint32_t)i));
}

}  // namespace sqg
