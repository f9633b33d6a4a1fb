// sqg_kernels.cuh — the CUDA kernels of the signal-generation path (sm_100a, hand-written).
//
// Work decomposition (DESIGN.md "Kernels"):
//   a read is 1-2 SEGMENTS (the 2nd only for the RNA stall of --prefix, src/genread.c:88-89);
//   a segment is cut into TILES of T consecutive k-mers; a tile is what one thread group turns into samples.
//
//   K0 tile_desc_kernel    per segment (warp): one 48-byte descriptor per tile                   -> tiles
//   K1 dwell_kernel        per tile (warp)  : draw the T dwells (Philox + table normals)         -> kpos (u16 prefix of the
//                                                                                                   dwells within the tile), tile_sum, ss
//   K2 read_plan_kernel    per read (warp)  : exclusive scan of its tile sums, per-read draws    -> tiles.B/S, siglen, offset, median_before
//   K3 read_offsets_kernel one CTA          : exclusive scan of the 64-sample-aligned lengths    -> sigoff, totals
//   K4 signal_kernel       (sqg_signal.cuh) : every warp walks a contiguous range of tiles: encode k-mers, gather their
//                          parameters, mark k-mer starts, emit every int16 sample with 128-bit stores -> signal  (the hot kernel)
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

// Everything the dwell pass and the signal kernel need to know about a tile, in one 48-byte record (written by K0,
// completed by K2) so that a warp reaches its bases, its dwells and its place in the read with independent loads.
enum : uint16_t { TILE_READ_FIRST = 1, TILE_READ_LAST = 2 };
struct __align__(16) TileDesc {
    int64_t a_off;   // window byte i < a_rem is bases[a_off + i]   (piece a of the segment)
    int64_t b_off;   // window byte i >= a_rem is bases[b_off + i]  (piece b)
    int32_t a_rem;   // may be <= 0 or beyond the window
    uint16_t nk;     // k-mers in the tile
    uint16_t flags;  // TILE_READ_FIRST / TILE_READ_LAST: first / last tile of its read
    int32_t read;    // local read index
    uint32_t kidx0;  // dwell draw index of the tile's first k-mer (multiple of 8)
    int64_t ss_pos;  // where the tile's dwells go in ss[]
    uint32_t B;      // first sample of the tile within the read, in generation order (K2)
    uint32_t S;      // samples in the tile (K2)
};
static_assert(sizeof(TileDesc) == 48, "TileDesc is loaded as three 16-byte words");

// what the signal kernel needs to know about a read, in one 32-byte record (K3)
struct __align__(16) ReadRec {
    int64_t sigoff;   // start of the read in the signal arena
    uint32_t L;       // samples in the read
    uint32_t pad0;
    double offset;    // ADC offset
    uint64_t pad1;
};
static_assert(sizeof(ReadRec) == 32, "ReadRec is fetched as two 16-byte words");

struct GenParams {
    // inputs
    const uint8_t *bases;
    const SegDesc *segs;
    const ReadDesc *reads;
    const float2 *model;     // (level_mean, level_stdv) by rank
    const float2 *model_am;  // (A', M) by rank: A' = (level_stdv*amp_noise)*scale, M = level_mean*scale, each rounded once (binary32)
    const float4 *pair_model;  // by (k+1)-mer rank: (A', M) of its two k-mers (first k bases, last k bases); base-4 models only
    const float *z32;     // Z32[32768]
    const float *z2;      // Z2[8192]
    // plan (written by K0-K3, read by K4)
    TileDesc *tiles;
    uint32_t *tile_sum;
    uint4 *kpos;          // per tile: exclusive prefix of its TK dwells as uint16 (32 x uint4), written by K1 (random-dwell modes)
    uint32_t *read_siglen;
    uint32_t *read_n0;
    int64_t *read_sigoff;
    ReadRec *read_rec;
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
    uint32_t pow5k;    // base 5: 5^k
    int32_t par_cap;      // k-mers the signal kernel's window may hold (<= PAR_N; fixed-dwell modes: bounded by the exact division)
    int32_t l2_vote;      // signal kernel: skip a chunk's third k-mer level when no lane of the warp has one (pays with long dwells)
    uint32_t tile_s_cap;  // a tile with more samples than this goes through the signal kernel's slow path (statistically unreachable)
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

constexpr int TK = 256;          // k-mers per tile at most (32 lanes x 8); also the row length of GenParams::kpos

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
// init: the pore model in the form the sample arithmetic uses.  A sample is trunc(fma.rz(z, A', B)) with
//   A' = (level_stdv * amp_noise) * scale      (two binary32 products, each rounded once; src/sim.c:249 makes the first)
//   B  = M + c_r,   M = level_mean * scale     (binary32), c_r = 32768 - (float)offset of the read
// A' and M depend on the context only, so they are tabulated once: the signal kernel's per-k-mer work is one FADD.
__global__ void __launch_bounds__(256) model_am_kernel(const float2 *__restrict__ model, float2 *__restrict__ am, uint32_t n, float amp_noise, float scale_f) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float2 m = model[i];
    am[i] = make_float2(__fmul_rn(__fmul_rn(m.y, amp_noise), scale_f), __fmul_rn(m.x, scale_f));
}

// The same table indexed by (k+1)-mer.  Two consecutive k-mers of a read overlap in k-1 bases, so one 16-byte
// entry addressed by the (k+1)-mer they span holds the parameters of both: the signal kernel's model gathers - one
// 32-byte sector request each, the scarcest resource of its k-mer phase - are halved.  4^(k+1) x 16 B (16 MB for 9-mers).
__global__ void __launch_bounds__(256) pair_model_kernel(const float2 *__restrict__ am, float4 *__restrict__ pair, uint32_t n_pair, uint32_t kmask) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pair) return;
    const float2 a = am[i >> 2], b = am[i & kmask];
    pair[i] = make_float4(a.x, a.y, b.x, b.y);
}
// base-5 (CpG) models: the (k+1)-mer of digits d0..dk has rank i = sum d_j 5^(k-j); its k-mers are i / 5 and i mod 5^k.
// 5^10 x 16 B = 156 MB for 9-mers, of which reads touch only the entries consistent with "M stands before G" (~25 MB).
__global__ void __launch_bounds__(256) pair_model5_kernel(const float2 *__restrict__ am, float4 *__restrict__ pair, uint32_t n_pair, uint32_t pow5k) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pair) return;
    const float2 a = am[i / 5u], b = am[i % pow5k];
    pair[i] = make_float4(a.x, a.y, b.x, b.y);
}

// ------------------------------------------------------------------------------------------------
// K0: tile descriptors (one warp per segment, one lane per tile)
__global__ void __launch_bounds__(256) tile_desc_kernel(const __grid_constant__ GenParams p) {
    const int si = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (si >= p.n_segs) return;
    const SegDesc seg = p.segs[si];
    const ReadDesc rd = p.reads[seg.read];
    const int64_t ss0 = rd.ss_off + seg.k0;
    const bool seg_first = si == rd.seg0, seg_last = si == rd.seg0 + rd.nseg - 1;
    const int nt = (seg.nk + p.T - 1) / p.T;
    for (int t = lane; t < nt; t += 32) {
        const int kstart = t * p.T;
        TileDesc d;
        d.a_off = seg.off_a + kstart;
        d.b_off = seg.off_b + kstart - seg.len_a;
        d.a_rem = seg.len_a - kstart;
        d.nk = (uint16_t)min(p.T, seg.nk - kstart);
        d.flags = (uint16_t)((seg_first && t == 0 ? TILE_READ_FIRST : 0) | (seg_last && t == nt - 1 ? TILE_READ_LAST : 0));
        d.read = seg.read;
        d.kidx0 = (uint32_t)(seg.k0_rng + kstart);
        d.ss_pos = ss0 + kstart;
        d.B = 0; d.S = 0;
        p.tiles[seg.tile0 + t] = d;
    }
}

// K1: the dwells (random-dwell modes only; src/gensig.c:255-256).  Persistent CTAs (one per SM, 32 warps) with the
// quantile table staged in shared memory by TMA; a warp takes a tile at a time, one lane per Philox block of 8 k-mers:
// 8 table normals -> 8 dwells -> their exclusive prefix WITHIN THE TILE (a warp scan), stored as one 16-byte row piece of
// uint16 - the signal kernel adds the tile's own start and has every k-mer's first sample without scanning anything -
// the tile's sample count, and - when asked for - the reference's aln->ss (src/gensig.c:273-281).
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
    __syncthreads();
    if (threadIdx.x == 0) {   // tail cells (NaN in the table file): +-infinity in this kernel's copy, see the dwell loop
        reinterpret_cast<float *>(smem1)[Z_TAIL_IDX] = __int_as_float(0x7F800000);
        reinterpret_cast<float *>(smem1)[Z_TAIL_IDX | 0x4000u] = __int_as_float(0xFF800000);
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int nw = K1_THREADS / 32;
    const int tile0 = blockIdx.x * nw + (threadIdx.x >> 5), tstride = gridDim.x * nw;
    if (tile0 >= p.n_tiles) return;
    // the descriptor words of the NEXT tile are loaded one iteration ahead (the kernel is otherwise bound by this latency)
    const uint4 *q0 = reinterpret_cast<const uint4 *>(p.tiles + tile0);
    uint4 nb = __ldg(q0 + 1), nc = __ldg(q0 + 2);
    const uint32_t r0_lo = (uint32_t)(uint64_t)p.first_read, r0_hi = (uint32_t)((uint64_t)p.first_read >> 32);
    for (int tile = tile0; tile < p.n_tiles; tile += tstride) {
        const uint4 b = nb, c = nc;
        {
            const uint4 *qn = reinterpret_cast<const uint4 *>(p.tiles + min(tile + tstride, p.n_tiles - 1));
            nb = __ldg(qn + 1); nc = __ldg(qn + 2);
        }
        const int nk_tile = (int)(b.y & 0xFFFFu);
        const uint32_t kidx0 = b.w;
        const int64_t ss_pos = (int64_t)(((uint64_t)c.y << 32) | c.x);
        const uint64_t rg = (((uint64_t)r0_hi << 32) | r0_lo) + (uint64_t)(int64_t)(int32_t)b.z;
        const RngKey key{p.key0, p.key1, (uint32_t)rg, (uint32_t)(rg >> 32)};
        uint32_t sum = 0;
        uint32_t d[8];
#pragma unroll
        for (int j = 0; j < 8; j++) d[j] = 0;
        if (lane * 8 < nk_tile) {
            const uint32_t blk = (kidx0 >> 3) + lane;
            const uint4 w = philox4x32_rk(blk, key.r_lo, key.r_hi, ST_DWELL, p.rk);
            const uint32_t class4 = (blk & 31u) << 2;
            // The shared copy of the table holds +-infinity in the two tail cells (patched in the prologue): their dwell
            // comes out of the conversion as INT_MAX / INT_MIN, so ONE test over the lane's eight dwells finds the draws
            // that need the tail refinement.  A dwell below 1 is folded: d -> 1 - d, i.e. max(d, 1 - d).
            uint32_t all = 0;
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const uint32_t off = z_offset(draw_word(w, j), class4);
                const int dd = __float2int_rn(fmaf(*reinterpret_cast<const float *>(smem1 + off), p.dwell_std, p.dwell_mean));
                d[j] = (uint32_t)max(dd, 1 - dd);
                all |= d[j];
            }
            if (__builtin_expect(all >= 0x10000u, 0)) {   // (every real dwell is below 2^14)
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const uint32_t off = z_offset(draw_word(w, j), class4);
                    if (z_is_tail(off)) d[j] = (uint32_t)dwell_from_z(z_tail(p.z2, off, blk * 8 + j, key, ST_DWELL_TAIL), p.dwell_mean, p.dwell_std);
                }
            }
            if (nk_tile < TK) {
#pragma unroll
                for (int j = 0; j < 8; j++)
                    if (lane * 8 + j >= nk_tile) d[j] = 0;
            }
            if (p.want_ss) {
                const int64_t ss0 = ss_pos + lane * 8;
#pragma unroll
                for (int j = 0; j < 8; j++)
                    if (lane * 8 + j < nk_tile) p.ss[ss0 + j] = (int32_t)d[j];
            }
#pragma unroll
            for (int j = 0; j < 8; j++) sum += d[j];
        }
        uint32_t inc = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += v;
        }
        uint32_t e[8];   // exclusive prefix (every tile has fewer than 2^16 samples: ctx_setup bounds T * max dwell)
        e[0] = inc - sum;
#pragma unroll
        for (int j = 1; j < 8; j++) e[j] = e[j - 1] + d[j - 1];
        p.kpos[(size_t)tile * (TK / 8) + lane] = make_uint4(e[0] | (e[1] << 16), e[2] | (e[3] << 16), e[4] | (e[5] << 16), e[6] | (e[7] << 16));
        if (lane == 31) p.tile_sum[tile] = inc;
    }
}

// fixed-dwell modes with aln->ss requested: every k-mer has sps_fixed samples
__global__ void __launch_bounds__(256) fixed_ss_kernel(const __grid_constant__ GenParams p, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p.ss[i] = p.sps_fixed;
}

// ------------------------------------------------------------------------------------------------
// K2: per read (one warp, one lane per tile) — scan its tiles, draw offset / median_before (src/gensig.c:312-318),
// complete the tile descriptors (B, S)
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
        ReadRec *rr = p.read_rec + r;   // (sigoff: K3)
        rr->L = (uint32_t)total;
        rr->pad0 = 0;
        rr->offset = off;
        rr->pad1 = 0;
    }
}

// ------------------------------------------------------------------------------------------------
// K3: one CTA — exclusive scan of the 64-sample-aligned read lengths, 4096 reads per round (four consecutive reads per
// thread: coalesced 16-byte loads)
__global__ void __launch_bounds__(1024) read_offsets_kernel(const __grid_constant__ GenParams p) {
    __shared__ uint64_t s_warp[32];
    __shared__ uint64_t s_raw[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint64_t carry = 0, raw_total = 0;
    // the lengths of round i+1 are requested before round i is scanned (the kernel is one CTA: nothing else hides the latency)
    auto load4 = [&](int r) -> uint4 {
        uint4 v = make_uint4(0, 0, 0, 0);
        if (r + 3 < p.n_reads) {
            v = *reinterpret_cast<const uint4 *>(p.read_siglen + r);   // (cudaMalloc'd, r a multiple of 4)
        } else {
            if (r < p.n_reads) v.x = p.read_siglen[r];
            if (r + 1 < p.n_reads) v.y = p.read_siglen[r + 1];
            if (r + 2 < p.n_reads) v.z = p.read_siglen[r + 2];
        }
        return v;
    };
    uint4 vn = load4(4 * tid);
    for (int r0 = 0; r0 < p.n_reads; r0 += 4096) {
        const int r = r0 + 4 * tid;
        const uint4 vc = vn;
        vn = load4(r + 4096);
        const uint32_t l[4] = {vc.x, vc.y, vc.z, vc.w};
        uint64_t al[4], mine = 0, raw = 0;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            al[i] = ((uint64_t)l[i] + 63) & ~63ull;
            mine += al[i];
            raw += l[i];
        }
        uint64_t inc = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint64_t v = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += v;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) raw += __shfl_xor_sync(0xffffffffu, raw, o);
        __syncthreads();   // (the shared totals of the previous round have been read)
        if (lane == 31) s_warp[warp] = inc;
        if (lane == 0) s_raw[warp] = raw;
        __syncthreads();
        // every warp scans the 32 warp totals for itself
        const uint64_t wt = s_warp[lane];
        uint64_t winc = wt, rawall = s_raw[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint64_t v = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += v;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) rawall += __shfl_xor_sync(0xffffffffu, rawall, o);
        const uint64_t before = __shfl_sync(0xffffffffu, winc - wt, warp), all = __shfl_sync(0xffffffffu, winc, 31);
        uint64_t off = carry + before + inc - mine;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            if (r + i < p.n_reads) {
                p.read_sigoff[r + i] = (int64_t)off;
                p.read_rec[r + i].sigoff = (int64_t)off;   // (L and the ADC offset of the record: K2)
            }
            off += al[i];
        }
        carry += all;
        raw_total += rawall;
    }
    if (tid == 0) {
        p.meta[0] = (int64_t)carry;
        p.meta[1] = (int64_t)raw_total;
    }
}

// ------------------------------------------------------------------------------------------------
// shared-memory / asynchronous-copy helpers used by K1 and K4

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
