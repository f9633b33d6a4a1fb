// sqg_kernels.cuh — the CUDA kernels of the signal-generation path (sm_100a, hand-written).
//
// Work decomposition (DESIGN.md "Kernels"):
//   a read is 1-2 SEGMENTS (the 2nd only for the RNA stall of --prefix, src/genread.c:88-89);
//   a segment is cut into TILES of T consecutive k-mers; a tile is what one thread group turns into samples.
//
//   K0 tile_map_kernel     per segment: tile -> segment                                 -> tile_seg
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

// Everything the dwell pass and the signal kernel need to know about a tile, in one 48-byte record
// (written by K0, completed by K2) so that a producer warp reaches its bases with a single dependent load.
struct __align__(16) TileDesc {
    int64_t a_off;   // window byte i < a_rem is bases[a_off + i]   (piece a of the segment)
    int64_t b_off;   // window byte i >= a_rem is bases[b_off + i]  (piece b)
    int32_t a_rem;   // may be <= 0 or beyond the window
    int32_t nk;      // k-mers in the tile
    int32_t read;    // local read index
    uint32_t kidx0;  // dwell draw index of the tile's first k-mer (multiple of 8)
    int64_t ss_pos;  // where the tile's dwells go in ss[]
    uint32_t B;      // first logical sample of the tile within the read (filled by K2)
    uint32_t pad;
};

// Shared-memory layout of the signal kernel (byte offsets into dynamic shared memory; computed on the host by
// k4_layout() and passed in the kernel parameters, i.e. the constant bank)
struct K4Layout {
    uint32_t par, map, bmap, digit, lut, code, mbar, z16, model, total;
};

struct GenParams {
    // inputs
    const uint8_t *bases;
    const SegDesc *segs;
    const ReadDesc *reads;
    const float2 *model;  // (level_mean, level_stdv) by rank
    const __half *z16;    // Z16[65536]
    const float *z2;      // Z2[2*8192]
    // plan (written by K0-K3, read by K4)
    TileDesc *tiles;
    uint32_t *tile_sum;
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
    uint32_t rk[20];      // the ten Philox round keys (k0 + r*W0, k1 + r*W1), precomputed
    int64_t first_read;
    int32_t want_ss;
    int32_t shift_val;  // (int16)(30*digitisation/range)
    K4Layout lay;       // shared-memory layout of the signal kernel for the launch's warp count
};

constexpr int K1_THREADS = 128;  // 4 tiles per CTA, one warp each

// ------------------------------------------------------------------------------------------------
// small helpers

__device__ __forceinline__ RngKey make_key(const GenParams &p, int read_local) {
    const uint64_t r = (uint64_t)(p.first_read + read_local);
    return RngKey{p.key0, p.key1, (uint32_t)r, (uint32_t)(r >> 32)};
}

// the j-th 16-bit draw (j in 0..7) of a Philox block: even draws are bits 1..16 of word j/2, odd draws bits 1..16 of
// the same word rotated by 16.  Defined this way so that `word & 0x1FF82` is already the BYTE offset of the stratified
// table entry (bit 0 of a binary16 offset is 0, bits 2-6 are the bank): no shift, no multiply in the sample loop.
__device__ __forceinline__ uint32_t draw_word(const uint4 &w, int j) {  // j compile-time after unrolling
    const uint32_t x = (j >> 1) == 0 ? w.x : (j >> 1) == 1 ? w.y : (j >> 1) == 2 ? w.z : w.w;
    return (j & 1) ? __byte_perm(x, x, 0x1032) : x;  // rotate by 16
}
__device__ __forceinline__ uint32_t halfword(const uint4 &w, int j) { return (draw_word(w, j) >> 1) & 0xFFFFu; }
// byte offset of the stratified table entry of a draw: ((h & 0xFFC1) | bank<<1) * 2
__device__ __forceinline__ uint32_t draw_offset(uint32_t word, uint32_t bank4) { return (word & 0x1FF82u) | bank4; }

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
// K0: tile descriptors (one thread per segment)
__global__ void __launch_bounds__(256) tile_desc_kernel(const __grid_constant__ GenParams p) {
    const int si = blockIdx.x * blockDim.x + threadIdx.x;
    if (si >= p.n_segs) return;
    const SegDesc seg = p.segs[si];
    const int64_t ss0 = p.reads[seg.read].ss_off + seg.k0;
    const int nt = (seg.nk + p.T - 1) / p.T;
    for (int t = 0; t < nt; t++) {
        const int kstart = t * p.T;
        TileDesc d;
        d.a_off = seg.off_a + kstart;
        d.b_off = seg.off_b + kstart - seg.len_a;
        d.a_rem = seg.len_a - kstart;
        d.nk = min(p.T, seg.nk - kstart);
        d.read = seg.read;
        d.kidx0 = (uint32_t)(seg.k0_rng + kstart);
        d.ss_pos = ss0 + kstart;
        d.B = 0;
        d.pad = 0;
        p.tiles[seg.tile0 + t] = d;
    }
}

// K1: per-tile sum of dwells (random-dwell modes only).  One warp per tile, one lane per Philox block of 8 k-mers.
__global__ void __launch_bounds__(K1_THREADS) dwell_sum_kernel(const __grid_constant__ GenParams p) {
    const int tile = blockIdx.x * (K1_THREADS / 32) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (tile >= p.n_tiles) return;
    const int nk_tile = p.tiles[tile].nk;
    const RngKey key = make_key(p, p.tiles[tile].read);
    uint32_t sum = 0;
    if (lane * 8 < nk_tile) {
        const uint32_t blk = (p.tiles[tile].kidx0 >> 3) + lane;
        const uint4 w = philox4x32_10_rk(blk, key.r_lo, key.r_hi, ST_DWELL, p.rk);
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const float z = z16(p.z16, p.z2, stratify(halfword(w, j), blk), blk * 8 + j, key, ST_DWELL_TAIL);
            const int d = dwell_from_z(z, p.dwell_mean, p.dwell_std);
            if (lane * 8 + j < nk_tile) sum += (uint32_t)d;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (lane == 0) p.tile_sum[tile] = sum;
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
            }
            p.tiles[tile].B = (uint32_t)total;
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
            z[d] = __dadd_rn(__dmul_rn((double)za, 0.8191520442889918), __dmul_rn((double)zb, 0.573576436351046));
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
//
// Every WARP is autonomous: it owns a private tile buffer in shared memory and walks its own sequence of tiles,
// alternating a k-mer phase (A: latency-bound — descriptor, base window, model gathers, warp scan) and a sample phase
// (B: issue-bound — Philox, table normals, FFMA, 128-bit stores).  There are no CTA-wide barriers and no
// producer/consumer hand-offs after the prologue; the 16-24 resident warps of an SM are at different points of their
// tiles, so phase-A latency of some warps is covered by phase-B work of the others.  The CTA shares only read-only
// tables: the binary16 quantile table and (when it fits) the pore model, both staged by TMA bulk copies, the boundary
// LUT and the base-code table.
//   phase A, lane = 8 consecutive k-mers = one Philox dwell block: 8 dwells, warp scan, chunk->k-mer map + boundary
//            bitmap; digits -> ranks -> (mean,stdv) gathers -> (A',B') into par[]
//   phase B, lane = one 16-byte chunk (8 samples) of the emitted signal per iteration: k-mer of the first sample from
//            the map, the chunk's boundary byte -> LUT -> the 8 parameter addresses, one Philox block -> 8 table
//            normals (bank-stratified lookups) -> FFMA -> cvt.rzi -> one 128-bit store

constexpr int K4_MAX_WARPS = 24;
constexpr int K4_MAX_THREADS = K4_MAX_WARPS * 32;  // register budget: 65536 / 768 = 85
constexpr int TK = 256;      // k-mers per tile (32 lanes x 8)
constexpr int MAPC = 1344;   // 8-sample chunks per tile: >= (TK*max_dwell + 14)/8 (dna-r10: 256 k-mers x 40)
constexpr int DIG_BYTES = TK + 32;
constexpr int LUT_COPIES = 4;
constexpr int NCHUNK = 1;    // 16-byte chunks per lane per phase-B iteration (more = more ILP but more code)
constexpr int WARP_TILE_BYTES = TK * 8 + 2 * MAPC + DIG_BYTES;  // par + map + bmap + digits

// Shared-memory layout (dynamic), for nw warps:
//   [par: nw*TK float2]  (first, so that parameter addresses fit 16 bits)   [map: nw*MAPC u8] [bmap: nw*MAPC u8]
//   [digit: nw*DIG_BYTES u8] [lut: 128*LUT_COPIES uint4] [code: 256 u8] [mbar: 8 B, 16-aligned]
//   [Z16: 128 KB, 128-aligned, if USE_Z] [model: num_kmer*8 B if MODEL_SMEM]
__host__ __device__ inline K4Layout k4_layout(int nw, bool use_z, uint32_t model_bytes) {
    K4Layout L;
    uint32_t o = 0;
    L.par = o; o += (uint32_t)nw * TK * 8;
    L.map = o; o += (uint32_t)nw * MAPC;
    L.bmap = o; o += (uint32_t)nw * MAPC;
    L.digit = o; o += (uint32_t)nw * DIG_BYTES;
    o = (o + 15u) & ~15u;
    L.lut = o; o += 128 * LUT_COPIES * 16;
    L.code = o; o += 256;
    L.mbar = o; o += 16;
    o = (o + 127u) & ~127u;
    L.z16 = o; o += use_z ? (uint32_t)Z16_N * 2 : 0;
    L.model = o; o += model_bytes;
    L.total = o;
    return L;
}

// ---- raw shared-memory access by 32-bit shared address ----
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float lds_half(uint32_t addr) {
    unsigned short h;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(h) : "r"(addr));
    return __half2float(__ushort_as_half(h));
}
__device__ __forceinline__ uint32_t lds_u8(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint4 lds_u4(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ float2 lds_f2(uint32_t addr) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
    return v;
}

__device__ __forceinline__ void sts_u8(uint32_t addr, uint32_t v) { asm volatile("st.shared.u8 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ void sts_f2(uint32_t addr, float2 v) {
    asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(v.x), "f"(v.y) : "memory");
}
__device__ __forceinline__ void sts_zero16(uint32_t addr) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(addr), "r"(0u) : "memory");
}
__device__ __forceinline__ void atoms_or(uint32_t addr, uint32_t v) { asm volatile("red.shared.or.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }

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

// What phase B needs to know about the tile (registers, warp-uniform)
struct TileHdr {
    uint32_t S;      // samples in the tile
    uint32_t ph;     // chunk w covers tile samples [8w-ph, 8w-ph+8)
    uint32_t B;      // first logical sample of the tile within the read
    uint32_t L;      // samples in the read
    uint32_t r_lo, r_hi;  // global read index (Philox counter words 1,2)
    int16_t *out;    // start of the read in the signal arena
};

// ---- phase B ---------------------------------------------------------------------------------------------------

// 32-bit shared addresses of a warp's tile buffer and of the CTA's tables, computed once per kernel
struct WarpSmem {
    uint32_t par, map, bmap, dig, lut, z;
};

// NCH chunks of the same lane (w, w+32, ...) in one straight-line block so that their Philox chains and table
// lookups interleave (instruction-level parallelism: a warp alone sustains ~2x the issue rate).
template <bool NOISY, bool RAND_DWELL, bool REV, int NCH>
__device__ __forceinline__ void emit_chunks_fast(const GenParams &p, const WarpSmem &ws, const TileHdr &h, int lane,
                                                 const uint32_t (&wv)[NCH], const bool (&st)[NCH]) {
    const uint32_t zbase = ws.z;
    const uint32_t par_addr = ws.par;  // < 64 KB by layout
    uint32_t q0[NCH], pa[NCH][4], v[NCH][8];
#pragma unroll
    for (int c = 0; c < NCH; c++) {
        const uint32_t wc = wv[c];
        const int s0 = (int)(8 * wc) - (int)h.ph;  // first tile sample of the chunk; < 0 only for a clipped chunk 0
        q0[c] = REV ? (h.L - h.B - (uint32_t)(s0 + 8)) : (h.B + (uint32_t)s0);  // emitted position, multiple of 8
        uint32_t k0, bm;
        if (RAND_DWELL) {
            k0 = lds_u8(ws.map + wc);
            bm = lds_u8(ws.bmap + wc) >> 1;  // (k-mer 0's own start is never marked: it is not a boundary to cross)
        } else {
            k0 = div_sps(p, (uint32_t)max(s0, 0));
            bm = 0;
            for (int b = (int)((k0 + 1) * (uint32_t)p.sps_fixed) - s0; b < 8; b += p.sps_fixed) bm |= 1u << (b - 1);
        }
        const uint4 lu = lds_u4(ws.lut + (bm * LUT_COPIES + (lane & (LUT_COPIES - 1))) * 16);
        const uint32_t rep = (k0 * 8 + par_addr) * 0x00010001u;
        pa[c][0] = lu.x + rep; pa[c][1] = lu.y + rep; pa[c][2] = lu.z + rep; pa[c][3] = lu.w + rep;
    }
    if (NOISY) {
        uint32_t rw[NCH][4];
#pragma unroll
        for (int c = 0; c < NCH; c++) {
            const uint4 r4 = philox4x32_10_rk(q0[c] >> 3, h.r_lo, h.r_hi, ST_AMP, p.rk);
            rw[c][0] = r4.x; rw[c][1] = r4.y; rw[c][2] = r4.z; rw[c][3] = r4.w;
        }
        float zmax = 0.f;
#pragma unroll
        for (int c = 0; c < NCH; c++) {
            const uint32_t bank4 = (q0[c] >> 1) & 0x7Cu;  // stratify(): (block & 31) << 2, the chunk's Philox block picks the bank
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const int e = REV ? 7 - j : j;  // slot in the emitted chunk = which 16-bit draw
                const float2 ab = lds_f2((j & 1) ? (pa[c][j >> 1] >> 16) : (pa[c][j >> 1] & 0xFFFFu));
                const uint32_t x = (e & 1) ? __byte_perm(rw[c][e >> 1], rw[c][e >> 1], 0x1032) : rw[c][e >> 1];
                const float z = lds_half(zbase + draw_offset(x, bank4));
                zmax = fmaxf(zmax, fabsf(z));
                v[c][e] = (uint32_t)__float2int_rz(fmaf(z, ab.x, ab.y));
            }
        }
        if (__builtin_expect(zmax >= Z_TAIL_THR, 0)) {
            // rare: some draw fell into one of the 16 outermost cells -> refine it (13 more bits)
            const RngKey key{p.key0, p.key1, h.r_lo, h.r_hi};
#pragma unroll
            for (int c = 0; c < NCH; c++) {
                const uint32_t bank4 = (q0[c] >> 1) & 0x7Cu;
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const int e = REV ? 7 - j : j;
                    const uint32_t x = (e & 1) ? __byte_perm(rw[c][e >> 1], rw[c][e >> 1], 0x1032) : rw[c][e >> 1];
                    const uint32_t hw = draw_offset(x, bank4) >> 1;
                    if ((hw & 0x7FFFu) >= Z_TAIL_FIRST) {
                        const float2 ab = lds_f2((j & 1) ? (pa[c][j >> 1] >> 16) : (pa[c][j >> 1] & 0xFFFFu));
                        const float z = z16_tail(p.z2, hw, q0[c] + e, key, ST_AMP_TAIL);
                        v[c][e] = (uint32_t)__float2int_rz(fmaf(z, ab.x, ab.y));
                    }
                }
            }
        }
    } else {
#pragma unroll
        for (int c = 0; c < NCH; c++)
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const int e = REV ? 7 - j : j;
                const float2 ab = lds_f2((j & 1) ? (pa[c][j >> 1] >> 16) : (pa[c][j >> 1] & 0xFFFFu));
                v[c][e] = __float_as_uint(ab.y);
            }
    }
#pragma unroll
    for (int c = 0; c < NCH; c++) {
        // low 16 bits of each int32 (the reference's wrap, src/gensig.c:270), packed little-endian
        const uint4 pk = make_uint4(__byte_perm(v[c][0], v[c][1], 0x5410), __byte_perm(v[c][2], v[c][3], 0x5410),
                                    __byte_perm(v[c][4], v[c][5], 0x5410), __byte_perm(v[c][6], v[c][7], 0x5410));
        if (st[c]) {
            const int s0 = (int)(8 * wv[c]) - (int)h.ph;
            if (s0 >= 0 && (uint32_t)(s0 + 8) <= h.S) {
                __stcs(reinterpret_cast<uint4 *>(h.out + q0[c]), pk);
            } else {
                // clipped chunk at an end of the tile (at most two per tile): store only the tile's own samples; the
                // neighbouring tile computes the same Philox block and stores the rest
                const uint32_t pw[4] = {pk.x, pk.y, pk.z, pk.w};
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const int e = REV ? 7 - j : j;
                    // (uint32_t)(s0 + j) < S also rejects the negative positions of a clipped first chunk
                    if ((uint32_t)(s0 + j) < h.S) h.out[q0[c] + e] = (int16_t)((e & 1) ? (pw[e >> 1] >> 16) : pw[e >> 1]);
                }
            }
        }
    }
}

template <bool NOISY, bool RAND_DWELL, bool REV>
__device__ __forceinline__ void emit_tile(const GenParams &p, const WarpSmem &ws, const TileHdr h, int lane) {
    const uint32_t nW = (h.S + h.ph + 7) >> 3;  // chunks touched by the tile (the first and last may be clipped)
    for (uint32_t w = lane; w < nW; w += 32 * NCHUNK) {
        uint32_t wv[NCHUNK];
        bool st[NCHUNK];
#pragma unroll
        for (int c = 0; c < NCHUNK; c++) {
            st[c] = w + 32u * c < nW;
            wv[c] = st[c] ? w + 32u * c : w;  // a missing partner is computed redundantly and not stored
        }
        emit_chunks_fast<NOISY, RAND_DWELL, REV, NCHUNK>(p, ws, h, lane, wv, st);
    }
}

// ---- phase A ---------------------------------------------------------------------------------------------------

__device__ __forceinline__ TileDesc load_tile_desc(const TileDesc *__restrict__ tiles, int tile) {
    const uint4 *q = reinterpret_cast<const uint4 *>(tiles + tile);
    const uint4 a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2);
    TileDesc d;
    d.a_off = (int64_t)(((uint64_t)a.y << 32) | a.x);
    d.b_off = (int64_t)(((uint64_t)a.w << 32) | a.z);
    d.a_rem = (int32_t)b.x; d.nk = (int32_t)b.y; d.read = (int32_t)b.z; d.kidx0 = b.w;
    d.ss_pos = (int64_t)(((uint64_t)c.y << 32) | c.x);
    d.B = c.z; d.pad = 0;
    return d;
}

constexpr int WIN_LOADS = (TK + 8 + 31) / 32;  // byte loads per lane for a tile's base window (k <= 9)

// Source order = latency order: global loads first (base window, per-read values), then the dwell draws (shared
// memory only) while they fly, then digits -> ranks -> model gathers, then the map/bitmap scatter while the gathers fly.
template <bool NOISY, bool RAND_DWELL, bool METH, bool REV, bool MODEL_SMEM>
__device__ __forceinline__ TileHdr prepare_tile(const GenParams &p, const TileDesc td, int lane, const WarpSmem &ws,
                                                const uint8_t *dig, const uint8_t *code, const float2 *__restrict__ model) {
    const int nk_tile = td.nk;
    // (1) the tile's base window: coalesced byte loads, all in flight at once
    uint32_t raw[WIN_LOADS];
    const int nb = nk_tile + p.k - 1;
    const uint8_t *pa_lane = p.bases + td.a_off + lane, *pb_lane = p.bases + td.b_off + lane;
    if (td.a_rem >= nb || td.a_rem <= 0) {  // the whole window lies in one piece (always, except around a prefix junction)
        const uint8_t *src = td.a_rem > 0 ? pa_lane : pb_lane;
#pragma unroll
        for (int u = 0; u < WIN_LOADS; u++) raw[u] = (lane + 32 * u < nb) ? __ldg(src + 32 * u) : 0u;
    } else {
#pragma unroll
        for (int u = 0; u < WIN_LOADS; u++) {
            const int i = lane + 32 * u;
            raw[u] = (i < nb) ? __ldg((i < td.a_rem ? pa_lane : pb_lane) + 32 * u) : 0u;
        }
    }
    // (2) per-read values
    const uint32_t L = __ldg(p.read_siglen + td.read);
    const double offset = __ldg(p.read_offset + td.read);
    const int64_t sigoff = __ldg(p.read_sigoff + td.read);
    const RngKey key = make_key(p, td.read);
    const uint32_t B = td.B;
    const uint32_t ph = REV ? ((B - L) & 7u) : (B & 7u);
    const int m0 = lane * 8;
    const bool active = m0 < nk_tile;

    // (3) dwells of this lane's 8 k-mers = one Philox block (src/gensig.c:255-256), then the warp scan
    int d[8];
    uint32_t o = 0, S = (uint32_t)nk_tile * (uint32_t)p.sps_fixed;
    if (RAND_DWELL) {
        const uint32_t blk = (td.kidx0 >> 3) + lane;
        const uint4 w = philox4x32_10_rk(blk, key.r_lo, key.r_hi, ST_DWELL, p.rk);
        const uint32_t zbase = ws.z;
        float zmax = 0.f;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const float z = lds_half(zbase + draw_offset(draw_word(w, j), (blk & 31u) << 2));
            zmax = fmaxf(zmax, fabsf(z));
            d[j] = dwell_from_z(z, p.dwell_mean, p.dwell_std);
        }
        if (__builtin_expect(zmax >= Z_TAIL_THR, 0)) {
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const uint32_t hw = stratify(halfword(w, j), blk);
                if ((hw & 0x7FFFu) >= Z_TAIL_FIRST)
                    d[j] = dwell_from_z(z16_tail(p.z2, hw, blk * 8 + j, key, ST_DWELL_TAIL), p.dwell_mean, p.dwell_std);
            }
        }
        if (nk_tile < TK) {  // only the last tile of a segment has lanes beyond its end
#pragma unroll
            for (int j = 0; j < 8; j++)
                if (m0 + j >= nk_tile) d[j] = 0;
        }
        uint32_t local = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) local += (uint32_t)d[j];
        uint32_t inc = local;
#pragma unroll
        for (int sh = 1; sh < 32; sh <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, inc, sh);
            if (lane >= sh) inc += v;
        }
        S = __shfl_sync(0xffffffffu, inc, 31);
        o = inc - local;
    } else {
#pragma unroll
        for (int j = 0; j < 8; j++) d[j] = (m0 + j < nk_tile) ? p.sps_fixed : 0;
    }

    // (4) bases -> digits (the code table folds IUPAC letters, src/seq.h:14-28 / :45-60)
#pragma unroll
    for (int u = 0; u < WIN_LOADS; u++) {
        const int i = lane + 32 * u;
        const uint32_t c = code[raw[u] & 0xFFu];
        if (i < nb) sts_u8(ws.dig + i, METH ? (c >> 4) : (c & 3u));
    }
    __syncwarp();

    // (5) ranks of this lane's 8 k-mers (src/seq.h:31-42 / :62-74).  They are visited in ROTATED order
    // jj(j) = (j + lane/2) & 7 so that the 8-byte parameter stores of a half-warp fall into 16 different bank pairs.
    const int rot = METH ? 0 : (lane >> 1);
    uint32_t ranks[8];
    {
        const uint2 dwa = *reinterpret_cast<const uint2 *>(dig + m0);
        const uint2 dwb = *reinterpret_cast<const uint2 *>(dig + m0 + 8);
        if (!METH) {
            // 16 two-bit digits packed first-digit-most-significant: ((w & 0x03030303) * 0x40100401) >> 24 packs 4 bytes
            const uint32_t P = ((((dwa.x & 0x03030303u) * 0x40100401u) >> 24) << 24) | ((((dwa.y & 0x03030303u) * 0x40100401u) >> 24) << 16) |
                               ((((dwb.x & 0x03030303u) * 0x40100401u) >> 24) << 8) | (((dwb.y & 0x03030303u) * 0x40100401u) >> 24);
            const int sh0 = 32 - 2 * p.k;
#pragma unroll
            for (int j = 0; j < 8; j++) ranks[j] = (P >> (sh0 - 2 * ((j + rot) & 7))) & p.kmask;
        } else {
            const uint32_t dw[4] = {dwa.x, dwa.y, dwb.x, dwb.y};
            const int km1 = p.k - 1;
            uint32_t rank = 0;
#pragma unroll
            for (int i = 0; i < 8; i++)
                if (i < km1) rank = rank * 5 + ((dw[i >> 2] >> (8 * (i & 3))) & 0xFFu);
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const int bi = km1 + j;
                const uint32_t word = bi < 4 ? dw[0] : bi < 8 ? dw[1] : bi < 12 ? dw[2] : dw[3];
                rank = (rank % p.kmask) * 5 + ((word >> (8 * (bi & 3))) & 0xFFu);
                ranks[j] = rank;
            }
        }
        if (nk_tile < TK) {
#pragma unroll
            for (int j = 0; j < 8; j++)
                if (m0 + ((j + rot) & 7) >= nk_tile) ranks[j] = 0;
        }
    }
    float2 mv[8];
#pragma unroll
    for (int j = 0; j < 8; j++) mv[j] = MODEL_SMEM ? model[ranks[j]] : __ldg(&model[ranks[j]]);

    // (6) chunk -> k-mer map and boundary bitmap (needs only the dwells: runs while the gathers are in flight)
    if (RAND_DWELL) {
        static_assert(MAPC / 16 <= 96, "three 16-byte stores per lane must cover the boundary bitmap");
#pragma unroll
        for (int u = 0; u < 3; u++)
            if (lane + 32 * u < MAPC / 16) sts_zero16(ws.bmap + 16 * (lane + 32 * u));
        __syncwarp();
        if (active) {
            uint32_t pos = o + ph;  // the k-mer starts at bit (pos&7) of chunk (pos>>3)
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const int m = m0 + j;
                if (nk_tile == TK || m < nk_tile) {
                    // chunks whose first (clipped) sample falls inside this k-mer get it as their first k-mer
                    uint32_t w0 = (j == 0 && m0 == 0) ? 0u : (pos + 7) >> 3;
                    const uint32_t end = pos + (uint32_t)d[j];
                    const uint32_t w1 = (end + 7) >> 3;
                    const uint32_t a = ws.map + w0;
                    if (w0 < w1) sts_u8(a, (uint32_t)m);
                    if (w0 + 1 < w1) sts_u8(a + 1, (uint32_t)m);
                    if (w0 + 2 < w1) sts_u8(a + 2, (uint32_t)m);
                    for (w0 += 3; w0 < w1; w0++) sts_u8(ws.map + w0, (uint32_t)m);
                    if (!(j == 0 && m0 == 0)) atoms_or(ws.bmap + ((pos >> 3) & ~3u), 1u << (pos & 31));
                    pos = end;
                }
            }
        }
    }
    if (p.want_ss && active) {
#pragma unroll
        for (int j = 0; j < 8; j++)
            if (m0 + j < nk_tile) p.ss[td.ss_pos + m0 + j] = d[j];
    }

    // (7) per-k-mer parameters from the gathered (level_mean, level_stdv)
    if (active) {
        const float scale_f = (float)p.scale, off_f = (float)offset;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            float2 pr;
            if (NOISY) {
                // single precision, rounded once each: A' = (stdv*amp_noise)*scale, B' = fma(mean, scale, -offset)
                pr = make_float2(__fmul_rn(__fmul_rn(mv[j].y, p.amp_noise), scale_f), fmaf(mv[j].x, scale_f, -off_f));
            } else {
                // src/gensig.c:266,270: (double)level_mean*digitisation/range - offset, truncated
                const double v = __dsub_rn(__ddiv_rn(__dmul_rn((double)mv[j].x, p.digitisation), p.range), offset);
                pr = make_float2(0.f, __uint_as_float(to_i16_bits(v)));
            }
            sts_f2(ws.par + 8 * (m0 + ((j + rot) & 7)), pr);  // rows beyond the tile are never read
        }
    }
    __syncwarp();
    TileHdr h;
    h.S = S; h.ph = ph; h.B = B; h.L = L; h.r_lo = key.r_lo; h.r_hi = key.r_hi;
    h.out = p.sig + sigoff;
    return h;
}

template <bool NOISY, bool RAND_DWELL, bool METH, bool REV, bool MODEL_SMEM>
__global__ void __launch_bounds__(K4_MAX_THREADS, 1) signal_kernel(const __grid_constant__ GenParams p) {
    constexpr bool USE_Z = NOISY || RAND_DWELL;
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x;
    const int nw = blockDim.x >> 5;
    const int warp = tid >> 5, lane = tid & 31;
    const K4Layout &lay = p.lay;
    uint4 *lut = reinterpret_cast<uint4 *>(smem + lay.lut);
    uint8_t *code = smem + lay.code;
    unsigned long long *stage_bar = reinterpret_cast<unsigned long long *>(smem + lay.mbar);
    __half *z16s = reinterpret_cast<__half *>(smem + lay.z16);
    float2 *models = reinterpret_cast<float2 *>(smem + lay.model);

    // ---- prologue: tables ----
    if (tid == 0) mbar_init(stage_bar, 1);
    for (int i = tid; i < 256; i += blockDim.x) code[i] = base_code(i);
    for (int i = tid; i < 128 * LUT_COPIES; i += blockDim.x) {
        const int m = (i / LUT_COPIES) << 1;  // boundary mask (bit 0 is never used)
        uint32_t f[8], cnt = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            if (j >= 1 && ((m >> j) & 1)) cnt++;
            f[j] = cnt * 8;
        }
        lut[i] = make_uint4(f[0] | (f[1] << 16), f[2] | (f[3] << 16), f[4] | (f[5] << 16), f[6] | (f[7] << 16));
    }
    __syncthreads();
    if (tid == 0) {
        uint32_t bytes = 0;
        if (USE_Z) bytes += Z16_N * 2;
        if (MODEL_SMEM) bytes += p.num_kmer * 8;
        if (bytes) {
            mbar_expect_tx(stage_bar, bytes);
            if (USE_Z) {
                tma_load_1d(z16s, p.z16, Z16_N, stage_bar);  // two 64 KB bulk copies
                tma_load_1d(z16s + Z16_N / 2, p.z16 + Z16_N / 2, Z16_N, stage_bar);
            }
            if (MODEL_SMEM) tma_load_1d(models, p.model, p.num_kmer * 8, stage_bar);
        }
    }
    if (USE_Z || MODEL_SMEM) mbar_wait(stage_bar, 0);
    float2 *par = reinterpret_cast<float2 *>(smem + lay.par) + warp * TK;
    uint8_t *map = smem + lay.map + warp * MAPC;
    uint8_t *bmap = smem + lay.bmap + warp * MAPC;
    uint8_t *dig = smem + lay.digit + warp * DIG_BYTES;
    if (smem_u32(par) + TK * 8 + 64 >= 0x10000u) __trap();
    const float2 *__restrict__ model = MODEL_SMEM ? models : p.model;
    const WarpSmem ws{smem_u32(par), smem_u32(map), smem_u32(bmap), smem_u32(dig), smem_u32(lut), smem_u32(z16s)};

    // ---- main loop: this warp's tiles ----
    const int gwarp = blockIdx.x * nw + warp;
    const int stride = gridDim.x * nw;
    if (gwarp >= p.n_tiles) return;
    TileDesc td_next = load_tile_desc(p.tiles, gwarp);
    for (int tile = gwarp; tile < p.n_tiles; tile += stride) {
        const TileDesc td = td_next;
        td_next = load_tile_desc(p.tiles, min(tile + stride, p.n_tiles - 1));  // in flight during this tile
        const TileHdr h = prepare_tile<NOISY, RAND_DWELL, METH, REV, MODEL_SMEM>(p, td, lane, ws, dig, code, model);
        emit_tile<NOISY, RAND_DWELL, REV>(p, ws, h, lane);
        __syncwarp();  // the tile buffer is rewritten by the next prepare_tile
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
