// sqg_signal.cuh — K4, the signal kernel (sm_100a, hand-written): src/gensig.c:226-288 for a whole batch.
//
// Work decomposition.  The batch's TILES (<= 256 consecutive k-mers of one read segment, sqg_kernels.cuh) form one
// global sequence; every WARP takes a contiguous range of it and walks it autonomously (no CTA barrier after the
// prologue).  Within a read the warp keeps a SLIDING WINDOW in shared memory:
//   par[]  (A', B) of the registered k-mers                       (B = M + c_r: one FADD per k-mer, see model_am_kernel)
//   map[]  per 32 samples {bit s: a k-mer starts at sample s, index of the k-mer that owns the entry's first sample}
// phase A registers the next tile's k-mers (digits -> ranks -> table gathers; start bits from the tile-relative prefix the
// dwell kernel stored), phase B emits every GROUP of 256 samples (32 lanes x one 16-byte chunk) that has become complete,
// then the k-mers still needed (those of the incomplete group) are moved to the front of the window.  Groups are aligned
// in the EMITTED signal, so:
//   * every store of the bulk is a full, aligned 16-byte chunk - nothing is clipped at tile boundaries;
//   * lane = emitted chunk & 31, so a warp-wide table lookup touches 32 different banks whatever bijection of the lanes
//     the group's class rotation applies;
//   * only where a warp's range begins or ends inside a read is a chunk shared with another warp: those (<= 2 per range)
//     take the exact path, which stores sample by sample.
// Everything a tile needs from global memory (descriptor, base window, prefix row, its read's offset/length/arena
// position) is fetched one tile ahead with cp.async (LDGSTS) while the previous tile's samples are being emitted.
//
// Shared memory (byte offsets into the dynamic array, all compile-time so that they fold into LDS/STS immediates):
//   [Z32: 128 KB quantile table (TMA bulk copy)] [code: 256 B] [mbar] [per warp: map, par, raw window, digits, prefix, descriptors]
#pragma once
#include <type_traits>

#include "sqg_kernels.cuh"

namespace sqg {

#ifndef SQG_K4_WARPS
#define SQG_K4_WARPS 16
#endif
constexpr int K4_WARPS = SQG_K4_WARPS;
constexpr int K4_THREADS = K4_WARPS * 32;  // register budget: 65536 / 512 = 128
constexpr uint32_t GROUP_S = 256;          // samples per group: 32 lanes x 8
constexpr uint32_t UNIT_G = 3;             // groups per unit: 96 chunks share 64 Philox blocks (12 draws each)
constexpr uint32_t UNIT_C = 32 * UNIT_G, UNIT_S = GROUP_S * UNIT_G;
#ifndef SQG_PAR_TAIL
#define SQG_PAR_TAIL 128
#endif
#ifndef SQG_MAP_ENT
#define SQG_MAP_ENT 168
#endif
constexpr int PAR_TAIL = SQG_PAR_TAIL;     // k-mers carried from tile to tile at most (else the run is cut)
constexpr int PAR_N = TK + PAR_TAIL;       // registered k-mers at most
constexpr int MAP_ENT = SQG_MAP_ENT;       // map entries (32 samples each): 168 = a window of 5376 samples
constexpr uint32_t TILE_S_CAP = (MAP_ENT - 8) * 32 - UNIT_S - 32;   // samples of one tile the window is guaranteed to hold
constexpr int DIG_BYTES = TK + 32;         // digits of the tile's base window; the same size holds the raw window (16-byte granules)
// per-warp buffer
constexpr uint32_t W_MAP = 0;                               // MAP_ENT (+2 that the one-ahead loads may touch) x {bits, base}
constexpr uint32_t W_PARG = W_MAP + (MAP_ENT + 2) * 8;      // guard entry par[-1] (the padding chunk of a reversed read)
constexpr uint32_t W_PAR = W_PARG + 16;                     // par[0 .. PAR_N) + 4 rows the sample loop may load past the end (k0 + 3 <= last + 3)
constexpr uint32_t W_RAW = W_PAR + (PAR_N + 4) * 8;         // prefetched base window (ASCII), 16-byte granules
constexpr uint32_t W_DIG = W_RAW;                           // base digits (table path only): converted in place
constexpr uint32_t W_PL = W_RAW + DIG_BYTES;                // prefetched prefix row of the tile: TK x uint16
constexpr uint32_t W_DESC = W_PL + TK * 2;                  // two TileDesc slots
constexpr uint32_t W_RD = W_DESC + 2 * 48;                  // two slots of {arena offset (8), samples in the read (4), pad, ADC offset (8), pad}
constexpr uint32_t WARP_BYTES = W_RD + 2 * 32;
constexpr uint32_t SM_Z = 0;
constexpr uint32_t SM_CODE = SM_Z + Z32_BYTES;
constexpr uint32_t SM_MBAR = SM_CODE + 256;
constexpr uint32_t SM_WARP = SM_MBAR + 16;
constexpr uint32_t SM_TOTAL = SM_WARP + K4_WARPS * WARP_BYTES;
static_assert(SM_WARP % 16 == 0 && WARP_BYTES % 16 == 0 && W_PAR % 16 == 0 && W_RAW % 16 == 0 && W_DIG % 16 == 0 && W_PL % 16 == 0 &&
              W_DESC % 16 == 0 && W_RD % 16 == 0, "alignment");
static_assert(SM_TOTAL <= 232448, "227 KB of shared memory per CTA");
static_assert(PAR_N % 2 == 0 && TK == 256 && MAP_ENT % 2 == 0, "layout assumptions");

// ---- per-lane asynchronous global -> shared copies (SASS: LDGSTS) for the next tile's inputs ----
__device__ __forceinline__ void cp_async16(uint32_t saddr, const void *g) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(saddr), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async8(uint32_t saddr, const void *g) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(saddr), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t saddr, const void *g) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(saddr), "l"(g) : "memory");
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

// ---- the amplitude draws.  A Philox block holds TWELVE 10-bit draws - three fields per word: bits
// 7..16 (F0), bits 17..26 (F1), bits 27..31,0..4 (F2) - so two blocks serve three chunks.  The chunks (= 8 samples) of
// the EMITTED signal are taken in UNITS of 96: chunk Cq = 96 u + 32 g + l (g = 0..2, l = 0..31) belongs to blocks
// A = 64 u + 2 l and B = A + 1 (stream ST_AMP), and its sample e uses
//   g = 0: word e/2 of A, field F0 (e even) or F1 (e odd)        g = 2: the same of B
//   g = 1: field F2 of word e of A (e < 4) or of word e - 4 of B (e >= 4)
// A lane of the signal kernel handles the three chunks l, 32 + l, 64 + l of a unit with two Philox calls.  The draw's
// table CLASS is l ^ h, h = five bits of ONE hash per unit and read (amp_mix; group g of the unit takes bits 27-5g .. 31-5g):
// within a group the classes are a bijection of the lanes - one bank per lane - and over the units every position of the
// signal meets every class.
__device__ __forceinline__ uint32_t amp_mix(uint32_t x) {   // x = unit * 0x9E3779B1 + amp_hmul(read): one xorshift-multiply round
    x ^= x >> 15;
    return x * 0x2C1B3C6Du;
}
// (h << 2) of group g (0..2) of a unit whose hash is x
template <int G>
__device__ __forceinline__ uint32_t amp_h4(uint32_t x) { return (x >> (25 - 5 * G)) & 0x7Cu; }
__device__ __forceinline__ uint32_t amp_h4(uint32_t x, uint32_t g) { return (x >> (25u - 5u * g)) & 0x7Cu; }
__device__ __forceinline__ uint32_t amp_class4(uint32_t Cq, uint32_t hmul) {   // any chunk, by itself
    const uint32_t u = Cq / UNIT_C, r = Cq - u * UNIT_C;
    return ((r & 31u) << 2) ^ amp_h4(amp_mix(u * 0x9E3779B1u + hmul), r >> 5);
}
__device__ __forceinline__ uint32_t amp_hmul(uint32_t r_lo) { return r_lo * 0x85EBCA6Bu; }
// a field moved to bits 7..16, where z_offset() masks it: F1 by a multiply-high (x >> 10 on the FMA pipe, which has
// room; the ALU pipe does not), F2 by a rotation
__device__ __forceinline__ uint32_t amp_f1(uint32_t x) {
    uint32_t r;
    asm("mul.hi.u32 %0, %1, 4194304;" : "=r"(r) : "r"(x));
    return r;
}
#ifdef SQG_F2_IMAD
__device__ __forceinline__ uint32_t amp_f2(uint32_t x) {   // the rotation as two multiplies: (x << 12) + (x >> 20), on the FMA pipe
    uint32_t hi, r;
    asm("mul.hi.u32 %0, %1, 4096;" : "=r"(hi) : "r"(x));
    asm("mad.lo.u32 %0, %1, 4096, %2;" : "=r"(r) : "r"(x), "r"(hi));
    return r;
}
#else
__device__ __forceinline__ uint32_t amp_f2(uint32_t x) { return __funnelshift_r(x, x, 20); }
#endif
template <int G>
__device__ __forceinline__ void amp_fields(const uint4 &A, const uint4 &B, uint32_t (&dw)[8]) {
    const uint32_t a[4] = {A.x, A.y, A.z, A.w}, b[4] = {B.x, B.y, B.z, B.w};
#pragma unroll
    for (int e = 0; e < 8; e++) {
        if (G == 1) dw[e] = amp_f2(e < 4 ? a[e] : b[e - 4]);
        else {
            const uint32_t x = G == 0 ? a[e >> 1] : b[e >> 1];
            dw[e] = (e & 1) ? amp_f1(x) : x;
        }
    }
}
// any chunk, by itself (exact path, masked groups, slow tiles): one or two Philox calls
__device__ __forceinline__ void amp_draws(const GenParams &p, uint32_t Cq, uint32_t r_lo, uint32_t r_hi, uint32_t (&dw)[8]) {
    const uint32_t u = Cq / UNIT_C, r = Cq - u * UNIT_C, g = r >> 5;
    const uint32_t blk = 64u * u + 2u * (r & 31u);
    uint4 A = make_uint4(0, 0, 0, 0), B = make_uint4(0, 0, 0, 0);
    // (the key schedule by additions, not from p.rk: in the out-of-line callers `p` is a generic pointer to the kernel's
    //  parameter block, and fourteen generic loads per block would sit on the critical path of a one-lane redo)
    const uint32_t k0 = p.key0, k1 = p.key1;
    if (g != 2) A = philox4x32(blk, r_lo, r_hi, ST_AMP, k0, k1);
    if (g != 0) B = philox4x32(blk + 1, r_lo, r_hi, ST_AMP, k0, k1);
    if (g == 0) amp_fields<0>(A, B, dw);
    else if (g == 1) amp_fields<1>(A, B, dw);
    else amp_fields<2>(A, B, dw);
}

// What a warp knows about the run it is in: consecutive tiles of one read, from where the warp's range (or the read)
// begins to where it ends.  FRAME coordinates f count samples in generation order from a point <= the run's first
// sample chosen such that f = 0 (mod 256) is a group boundary of the emitted signal.  All warp-uniform.
struct Run {
    uint32_t C0;          // emitted chunk (= Philox block) of frame chunk 0: frame chunk c is C0 + c, or C0 - c when reversed
    uint32_t r_lo, r_hi;  // global read index (Philox counter words 1, 2)
    uint32_t hmul;        // amp_hmul(r_lo)
    int16_t *out;         // start of the read in the signal arena
    float c_r;            // 32768 - (float)offset  (noisy modes)
    double offset;        // the read's ADC offset
    uint32_t L;           // samples in the read
    uint32_t clip_lo;     // first frame sample this run owns (0: the run starts its read - what lies before is padding)
    uint32_t f_end;       // one past the last registered frame sample
    uint32_t fmap;        // frame coordinate of map entry 0 (a multiple of 256)
    uint32_t cur_c;       // next frame chunk to emit
    int32_t nreg;         // k-mers in par[]
    int32_t fix_f0;       // fixed-dwell modes: frame coordinate of the first sample of par[0]'s k-mer
};

// per-lane constants of the sample loop
struct LaneC {
    uint32_t lw;        // the lane's chunk within a group in FRAME order (reversed reads: 31 - lane)
    uint32_t ent_sh;    // 8 * (lw & 3): the chunk's byte within its map entry
    uint32_t ent_lane;  // 8 * (lw >> 2): byte offset of its entry within the group's eight
    uint32_t lane4;     // lane << 2
};

// k-mer (index into par[]) of a chunk's first sample and the chunk's boundary mask (bit j, 1..7: a k-mer starts at slot j
// in generation order).  A start on the chunk's first sample is not a boundary to cross.
__device__ __forceinline__ void entry_kmers(uint2 ent, uint32_t sh, uint32_t &k0, uint32_t &m1) {
    k0 = ent.y + __popc(ent.x & ((2u << sh) - 1u));
    m1 = (ent.x >> sh) & 0xFEu;
}
template <bool RAND_DWELL>
__device__ __forceinline__ void chunk_kmers(const GenParams &p, const unsigned char *smem, uint32_t map_off, uint32_t fmap, int32_t fix_f0,
                                            uint32_t c, uint32_t &k0, uint32_t &m1) {
    if (RAND_DWELL) {
        entry_kmers(*reinterpret_cast<const uint2 *>(smem + map_off + W_MAP + 8 * ((8 * c - fmap) >> 5)), 8 * (c & 3), k0, m1);
    } else {
        const int s0 = (int)(8 * c) - fix_f0;
        k0 = div_sps(p, (uint32_t)max(s0, 0));
        m1 = 0;
        for (int b = (int)((k0 + 1) * (uint32_t)p.sps_fixed) - s0; b < 8; b += p.sps_fixed) m1 |= 1u << b;
    }
}

// ---- phase B ---------------------------------------------------------------------------------------------------

// The exact path of one chunk, start to finish (rare: a chunk shared with another warp's range, a chunk with a flagged
// sample - tail cell of the table, negative value, value beyond int16 - or with four or more k-mers, and every chunk in
// wide mode): samples are trunc(fma.rz(z, A', Bq)) with the tail cells refined and any number of boundaries; only frame
// samples in [clip_lo, clip_hi) are stored.
template <bool NOISY, bool RAND_DWELL, bool REV>
__device__ __noinline__ void exact_chunk(const GenParams &p, const unsigned char *smem, uint32_t map_off, uint32_t fmap, int32_t fix_f0,
                                         uint32_t C0, uint32_t r_lo, uint32_t r_hi, uint32_t hmul, float c_r, int16_t *out, uint32_t c,
                                         uint32_t clip_lo, uint32_t clip_hi) {
    uint32_t k0, m1;
    chunk_kmers<RAND_DWELL>(p, smem, map_off, fmap, fix_f0, c, k0, m1);
    const uint32_t par0 = k0 * 8 + map_off + W_PAR;
    const uint32_t Cq = REV ? C0 - c : C0 + c;
    const uint32_t class4 = amp_class4(Cq, hmul);
    const RngKey key{p.key0, p.key1, r_lo, r_hi};
    uint32_t dw[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (NOISY) amp_draws(p, Cq, r_lo, r_hi, dw);
    int16_t *dst = out + (size_t)Cq * 8;
#pragma unroll
    for (int e = 0; e < 8; e++) {
        const int j = REV ? 7 - e : e;
        const uint32_t f = 8 * c + j;
        if (f >= clip_lo && f < clip_hi) {
            const float2 ab = *reinterpret_cast<const float2 *>(smem + par0 + 8 * __popc(m1 & ((2u << j) - 1u)));
            uint32_t v;
            if (NOISY) {
                const uint32_t off = z_offset(dw[e], class4);
                float z = *reinterpret_cast<const float *>(smem + SM_Z + off);
                if (z_is_tail(off)) z = z_tail(p.z2, off, Cq * 8 + e, key, ST_AMP_TAIL);
                v = sample_exact(z, ab.x, __fsub_rn(__fadd_rn(ab.y, c_r), SAMPLE_MAGIC));   // par[] holds (A', M)
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
// FMA pipe, and no shared-memory traffic of its own.  Returns non-zero when the chunk has to be redone by the exact path
// (flagged sample, 4+ k-mers).
template <bool NOISY, bool REV>
__device__ __forceinline__ uint32_t fast_chunk(uint32_t k0, uint32_t m1, uint32_t par_base, const uint32_t (&dw)[8], uint32_t class4, float c_r,
                                               bool l2_vote, bool converged, uint4 &pk) {
    const uint32_t par0 = k0 * 8 + par_base;
    float2 q0 = lds_f2<0>(par0), q1 = lds_f2<8>(par0);
    if (NOISY) {   // par[] holds (A', M): B' + 32768 = M + c_r, one rounding
        q0.y = __fadd_rn(q0.y, c_r); q1.y = __fadd_rn(q1.y, c_r);
    }
    const uint32_t t1 = m1 - 1u;        // bit j clear  <=>  slot j lies at or after the 1st boundary
    const uint32_t m2 = m1 & t1;        // boundaries after the first
    uint32_t m3 = 0;                    // non-zero after the levels: more boundaries than they cover -> exact path
    uint32_t bad;
    if (NOISY) {
        float zz[8], v[8];
#pragma unroll
        for (int e = 0; e < 8; e++) {   // e = slot in the emitted chunk = which draw; j = slot in generation order
            zz[e] = lds_f32<SM_Z>(z_offset(dw[e], class4));
            v[e] = fma_rz(zz[e], q0.x, q0.y);
        }
#pragma unroll
        for (int e = 0; e < 8; e++) {   // level by level, so that one R2P per level sets the predicates
            const int j = REV ? 7 - e : e;
            if (j >= 1 && !(t1 & (1u << j))) v[e] = fma_rz(zz[e], q1.x, q1.y);
        }
        // a third k-mer in the chunk: with long dwells few chunks have one, and then (l2_vote) the level is skipped
        // whenever no lane of the warp needs it; a fourth one is rare with every profile, but not rare enough (dwell 9 +- 4:
        // one chunk in 300) to send its chunk through the exact path: its level runs when some lane of the warp has one
        if (!l2_vote || __any_sync(0xffffffffu, m2 != 0)) {
            float2 q2 = lds_f2<16>(par0);
            q2.y = __fadd_rn(q2.y, c_r);
            const uint32_t t2 = m2 - 1u;        // bit j clear  <=>  slot j lies at or after the 2nd boundary
            m3 = m2 & t2;
#pragma unroll
            for (int e = 0; e < 8; e++) {
                const int j = REV ? 7 - e : e;
                if (j >= 2 && !(t2 & (1u << j))) v[e] = fma_rz(zz[e], q2.x, q2.y);
            }
            if (converged && __any_sync(0xffffffffu, m3 != 0)) {
                float2 q3 = lds_f2<24>(par0);
                q3.y = __fadd_rn(q3.y, c_r);
                const uint32_t t3 = m3 - 1u;    // bit j clear  <=>  slot j lies at or after the 3rd boundary
#pragma unroll
                for (int e = 0; e < 8; e++) {
                    const int j = REV ? 7 - e : e;
                    if (j >= 3 && !(t3 & (1u << j))) v[e] = fma_rz(zz[e], q3.x, q3.y);
                }
                m3 &= t3;                        // non-zero: a 4th boundary -> exact path
            }
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
        const float2 q2 = lds_f2<16>(par0);
        const uint32_t t2 = m2 - 1u;
        m3 = m2 & t2;
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
    return bad;
}

// the chunk of this lane in frame group g: k-mers, class, samples from the draw words dw.  Returns the redo flag; the
// caller stores pk.
template <bool NOISY, bool RAND_DWELL, bool REV>
__device__ __forceinline__ uint32_t group_chunk(const GenParams &p, const unsigned char *smem, const Run &t, const LaneC &lc, uint32_t map_off,
                                                uint32_t g, const uint32_t (&dw)[8], uint32_t &Cq, uint4 &pk) {
    const uint32_t c = 32 * g + lc.lw;
    Cq = REV ? t.C0 - c : t.C0 + c;
    uint32_t k0, m1;
    if (RAND_DWELL) {
        const uint2 ent = lds_u2<W_MAP>(map_off + 64 * (g - (t.fmap >> 8)) + lc.ent_lane);
        entry_kmers(ent, lc.ent_sh, k0, m1);
    } else {
        chunk_kmers<false>(p, smem, map_off, t.fmap, t.fix_f0, c, k0, m1);
    }
    const uint32_t class4 = amp_class4(Cq, t.hmul);   // (= lane4 ^ the group's five hash bits: groups are aligned in the emitted signal)
    return fast_chunk<NOISY, REV>(k0, m1, map_off + W_PAR, dw, class4, t.c_r, false, false, pk);
}

// One group of 32 chunks by itself (where a run begins or ends): only frame chunks [c_lo, c_hi) are emitted, and those
// reaching outside [clip_lo, clip_hi) - or all of them when `all_exact` - take the exact path.
template <bool NOISY, bool RAND_DWELL, bool REV>
__device__ __forceinline__ void emit_group_single(const GenParams &p, const unsigned char *smem, const Run &t, const LaneC &lc, uint32_t map_off,
                                                  uint32_t g, uint32_t c_lo, uint32_t c_hi, uint32_t clip_hi, bool all_exact) {
    const uint32_t c = 32 * g + lc.lw;
    if (c < c_lo || c >= c_hi) return;
    uint32_t lo = t.clip_lo, hi = clip_hi;
    bool redo = all_exact || 8 * c < t.clip_lo || 8 * c + 8 > clip_hi;
    if (!redo) {
        uint32_t Cq, dw[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        uint4 pk;
        if (NOISY) amp_draws(p, REV ? t.C0 - c : t.C0 + c, t.r_lo, t.r_hi, dw);
        redo = group_chunk<NOISY, RAND_DWELL, REV>(p, smem, t, lc, map_off, g, dw, Cq, pk) != 0;
        st_cs_v4(t.out + (size_t)Cq * 8, pk);
        lo = 0u; hi = 0xFFFFFFFFu;
    }
    if (redo) exact_chunk<NOISY, RAND_DWELL, REV>(p, smem, map_off, t.fmap, t.fix_f0, t.C0, t.r_lo, t.r_hi, t.hmul, t.c_r, t.out, c, lo, hi);
}

// One whole unit: frame groups 3 uf .. 3 uf + 2.  Three independent chunks per lane from two Philox blocks, one check for
// the rare redo behind them.  ent_addr / dst / blk / H walk from unit to unit in the caller: the lane's map entry of the
// unit's first group, where its first chunk goes, its first Philox block, the class hash's argument of the first group.
template <bool NOISY, bool RAND_DWELL, bool REV, bool L2V>
__device__ __forceinline__ void emit_unit_fast(const GenParams &p, const unsigned char *smem, const Run &t, const LaneC &lc, uint32_t map_off,
                                               uint32_t uf, uint32_t ent_addr, int16_t *dst, uint32_t blk, uint32_t H) {
    uint4 A = make_uint4(0, 0, 0, 0), B = make_uint4(0, 0, 0, 0);
    if (NOISY) {
        A = philox4x32_rk(blk, t.r_lo, t.r_hi, ST_AMP, p.rk);
        B = philox4x32_rk(blk + 1, t.r_lo, t.r_hi, ST_AMP, p.rk);
    }
    uint32_t bad[3];
    uint4 pk[3];
    const uint32_t hx = amp_mix(H);
#pragma unroll
    for (int i = 0; i < 3; i++) {
        uint32_t k0, m1, dw[8];
        if (RAND_DWELL) {
            const uint2 ent = i == 0 ? lds_u2<W_MAP>(ent_addr) : i == 1 ? lds_u2<W_MAP + 64>(ent_addr) : lds_u2<W_MAP + 128>(ent_addr);
            entry_kmers(ent, lc.ent_sh, k0, m1);
        } else {
            chunk_kmers<false>(p, smem, map_off, t.fmap, t.fix_f0, 32 * (3 * uf + i) + lc.lw, k0, m1);
        }
        // frame order = emitted order (reversed reads: the unit's chunks come last group first)
        if (i == 0) amp_fields<REV ? 2 : 0>(A, B, dw);
        else if (i == 1) amp_fields<1>(A, B, dw);
        else amp_fields<REV ? 0 : 2>(A, B, dw);
        const uint32_t class4 = lc.lane4 ^ (i == 0 ? amp_h4<REV ? 2 : 0>(hx) : i == 1 ? amp_h4<1>(hx) : amp_h4<REV ? 0 : 2>(hx));
        bad[i] = fast_chunk<NOISY, REV>(k0, m1, map_off + W_PAR, dw, class4, t.c_r, L2V, true, pk[i]);
#ifdef SQG_STORE_EARLY
        st_cs_v4(REV ? dst - 256 * i : dst + 256 * i, pk[i]);
#endif
    }
#ifndef SQG_STORE_EARLY
#pragma unroll
    for (int i = 0; i < 3; i++) {
#ifdef SQG_KO_STORE
        if (pk[i].x == 0x12345678u && pk[i].y == 0x9abcdef0u)
#endif
        st_cs_v4(REV ? dst - 256 * i : dst + 256 * i, pk[i]);
    }
#endif
#ifdef SQG_KO_EXACT
    if (pk[0].x == 0x12345678u && pk[1].y == 0x9abcdef0u && (bad[0] | bad[1] | bad[2]) != 0) {
#else
    if (__builtin_expect((bad[0] | bad[1] | bad[2]) != 0, 0)) {
#endif
#pragma unroll
        for (int i = 0; i < 3; i++)
            if (bad[i])
                exact_chunk<NOISY, RAND_DWELL, REV>(p, smem, map_off, t.fmap, t.fix_f0, t.C0, t.r_lo, t.r_hi, t.hmul, t.c_r, t.out, 32 * (3 * uf + i) + lc.lw, 0u, 0xFFFFFFFFu);
    }
}

// Emit what has become complete: whole units by the fast path; where a run begins (up to its first unit boundary) and
// where it ends, single groups.  `last`: the run ends with the samples registered so far.
template <bool NOISY, bool RAND_DWELL, bool REV>
__device__ __forceinline__ void emit_ready(const GenParams &p, const unsigned char *smem, Run &t, const LaneC &lc, uint32_t map_off, int lane,
                                           bool last, uint32_t clip_hi) {
    const uint32_t hi_c = last ? (t.f_end + 7) >> 3 : t.f_end >> 3;   // chunks below hi_c can be computed
    const bool all_exact = NOISY && p.wide;
#ifdef SQG_KO_EMIT
    if (t.f_end != 0x7FFFFFFFu) { t.cur_c = last ? hi_c : (hi_c / UNIT_C) * UNIT_C; return; }
#endif
    while (t.cur_c < hi_c) {
        if (!all_exact && t.cur_c % UNIT_C == 0 && 8 * t.cur_c >= t.clip_lo) {
            uint32_t uf = t.cur_c / UNIT_C;
            const uint32_t u_end = min(hi_c, clip_hi >> 3) / UNIT_C;   // whole units: [uf, u_end)
            if (uf < u_end) {
                // the lane's walk over the units: map entry, output chunk, Philox block, class hash of the first group
                const uint32_t c0 = t.cur_c + lc.lw;
                const uint32_t Cq0 = REV ? t.C0 - c0 : t.C0 + c0;
                uint32_t ent_addr = map_off + 8 * ((8 * t.cur_c - t.fmap) >> 5) + lc.ent_lane;
                int16_t *dst = t.out + (size_t)Cq0 * 8;
                const uint32_t u0 = REV ? (t.C0 / UNIT_C) - uf : (t.C0 / UNIT_C) + uf;   // emitted unit (frame units are aligned with them)
                uint32_t blk = 64u * u0 + 2u * (uint32_t)lane;
                uint32_t H = u0 * 0x9E3779B1u + t.hmul;
                // (two copies of the loop: whether a chunk's third k-mer level is voted on is a per-profile constant, and a
                //  test of it inside the loop costs seven issue slots per unit)
                if (p.l2_vote) {
                    for (; uf < u_end; uf++) {
                        emit_unit_fast<NOISY, RAND_DWELL, REV, true>(p, smem, t, lc, map_off, uf, ent_addr, dst, blk, H);
                        ent_addr += 3 * 64;
                        dst = REV ? dst - 768 : dst + 768;
                        blk = REV ? blk - 64 : blk + 64;
                        H = REV ? H - 0x9E3779B1u : H + 0x9E3779B1u;
                    }
                } else {
                    for (; uf < u_end; uf++) {
                        emit_unit_fast<NOISY, RAND_DWELL, REV, false>(p, smem, t, lc, map_off, uf, ent_addr, dst, blk, H);
                        ent_addr += 3 * 64;
                        dst = REV ? dst - 768 : dst + 768;
                        blk = REV ? blk - 64 : blk + 64;
                        H = REV ? H - 0x9E3779B1u : H + 0x9E3779B1u;
                    }
                }
                t.cur_c = UNIT_C * u_end;
                if (!last) break;   // (what is left is less than a unit: it completes with the next tile)
                continue;
            }
            if (!last) break;   // the unit completes with the next tile
        }
        const uint32_t g = t.cur_c >> 5, gend_c = 32 * (g + 1);
        if (!last && gend_c > hi_c) break;   // the group completes with the next tile
        const uint32_t ce = min(gend_c, hi_c);
        emit_group_single<NOISY, RAND_DWELL, REV>(p, smem, t, lc, map_off, g, t.cur_c, ce, clip_hi, all_exact);
        t.cur_c = ce;
    }
}

// ---- phase A ---------------------------------------------------------------------------------------------------

constexpr int WIN_LOADS = (TK + 8 + 31) / 32;  // bytes per lane of a tile's base window (k <= 9)

// a tile whose base window straddles the two pieces of its segment (only around a --prefix junction)
__device__ __forceinline__ bool tile_is_junction(const GenParams &p, int32_t a_rem, int32_t nk) {
    return a_rem > 0 && a_rem < nk + p.k - 1;
}

// Asynchronous fetch of a tile's inputs into the warp's buffer: the base window as 16-byte granules (the aligned
// superset of the window), its prefix row, its read's record (arena offset, length, ADC offset).  `desc_off` = the tile's
// descriptor, already in shared memory.
template <bool RAND_DWELL>
__device__ __forceinline__ void fetch_tile_inputs(const GenParams &p, const unsigned char *smem, uint32_t wbase, uint32_t desc_off,
                                                  uint32_t rd_rel, int tile, int lane) {
    const uint4 a = *reinterpret_cast<const uint4 *>(smem + desc_off);
    const uint4 b = *reinterpret_cast<const uint4 *>(smem + desc_off + 16);
    const int32_t a_rem = (int32_t)b.x, nk = (int32_t)(b.y & 0xFFFFu), read = (int32_t)b.z;
    if (!tile_is_junction(p, a_rem, nk)) {
        const int64_t off = a_rem > 0 ? (int64_t)(((uint64_t)a.y << 32) | a.x) : (int64_t)(((uint64_t)a.w << 32) | a.z);
        const uint8_t *g = p.bases + off;
        const uint32_t shift = (uint32_t)(reinterpret_cast<uintptr_t>(g) & 15u);
        const uint32_t nbytes = shift + (uint32_t)(nk + p.k - 1);
        if ((uint32_t)lane * 16 < nbytes) cp_async16(wbase + W_RAW + lane * 16, g - shift + lane * 16);
    }
    if (RAND_DWELL) cp_async16(wbase + W_PL + lane * 16, p.kpos + (size_t)tile * (TK / 8) + lane);
    if (lane < 2) cp_async16(wbase + rd_rel + lane * 16, reinterpret_cast<const uint4 *>(p.read_rec + read) + lane);
}
__device__ __forceinline__ void fetch_tile_desc(const GenParams &p, uint32_t wbase, uint32_t desc_rel, int tile, int lane) {
    if (lane < 3) cp_async16(wbase + desc_rel + lane * 16, reinterpret_cast<const uint4 *>(p.tiles + tile) + lane);
}

struct TileIn {   // the descriptor fields phase A works from (warp-uniform)
    int64_t a_off, b_off;
    int32_t a_rem, nk;
    uint32_t S;
};

// Phase A of one tile: its k-mers are appended to the run's window.  Descriptor, base window and prefix row are already
// in the warp's buffer.
template <bool NOISY, bool RAND_DWELL, bool METH, bool REV>
__device__ __forceinline__ void register_tile(const GenParams &p, int lane, unsigned char *smem, uint32_t map_off, uint32_t wbase, Run &t,
                                              const TileIn &td) {
    const int nk_tile = td.nk;
#ifdef SQG_KO_PHASEA
    if (t.f_end != 0x7FFFFFFFu) return;   // (timing only: the sample loop runs on the guard entries)
#endif
    const int nb = nk_tile + p.k - 1;
    const uint32_t dig_off = map_off + W_DIG;
    const int m0 = lane * 8;

    // (1) bases -> digits.  Fast path (base-4 models, window in one piece): every lane takes the 16 raw bytes of its own
    // 8 k-mers straight from the prefetched window (three aligned 8-byte loads + a funnel shift by the window's
    // misalignment) and turns A/C/G/T of either case into digits arithmetically, ((c>>1) ^ (c>>2)) & 3, four bytes at
    // a time; a PRMT maps the digits back to letters to check that every byte really was one of those eight.  Any
    // other byte in the tile (IUPAC codes, U, N: src/seq.h:14-28 folds them) sends the whole warp through the
    // 256-entry code table, which is also the path of base-5 (CpG) models (src/seq.h:45-60) and of prefix junctions.
    uint32_t dg[4] = {0, 0, 0, 0};   // the lane's 16 digits, one per byte
    bool table_path = tile_is_junction(p, td.a_rem, nk_tile);
    if (!table_path) {
        const uint32_t shift = (uint32_t)(reinterpret_cast<uintptr_t>(p.bases + (td.a_rem > 0 ? td.a_off : td.b_off)) & 15u);
        const uint32_t s0 = shift + 8u * (uint32_t)lane;     // lane's first byte within the raw buffer
        const uint2 *rw = reinterpret_cast<const uint2 *>(smem + map_off + W_RAW + (s0 & ~7u));
        const uint2 w0 = rw[0], w1 = rw[1], w2 = rw[2];
        const bool hi = (s0 & 4u) != 0;                       // (warp-uniform: shift & 4)
        const uint32_t q0 = hi ? w0.y : w0.x, q1 = hi ? w1.x : w0.y, q2 = hi ? w1.y : w1.x, q3 = hi ? w2.x : w1.y, q4 = hi ? w2.y : w2.x;
        const uint32_t fs = 8u * (s0 & 3u);
        const uint32_t x[4] = {__funnelshift_r(q0, q1, fs), __funnelshift_r(q1, q2, fs), __funnelshift_r(q2, q3, fs), __funnelshift_r(q3, q4, fs)};
        // base 4: ACGTacgt -> ((c>>1) ^ (c>>2)) & 3, checked by mapping the digits back ("ACGT" by PRMT) against the byte
        // with its case bit cleared.  Base 5 (src/seq.h:45-60, upper case only): A C G M T = 41 43 47 4D 54 ->
        // b1 + (b1 & b2) + 3 b3 + 4 b4 (b_i = bit i of the byte) = 0 1 2 3 4, checked against "ACGMT" exactly.
        auto digits_of = [&](uint32_t w) -> uint32_t {
            if (!METH) return ((w >> 1) ^ (w >> 2)) & 0x03030303u;
            const uint32_t b1 = (w >> 1) & 0x01010101u, b2 = (w >> 2) & 0x01010101u, b3 = (w >> 3) & 0x01010101u, b4 = (w >> 4) & 0x01010101u;
            return b1 + (b1 & b2) + 3u * b3 + 4u * b4;
        };
        auto check_of = [&](uint32_t c, uint32_t w) -> uint32_t {
            const uint32_t u = (c | (c >> 4)) & 0x00FF00FFu;
            const uint32_t sel = (u | (u >> 8)) & 0xFFFFu;                    // the four digits as PRMT selectors
            return METH ? (__byte_perm(0x4D474341u /* "ACGM" */, 0x00000054u /* "T" */, sel) ^ w)
                        : (__byte_perm(0x54474341u /* "ACGT" */, 0u, sel) ^ (w & 0xDFDFDFDFu));
        };
        uint32_t bad = 0;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            dg[i] = digits_of(x[i]);
            bad |= check_of(dg[i], x[i]);
        }
        // bytes past the window's end are whatever the 16-byte granules held: they must not force the table path, so
        // only the lane's bytes inside the window count - and their digits are cleared (base 5: such a "digit" can exceed
        // 4, and the pair rank of the tile's last k-mer takes one digit from behind the window)
        const int inside = nb - 8 * lane;
        if (inside < 16) {
            uint32_t keep = 0;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int nbytes = min(max(inside - 4 * i, 0), 4);
                const uint32_t m = nbytes >= 4 ? 0xFFFFFFFFu : ((1u << (8 * nbytes)) - 1u);
                keep |= check_of(dg[i], x[i]) & m;
                dg[i] &= m;
            }
            bad = keep;
        }
        table_path = __any_sync(0xffffffffu, bad != 0);
    }
    if (table_path) {
        if (!tile_is_junction(p, td.a_rem, nk_tile)) {
            const uint32_t shift = (uint32_t)(reinterpret_cast<uintptr_t>(p.bases + (td.a_rem > 0 ? td.a_off : td.b_off)) & 15u);
            const uint32_t raw_off = map_off + W_RAW + shift + lane;
            uint32_t cc[WIN_LOADS];
#pragma unroll
            for (int u = 0; u < WIN_LOADS; u++) cc[u] = (lane + 32 * u < nb) ? smem[SM_CODE + smem[raw_off + 32 * u]] : 0u;
            __syncwarp();   // (the digits replace the raw bytes in place)
#pragma unroll
            for (int u = 0; u < WIN_LOADS; u++)
                if (lane + 32 * u < nb) smem[dig_off + lane + 32 * u] = (unsigned char)(METH ? (cc[u] >> 4) : (cc[u] & 3u));
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
        // (what lies behind the window in this buffer is raw bytes, not digits: cleared, as on the arithmetic path)
        const int inside = nb - 8 * lane;
        if (inside < 16) {
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int nbytes = min(max(inside - 4 * i, 0), 4);
                dg[i] &= nbytes >= 4 ? 0xFFFFFFFFu : ((1u << (8 * nbytes)) - 1u);
            }
        }
        __syncwarp();   // (the digit buffer is rewritten by this warp's next tile)
    }

    // (1b) the table gathers of this lane.  Noisy modes: asynchronous copies (LDGSTS) from the (A', M) tables STRAIGHT INTO
    // the window - no registers, no stores of our own, and a random-line gather costs the load/store unit far less this way
    // than as LDG (measured: 0.17 ms instead of 0.42 ms per 2.1 G samples); the read's constant c_r is added where the
    // samples are made.  Base-4 models: 16 two-bit digits packed first-digit-most-significant
    // (((w & 0x03030303) * 0x40100401) >> 24 packs 4 bytes); k-mers 2j, 2j+1 of the lane = the two k-mers of the
    // (k+1)-mer at digit 2j: ONE 16-byte gather for both.  Ideal-amplitude modes: the raw table into registers (double
    // arithmetic below).
    float2 mv[8];
    const uint32_t par_w = map_off + W_PAR + 8 * (uint32_t)(t.nreg + m0);
    const uint32_t par_a = wbase + W_PAR + 8 * (uint32_t)(t.nreg + m0);   // (the same, as an address for the asynchronous copies)
    const bool pairs = NOISY && (t.nreg & 1) == 0;   // (16-byte copies need an even window position)
    uint32_t ranks[8], pr[4];
    if (!METH) {
        const uint32_t P = ((((dg[0] & 0x03030303u) * 0x40100401u) >> 24) << 24) | ((((dg[1] & 0x03030303u) * 0x40100401u) >> 24) << 16) |
                           ((((dg[2] & 0x03030303u) * 0x40100401u) >> 24) << 8) | (((dg[3] & 0x03030303u) * 0x40100401u) >> 24);
        if (pairs) {
            const int sh0 = 30 - 2 * p.k;                 // 32 - 2(k+1)
            const uint32_t pmask = (p.kmask << 2) | 3u;   // 4^(k+1) - 1
#pragma unroll
            for (int j = 0; j < 4; j++) pr[j] = (P >> (sh0 - 4 * j)) & pmask;
        } else {
            const int sh0 = 32 - 2 * p.k;
#pragma unroll
            for (int j = 0; j < 8; j++) ranks[j] = (P >> (sh0 - 2 * j)) & p.kmask;
        }
    } else {
        // base-5 (CpG) ranks, src/seq.h:62-74, rolled: rank' = 5*rank - 5^k*(leading digit) + (new digit); the (k+1)-mer
        // at an even position = 5 * (its first k-mer) + (the digit behind it)
        const uint32_t dw[4] = {dg[0], dg[1], dg[2], dg[3]};
        const int km1 = p.k - 1;
        uint32_t rank = 0;
#pragma unroll
        for (int i = 0; i < 8; i++)
            if (i < km1) rank = rank * 5 + ((dw[i >> 2] >> (8 * (i & 3))) & 0xFFu);
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const int bi = km1 + j;
            const uint32_t word = bi < 4 ? dw[0] : bi < 8 ? dw[1] : bi < 12 ? dw[2] : dw[3];
            rank = rank * 5 + ((word >> (8 * (bi & 3))) & 0xFFu);             // k digits: the k-mer at lane position j
            ranks[j] = rank;
            if (j & 1) pr[j >> 1] = ranks[j - 1] * 5 + ((word >> (8 * (bi & 3))) & 0xFFu);
            rank -= ((dw[j >> 2] >> (8 * (j & 3))) & 0xFFu) * p.kmask;        // drop its leading digit (kmask = 5^(k-1))
        }
    }
    if (pairs) {
        // (the 16-byte shared-memory writes of these copies collide in the banks - lane stride 64 bytes - but visiting the
        // pairs in a rotated order to avoid that costs more in instructions than the replays do: measured)
#pragma unroll
        for (int j = 0; j < 4; j++) {
#ifdef SQG_KO_GATHER
            if (m0 + 2 * j < nk_tile) cp_async16(par_a + 16 * j, &p.pair_model[(td.nk & 0xFF) * 128 + j * 32 + lane]);   // coalesced (timing only)
#else
            if (m0 + 2 * j < nk_tile) cp_async16(par_a + 16 * j, &p.pair_model[pr[j]]);
#endif
        }
    } else {
        // one gather per k-mer, by rank: ideal amplitudes (the raw table, double arithmetic below), and a window position
        // the 16-byte copies cannot take (a second segment behind an odd number of k-mers)
        if (NOISY) {
#pragma unroll
            for (int j = 0; j < 8; j++)
                if (m0 + j < nk_tile) cp_async8(par_a + 8 * j, &p.model_am[ranks[j]]);
        } else {
#pragma unroll
            for (int j = 0; j < 8; j++)
                if (m0 + j < nk_tile) mv[j] = __ldg(&p.model[ranks[j]]);
        }
    }
    if (NOISY) cp_async_commit();

    // (2) k-mer starts into the map.  Frame position of k-mer i of the tile = (frame position of the tile) + (prefix of
    // the dwells before it, from K1): bit (pos & 31) of entry pos >> 5.  Then the entries' `base` words: the k-mer (index
    // into par[]) of a sample is base + popcount(start bits at or before the sample), so base[e+1] = base[e] +
    // popcount(bits[e]) - a warp scan over the window's entries, six per lane, starting from the entry that holds the
    // tile's first sample (whose base is already valid) rounded down to a 16-byte boundary.
#ifdef SQG_KO_MAP
    if (false) {
#else
    if (RAND_DWELL) {
#endif
        const uint4 pq = *reinterpret_cast<const uint4 *>(smem + map_off + W_PL + lane * 16);
        const uint32_t mrel = t.f_end - t.fmap;      // the tile's first sample, relative to map entry 0
        const uint32_t pw[4] = {pq.x, pq.y, pq.z, pq.w};
        if (nk_tile == TK) {
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const uint32_t pos = mrel + ((j & 1) ? (pw[j >> 1] >> 16) : (pw[j >> 1] & 0xFFFFu));
                atomicOr(reinterpret_cast<uint32_t *>(smem + map_off + W_MAP + 8 * (pos >> 5)), __funnelshift_l(0u, 1u, pos));
            }
        } else {
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const uint32_t pos = mrel + ((j & 1) ? (pw[j >> 1] >> 16) : (pw[j >> 1] & 0xFFFFu));
                if (m0 + j < nk_tile) atomicOr(reinterpret_cast<uint32_t *>(smem + map_off + W_MAP + 8 * (pos >> 5)), __funnelshift_l(0u, 1u, pos));
            }
        }
        __syncwarp();
        const uint32_t e_al = (mrel >> 5) & ~1u;
        const uint32_t e_mine = e_al + 6u * (uint32_t)lane;
        const bool inw = e_mine + 6u <= (uint32_t)(MAP_ENT + 2);   // (MAP_ENT + 2 is a multiple of 2; partial sixes at the end are left alone:
                                                                    //  the window check keeps the tile's samples below them)
        uint4 w[3] = {make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0)};
        uint4 *ep = reinterpret_cast<uint4 *>(smem + map_off + W_MAP + 8 * e_mine);
        if (inw) { w[0] = ep[0]; w[1] = ep[1]; w[2] = ep[2]; }
        const uint32_t c0 = __popc(w[0].x), c1 = __popc(w[0].z), c2 = __popc(w[1].x), c3 = __popc(w[1].z), c4 = __popc(w[2].x), c5 = __popc(w[2].z);
        const uint32_t mine = c0 + c1 + c2 + c3 + c4 + c5;
        uint32_t run = mine;
#pragma unroll
        for (int sh = 1; sh < 32; sh <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, run, sh);
            if (lane >= sh) run += v;
        }
        const uint32_t b0 = __shfl_sync(0xffffffffu, w[0].y, 0) + run - mine;   // base of this lane's first entry
        if (inw) {
            ep[0] = make_uint4(w[0].x, b0, w[0].z, b0 + c0);
            ep[1] = make_uint4(w[1].x, b0 + c0 + c1, w[1].z, b0 + c0 + c1 + c2);
            ep[2] = make_uint4(w[2].x, b0 + c0 + c1 + c2 + c3, w[2].z, b0 + c0 + c1 + c2 + c3 + c4);
        }
    }

    // (3) the parameters of this lane's 8 k-mers.  Noisy modes: they are arriving by themselves.  Ideal amplitudes:
    // src/gensig.c:266,270: (double)level_mean*digitisation/range - offset, truncated - the sample itself.
    if (NOISY) {
        cp_async_wait_all();
    } else {
        const double off_d = t.offset;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            if (m0 + j < nk_tile) {
                const double v = __dsub_rn(__ddiv_rn(__dmul_rn((double)mv[j].x, p.digitisation), p.range), off_d);
                *reinterpret_cast<float2 *>(smem + par_w + 8 * j) = make_float2(0.f, __uint_as_float(to_i16_bits(v)));
            }
        }
    }
    t.nreg += nk_tile;
    __syncwarp();
}

// After a tile's groups have been emitted: move what the next groups still need to the front of the window.
template <bool RAND_DWELL>
__device__ __forceinline__ void slide_window(const GenParams &p, int lane, unsigned char *smem, uint32_t map_off, Run &t) {
    const uint32_t gdone = (8 * t.cur_c - t.fmap) >> 8;   // whole groups emitted since the map's origin
    if (gdone == 0) return;
#ifdef SQG_KO_SLIDE
    if (t.f_end != 0x7FFFFFFFu) { t.nreg = 0; t.fmap += GROUP_S * gdone; return; }
#endif
    const uint32_t fmap2 = t.fmap + GROUP_S * gdone;
    int32_t kt;       // first k-mer still needed (kept even: parameter stores are 16 bytes wide)
    int32_t n_ent = 0;
    const uint32_t e_sh = 8 * gdone;
    if (RAND_DWELL) {
        // (every entry of the window has a valid base after register_tile's scan, the one behind the last sample included)
        const uint2 ent = *reinterpret_cast<const uint2 *>(smem + map_off + W_MAP + 8 * e_sh);
        kt = (int32_t)(ent.y + (ent.x & 1u));
        n_ent = (int32_t)((t.f_end - t.fmap) >> 5) - (int32_t)e_sh + 1;   // <= 9: less than a group, and the entry of the next sample
    } else {
        const int32_t d = (int32_t)fmap2 - t.fix_f0;
        kt = d > 0 ? (int32_t)div_sps(p, (uint32_t)d) : 0;
    }
    kt = max(min(kt, t.nreg), 0) & ~1;
    __syncwarp();
    if (kt > 0) {
        const int32_t n = t.nreg - kt;
        for (int32_t i0 = 0; i0 < n; i0 += 64) {
            float2 v[2];
#pragma unroll
            for (int u = 0; u < 2; u++) {
                const int32_t i = i0 + lane + 32 * u;
                if (i < n) v[u] = *reinterpret_cast<const float2 *>(smem + map_off + W_PAR + 8 * (kt + i));
            }
            __syncwarp();
#pragma unroll
            for (int u = 0; u < 2; u++) {
                const int32_t i = i0 + lane + 32 * u;
                if (i < n) *reinterpret_cast<float2 *>(smem + map_off + W_PAR + 8 * i) = v[u];
            }
            __syncwarp();
        }
    }
    if (RAND_DWELL) {
        // entries [e_sh, e_sh + n_ent) -> [0, n_ent), bases re-indexed; everything behind them cleared
        uint2 ent = make_uint2(0u, 0u);
        if (lane < n_ent) ent = *reinterpret_cast<const uint2 *>(smem + map_off + W_MAP + 8 * (e_sh + lane));
        __syncwarp();
        for (int32_t x = lane; x < (MAP_ENT + 2) / 2; x += 32) *reinterpret_cast<uint4 *>(smem + map_off + W_MAP + 16 * x) = make_uint4(0u, 0u, 0u, 0u);
        __syncwarp();
        if (lane < n_ent) *reinterpret_cast<uint2 *>(smem + map_off + W_MAP + 8 * lane) = make_uint2(ent.x, ent.y - (uint32_t)kt);
    } else {
        t.fix_f0 += kt * p.sps_fixed;
    }
    t.nreg -= kt;
    t.fmap = fmap2;
    __syncwarp();
}

// A tile with more samples than the window holds (T is chosen so that this takes a six-sigma run of long dwells): every
// lane walks k-mers of its own and stores sample by sample.  Unconditionally correct, never fast.
template <bool NOISY, bool RAND_DWELL, bool METH, bool REV>
__device__ __noinline__ void slow_tile(const GenParams &p, const unsigned char *smem, int lane, int tile, int64_t a_off, int64_t b_off,
                                       int32_t a_rem, int32_t nk, uint32_t B, uint32_t S, uint32_t L, int16_t *out, double offset,
                                       uint32_t r_lo, uint32_t r_hi) {
    const RngKey key{p.key0, p.key1, r_lo, r_hi};
    const uint32_t hmul = amp_hmul(r_lo);
    const float c_r = __fsub_rn(SAMPLE_MAGIC, (float)offset);
    const uint16_t *row = reinterpret_cast<const uint16_t *>(p.kpos + (size_t)tile * (TK / 8));
    for (int m = lane; m < nk; m += 32) {
        uint32_t rank = 0;
        for (int i = 0; i < p.k; i++) {
            const int pos = m + i;
            const uint8_t c = base_code(p.bases[(pos < a_rem ? a_off : b_off) + pos]);
            rank = METH ? rank * 5 + (c >> 4) : (rank << 2) | (c & 3);
        }
        uint32_t start, d;
        if (RAND_DWELL) {
            start = row[m];
            d = (m + 1 < nk ? (uint32_t)row[m + 1] : S) - start;
        } else {
            start = (uint32_t)m * (uint32_t)p.sps_fixed;
            d = (uint32_t)p.sps_fixed;
        }
        float A = 0.f, Bq = 0.f;
        uint32_t fixed_v = 0;
        if (NOISY) {
            const float2 am = p.model_am[rank];
            A = am.x;
            Bq = __fsub_rn(__fadd_rn(am.y, c_r), SAMPLE_MAGIC);
        } else {
            fixed_v = to_i16_bits(__dsub_rn(__ddiv_rn(__dmul_rn((double)p.model[rank].x, p.digitisation), p.range), offset));
        }
        for (uint32_t s = 0; s < d; s++) {
            const uint32_t n = B + start + s;
            const uint32_t q = REV ? L - 1 - n : n;
            uint32_t v = fixed_v;
            if (NOISY) {
                const uint32_t Cq = q >> 3;
                uint32_t dw[8];
                amp_draws(p, Cq, r_lo, r_hi, dw);
                uint32_t word = dw[0];
#pragma unroll
                for (int e = 1; e < 8; e++) word = (q & 7u) == (uint32_t)e ? dw[e] : word;
                const uint32_t off = z_offset(word, amp_class4(Cq, hmul));
                float z = *reinterpret_cast<const float *>(smem + SM_Z + off);
                if (z_is_tail(off)) z = z_tail(p.z2, off, q, key, ST_AMP_TAIL);
                v = sample_exact(z, A, Bq);
            }
            out[q] = (int16_t)v;
        }
    }
}

template <bool NOISY, bool RAND_DWELL, bool METH, bool REV>
__global__ void __launch_bounds__(K4_THREADS, 1) signal_kernel(const __grid_constant__ GenParams p) {
    constexpr bool USE_Z = NOISY;
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    unsigned long long *stage_bar = reinterpret_cast<unsigned long long *>(smem + SM_MBAR);
    const uint32_t map_off = SM_WARP + (uint32_t)warp * WARP_BYTES;
    const uint32_t wbase = opaque_smem_addr(smem + map_off);  // this warp's buffer, for the asynchronous copies

    // ---- this warp's range of tiles; its first descriptor ----
    const int64_t gwarp = (int64_t)blockIdx.x * K4_WARPS + warp, nwarp = (int64_t)gridDim.x * K4_WARPS;
    const int lo = (int)(((int64_t)p.n_tiles * gwarp) / nwarp), hi = (int)(((int64_t)p.n_tiles * (gwarp + 1)) / nwarp);
    const bool has_work = lo < hi;
    if (has_work) {
        fetch_tile_desc(p, wbase, W_DESC, lo, lane);
        cp_async_commit();
    }
    // ---- prologue: tables ----
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
    fetch_tile_inputs<RAND_DWELL>(p, smem, wbase, map_off + W_DESC, W_RD, lo, lane);
    fetch_tile_desc(p, wbase, W_DESC + 48, min(lo + 1, p.n_tiles - 1), lane);
    cp_async_commit();

    LaneC lc;
    lc.lw = REV ? 31u - (uint32_t)lane : (uint32_t)lane;
    lc.ent_sh = 8u * (lc.lw & 3u);
    lc.ent_lane = 8u * (lc.lw >> 2);
    lc.lane4 = (uint32_t)lane << 2;

    Run t;
    bool active = false;
    const uint32_t r0_lo = (uint32_t)(uint64_t)p.first_read, r0_hi = (uint32_t)((uint64_t)p.first_read >> 32);

    // ---- main loop: this warp's tiles; the inputs of tile i+1 and the descriptor of tile i+2 fly while tile i's samples are emitted ----
    uint32_t slot = 0;
    for (int tile = lo; tile < hi; tile++, slot ^= 1) {
        cp_async_wait_all();
        __syncwarp();
        const uint32_t desc_off = map_off + W_DESC + slot * 48, rd_off = map_off + W_RD + slot * 32;
        const uint4 d0 = *reinterpret_cast<const uint4 *>(smem + desc_off);
        const uint4 d1 = *reinterpret_cast<const uint4 *>(smem + desc_off + 16);
        const uint4 d2 = *reinterpret_cast<const uint4 *>(smem + desc_off + 32);
        TileIn td;
        td.a_off = (int64_t)(((uint64_t)d0.y << 32) | d0.x);
        td.b_off = (int64_t)(((uint64_t)d0.w << 32) | d0.z);
        td.a_rem = (int32_t)d1.x;
        td.nk = (int32_t)(d1.y & 0xFFFFu);
        td.S = d2.w;
        const uint32_t flags = d1.y >> 16;
        const int32_t read = (int32_t)d1.z;
        const uint32_t B = d2.z;
        const bool rfirst = (flags & TILE_READ_FIRST) != 0, rlast = (flags & TILE_READ_LAST) != 0;
        const bool last = rlast || tile == hi - 1;
        auto prefetch_next = [&]() {
            __syncwarp();   // (every lane is done with this tile's base window and prefix row)
            if (tile + 1 < hi) {
                fetch_tile_inputs<RAND_DWELL>(p, smem, wbase, map_off + W_DESC + (slot ^ 1) * 48, W_RD + (slot ^ 1) * 32, tile + 1, lane);
                fetch_tile_desc(p, wbase, W_DESC + slot * 48, min(tile + 2, p.n_tiles - 1), lane);
                cp_async_commit();
            }
        };

        // ---- does the tile fit behind what the window still holds?  (else the run is cut here: the group in progress is
        // finished by the exact path on both sides of the cut, exactly like a range boundary) ----
        bool cut = false;
        if (active) {
            bool fits = t.nreg + td.nk <= p.par_cap;
            if (RAND_DWELL) fits = fits && ((t.f_end - t.fmap + td.S + 31) >> 5) + 8 <= (uint32_t)MAP_ENT;
            cut = !fits || td.S > p.tile_s_cap;
        }
        const bool slow = td.S > p.tile_s_cap;
        // pass 0 (only after a cut): the run ends with what is registered; pass 1: the tile itself
        for (int pass = cut ? 0 : 1; pass < 2; pass++) {
            bool lst = true;
            uint32_t clip_hi = t.f_end;
            if (pass == 1) {
                if (!active) {
                    // ---- a run begins: the read's constants, the frame, an empty window ----
                    const ReadRec rr = *reinterpret_cast<const ReadRec *>(smem + rd_off);
                    t.L = rr.L;
                    t.offset = rr.offset;
                    t.out = p.sig + rr.sigoff;
                    t.c_r = __fsub_rn(SAMPLE_MAGIC, (float)t.offset);
                    const uint64_t rg = (((uint64_t)r0_hi << 32) | r0_lo) + (uint64_t)(int64_t)read;
                    t.r_lo = (uint32_t)rg; t.r_hi = (uint32_t)(rg >> 32);
                    t.hmul = amp_hmul(t.r_lo);
                    if (!slow) {
                        // frame origin: the unit boundary of the emitted signal at or before the run's first sample
                        const uint32_t delta = REV ? (UNIT_S - (t.L - B) % UNIT_S) % UNIT_S : B % UNIT_S;
                        t.C0 = REV ? ((t.L - B + delta) >> 3) - 1u : (B - delta) >> 3;
                        // a run that does not start its read starts where another warp's range (or a cut) ended: exactly there
                        t.clip_lo = rfirst ? 0u : delta;
                        t.f_end = delta;
                        t.fmap = delta & ~(GROUP_S - 1);
                        t.cur_c = delta >> 3;
                        t.nreg = 0;
                        t.fix_f0 = (int32_t)delta;
                        if (RAND_DWELL)
                            for (int32_t x = lane; x < MAP_ENT + 2; x += 32) *reinterpret_cast<uint2 *>(smem + map_off + W_MAP + 8 * x) = make_uint2(0u, 0xFFFFFFFFu);
                        if (lane == 0) *reinterpret_cast<float2 *>(smem + map_off + W_PARG + 8) = make_float2(0.f, NOISY ? __fsub_rn(SAMPLE_MAGIC, t.c_r) : 0.f);
#ifdef SQG_KO_PHASEA
                        *reinterpret_cast<float2 *>(smem + map_off + W_PAR + 8 * lane) = make_float2(0.f, NOISY ? __fsub_rn(SAMPLE_MAGIC, t.c_r) : 0.f);
#endif
                        active = true;
                        __syncwarp();
                    }
                }
                if (slow) {
                    slow_tile<NOISY, RAND_DWELL, METH, REV>(p, smem, lane, tile, td.a_off, td.b_off, td.a_rem, td.nk, B, td.S, t.L, t.out, t.offset, t.r_lo, t.r_hi);
                    prefetch_next();
                    break;
                }
                register_tile<NOISY, RAND_DWELL, METH, REV>(p, lane, smem, map_off, wbase, t, td);
                prefetch_next();
                t.f_end += td.S;
                lst = last;
                clip_hi = rlast ? 0xFFFFFFFFu : t.f_end;
            }
            emit_ready<NOISY, RAND_DWELL, REV>(p, smem, t, lc, map_off, lane, lst, clip_hi);
            if (lst) active = false;
            else slide_window<RAND_DWELL>(p, lane, smem, map_off, t);
        }
    }
}

}  // namespace sqg
