// sqg_device.cuh — device-side building blocks of the signal path (sm_100a).
//
// Everything here is the CUDA statement of one piece of the reference hot path
// (/root/reference src/gensig.c:226-356, src/seq.h:14-74, src/rand.h:79-94); the arithmetic of the
// Philox mode is specified in DESIGN.md ("Philox mode"); the test suite holds an independent CPU restatement.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace sqg {

// ---- Philox4x32-R (Salmon et al. 2011).  R = 7 rounds: the smallest round count the authors report as
// Crush-resistant (10 is their safety-margin default); every round is 4 issue slots per 8 samples in a kernel that is
// issue-bound, so the margin is not free here.  The oracle uses the same constant (SQO_PHILOX_ROUNDS). ----
constexpr int PHILOX_ROUNDS = 7;

__device__ __forceinline__ uint4 philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < PHILOX_ROUNDS; r++) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0;
        const uint32_t n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}

// Same function with the round keys precomputed on the host (GenParams::rk lives in the constant bank, so
// the key schedule costs no issue slots inside the sample loop)
__device__ __forceinline__ uint4 philox4x32_rk(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, const uint32_t *__restrict__ rk) {
#pragma unroll
    for (int r = 0; r < PHILOX_ROUNDS; r++) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ rk[2 * r];
        const uint32_t n2 = hi0 ^ c3 ^ rk[2 * r + 1];
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    }
    return make_uint4(c0, c1, c2, c3);
}

// counter word 3: which family of draws (DESIGN.md "Philox mode")
enum : uint32_t { ST_AMP = 0, ST_DWELL = 1, ST_READ = 2, ST_AMP_TAIL = 3, ST_DWELL_TAIL = 4, ST_READ_TAIL = 5 };

// ---- the table normal (scripts/make_ztable.py, DESIGN.md "Z32") ----
// Z32[(r << 5) | c]: r = ten random bits (bit 9 = sign), c = class = low five bits of the draw's Philox block number.
// The 4-byte entry's word index has c in its low five bits = the shared-memory bank: the 32 lanes of a warp work on 32
// consecutive blocks, so a warp-wide lookup touches 32 different banks (one wavefront).
constexpr int Z32_N = 32768;
constexpr uint32_t Z32_BYTES = Z32_N * 4;
constexpr uint32_t Z_TAIL_IDX = 0x3FE0u;   // (m = 511) << 5 | class 0: the outermost cell, refined by Z2; Z32 holds NaN there
constexpr int Z2_SUB = 8192;
constexpr float Z_MAX = 5.9104958f;        // largest |z| (bounds the dwell per k-mer and the sample range check)
constexpr uint32_t Z_DRAW_MASK = 0x1FF80u;  // bits 7..16 of a draw word = byte offset of the class-0 entry

struct RngKey {
    uint32_t k0, k1;       // Philox key = seed
    uint32_t r_lo, r_hi;   // global read index = counter words 1,2
};

// the j-th 10-bit draw (j in 0..7) of a Philox block, left in place (bits 7..16): even draws come from word j/2, odd
// draws from the same word rotated by 16 bits.  (word & Z_DRAW_MASK) | (class << 2) is the BYTE offset of the entry.
__device__ __forceinline__ uint32_t draw_word(const uint4 &w, int j) {  // j compile-time after unrolling
    const uint32_t x = (j >> 1) == 0 ? w.x : (j >> 1) == 1 ? w.y : (j >> 1) == 2 ? w.z : w.w;
    return (j & 1) ? __byte_perm(x, x, 0x1032) : x;
}
__device__ __forceinline__ uint32_t z_offset(uint32_t word, uint32_t class4) { return (word & Z_DRAW_MASK) | class4; }
__device__ __forceinline__ bool z_is_tail(uint32_t byte_off) { return ((byte_off >> 2) & 0x3FFFu) == Z_TAIL_IDX; }

// rare path: 13 fresh bits pick the sub-cell of the outermost cell
__device__ __noinline__ float z_tail(const float *__restrict__ z2g, uint32_t byte_off, uint32_t c0, RngKey key, uint32_t stream) {
    const uint4 w = philox4x32(c0, key.r_lo, key.r_hi, stream, key.k0, key.k1);
    const float z = __ldg(z2g + (w.x & (Z2_SUB - 1)));
    return (byte_off & 0x10000u) ? -z : z;  // bit 14 of the index = bit 16 of the byte offset = sign
}

// table normal from global memory (dwell pass, per-read draws)
__device__ __forceinline__ float z_global(const float *__restrict__ z32g, const float *__restrict__ z2g, uint32_t byte_off,
                                          uint32_t tail_c0, RngKey key, uint32_t tail_stream) {
    float z = __ldg(reinterpret_cast<const float *>(reinterpret_cast<const unsigned char *>(z32g) + byte_off));
    if (__builtin_expect(z_is_tail(byte_off), 0)) z = z_tail(z2g, byte_off, tail_c0, key, tail_stream);
    return z;
}

// ---- base -> digit (src/seq.h:14-28 and :45-60).  256-entry table: low nibble = base-4 rank with the
// reference's IUPAC folding, high nibble = base-5 {A,C,G,M,T} rank; everything else ranks 0. ----
__host__ __device__ inline uint8_t base_code(int c) {
    uint8_t r4 = 0, r5 = 0;
    switch (c) {
        case 'C': case 'c': case 'Y': case 'B': r4 = 1; break;
        case 'G': case 'g': case 'S': case 'K': r4 = 2; break;
        case 'T': case 't': case 'U': r4 = 3; break;
        default: r4 = 0;
    }
    switch (c) {
        case 'C': r5 = 1; break;
        case 'G': r5 = 2; break;
        case 'M': r5 = 3; break;
        case 'T': r5 = 4; break;
        default: r5 = 0;
    }
    return (uint8_t)(r4 | (r5 << 4));
}

// double/float -> int16 as the reference's `raw_signal[n] = <double>` store does on x86-64
// (src/gensig.c:270): truncate toward zero to int32, keep the low 16 bits, no clamp.
__device__ __forceinline__ uint32_t to_i16_bits(double v) { return (uint32_t)__double2int_rz(v) & 0xFFFFu; }
__device__ __forceinline__ uint32_t to_i16_bits(float v) { return (uint32_t)__float2int_rz(v) & 0xFFFFu; }

// Sample arithmetic of the Philox mode: the single-precision FMA z*A + Bq ROUNDED TOWARD ZERO, then truncated.
// SAMPLE_MAGIC: Bm = Bq + 32768 puts the sum into [32768, 65536), where floats are spaced 2^-8 apart, so for
// 0 <= value < 32768 bits 8..23 of fma.rz(z, A, Bm) ARE the truncated sample (bit 23, the exponent's low bit, is 0
// exactly in that binade): no float->int conversion, and a set bit 23 flags everything else (negative values, values
// beyond int16, the NaN of a tail cell) for the exact path below.
constexpr float SAMPLE_MAGIC = 32768.0f;
__device__ __forceinline__ float fma_rz(float a, float b, float c) { return __fmaf_rz(a, b, c); }
__device__ __forceinline__ uint32_t sample_exact(float z, float A, float Bq) { return to_i16_bits(fma_rz(z, A, Bq)); }

// dwell of one k-mer from a table normal (src/gensig.c:255-256): Philox mode rounds the single-precision
// FMA to nearest (ties to even) and folds values below 1 exactly like the reference
__device__ __forceinline__ int dwell_from_z(float z, float dwell_mean, float dwell_std) {
    const int d = __float2int_rn(fmaf(z, dwell_std, dwell_mean));
    return d < 1 ? -d + 1 : d;
}

}  // namespace sqg
