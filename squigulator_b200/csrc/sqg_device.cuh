// sqg_device.cuh — device-side building blocks of the signal path (sm_100a).
//
// Everything here is the CUDA statement of one piece of the reference hot path
// (/root/reference src/gensig.c:226-356, src/seq.h:14-74, src/rand.h:79-94); the arithmetic of the
// Philox mode is specified in DESIGN.md ("Philox mode"); the test suite holds an independent CPU restatement.
#pragma once
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

namespace sqg {

// ---- Philox4x32-10 (Salmon et al. 2011); round keys are warp-uniform so they live in uniform registers ----
__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                               uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
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

// Same function with the ten round keys precomputed on the host (GenParams::rk lives in the constant bank, so
// the key schedule costs no issue slots inside the sample loop)
__device__ __forceinline__ uint4 philox4x32_10_rk(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                  const uint32_t *__restrict__ rk) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
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

constexpr int Z16_N = 65536;          // Z16[h]: bit 15 of h = sign, bits 0-14 = half-normal cell of probability 2^-16
constexpr int Z_TAIL_FIRST = 32766;   // the 2 outermost cells (p = 2^-14) are refined ...
constexpr int Z2_SUB = 8192;          // ... into 8192 sub-cells each (Z2, float32)
constexpr int Z2_N = 2 * Z2_SUB;
constexpr float Z_MAX = 6.0590086f;   // largest table entry (bounds the dwell per k-mer)
constexpr float Z_TAIL_THR = 4.08203125f;  // fp16(Z1[32766]): |z| >= this <=> possibly a tail cell

struct RngKey {
    uint32_t k0, k1;       // Philox key = seed
    uint32_t r_lo, r_hi;   // global read index = counter words 1,2
};

// Bank-stratified table index: bits 1-5 of a 16-bit draw (= the shared-memory bank of its 2-byte table entry) are
// replaced by the low five bits of the draw's Philox block number.  The 32 lanes of a warp work on 32 consecutive
// blocks, so a warp-wide table lookup touches 32 different banks: one wavefront instead of ~3.4.  Each draw still
// picks uniformly among 2^11 cells spread evenly over the table, and all cells are used across block residues.
__device__ __forceinline__ uint32_t stratify(uint32_t h, uint32_t block) { return (h & 0xFFC1u) | ((block & 31u) << 1); }

// rare path of z16: 13 fresh bits pick the sub-cell
__device__ __noinline__ float z16_tail(const float *__restrict__ z2g, uint32_t h, uint32_t c0, RngKey key,
                                       uint32_t stream) {
    const uint4 w = philox4x32_10(c0, key.r_lo, key.r_hi, stream, key.k0, key.k1);
    const float z = __ldg(z2g + ((h & 0x7FFFu) - Z_TAIL_FIRST) * Z2_SUB + (w.x & (Z2_SUB - 1)));
    return (h & 0x8000u) ? -z : z;
}

// 16-bit uniform -> N(0,1): one lookup in the signed binary16 quantile table (global or shared memory)
__device__ __forceinline__ float z16(const __half *__restrict__ z16t, const float *__restrict__ z2g, uint32_t h,
                                     uint32_t tail_c0, RngKey key, uint32_t tail_stream) {
    float z = __half2float(z16t[h]);
    if (__builtin_expect((h & 0x7FFFu) >= Z_TAIL_FIRST, 0)) z = z16_tail(z2g, h, tail_c0, key, tail_stream);
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

// dwell of one k-mer from a table normal (src/gensig.c:255-256): Philox mode rounds the single-precision
// FMA to nearest (ties to even) and folds values below 1 exactly like the reference
__device__ __forceinline__ int dwell_from_z(float z, float dwell_mean, float dwell_std) {
    const int d = __float2int_rn(fmaf(z, dwell_std, dwell_mean));
    return d < 1 ? -d + 1 : d;
}

}  // namespace sqg
