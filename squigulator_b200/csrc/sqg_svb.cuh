// sqg_svb.cuh — svb-zd on the GPU: every read's int16 signal re-coded, in HBM, as the byte stream
// slow5lib's ptr_compress_svb_zd produces (slow5lib/src/slow5_press.c:1055-1087: uint32 sample count, then StreamVByte
// of the zig-zag deltas - ceil(n/4) key bytes with 2 bits per value, then 1-4 little-endian data bytes per value,
// thirdparty/streamvbyte/src/streamvbyte_encode.c:31-79, streamvbyte_zigzag.c:15-27), so that the device->host copy
// moves ~1.3 bytes per sample instead of 2 and the caller's BLOW5 writer can skip its own signal compression.
//
//   S1 svb_size_kernel    one CTA per read: bytes of its stream                              -> svb_len
//   S2 svb_offsets_kernel one CTA        : exclusive scan of the 16-byte-aligned lengths     -> svb_off, total
//   S3 svb_encode_kernel  one CTA per read: keys + data, strip by strip with a running carry -> svb
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace sqg {

struct SvbParams {
    const int16_t *sig;          // signal arena
    const int64_t *read_sigoff;  // start of read r in the arena (multiple of 64 samples)
    const uint32_t *read_siglen;
    int64_t *svb_len;            // per read: bytes of its stream
    int64_t *svb_off;            // n_reads + 1: start of read r's stream in svb (16-byte aligned); [n_reads] = total
    uint8_t *svb;
    int32_t n_reads;
};

constexpr int SVB_THREADS = 256;
constexpr int SVB_STRIP = SVB_THREADS * 8;   // samples per strip: 8 consecutive samples per thread

// zig-zag deltas of 8 consecutive samples starting at i0 (multiple of 8); values at or beyond n come out as 0 bytes
__device__ __forceinline__ void svb_load8(const int16_t *x, uint32_t i0, uint32_t n, uint32_t (&v)[8], uint32_t (&nb)[8]) {
    const uint4 q = *reinterpret_cast<const uint4 *>(x + i0);   // reads start on 128-byte boundaries and are padded to them
    int32_t prev = i0 ? (int32_t)x[i0 - 1] : 0;
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const int32_t s = (int32_t)(int16_t)((j & 1) ? (w[j >> 1] >> 16) : (w[j >> 1] & 0xFFFFu));
        const int32_t d = s - prev;
        prev = s;
        v[j] = ((uint32_t)d + (uint32_t)d) ^ (uint32_t)(d >> 31);
        nb[j] = (i0 + j < n) ? 1u + (v[j] >= 256u) + (v[j] >= 65536u) : 0u;   // |delta| < 2^16: at most 3 bytes
    }
}

__device__ __forceinline__ uint32_t svb_block_sum(uint32_t x, uint32_t *s_warp) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = x;
    __syncthreads();
    uint32_t t = 0;
#pragma unroll
    for (int w = 0; w < SVB_THREADS / 32; w++) t += s_warp[w];
    return t;
}

__global__ void __launch_bounds__(SVB_THREADS) svb_size_kernel(const SvbParams p) {
    __shared__ uint32_t s_warp[SVB_THREADS / 32];
    const int r = blockIdx.x;
    const uint32_t n = p.read_siglen[r];
    const int16_t *x = p.sig + p.read_sigoff[r];
    uint32_t bytes = 0;
    for (uint32_t i0 = threadIdx.x * 8; i0 < n; i0 += SVB_STRIP) {
        uint32_t v[8], nb[8];
        svb_load8(x, i0, n, v, nb);
#pragma unroll
        for (int j = 0; j < 8; j++) bytes += nb[j];
    }
    const uint32_t total = svb_block_sum(bytes, s_warp);
    if (threadIdx.x == 0) p.svb_len[r] = 4 + (int64_t)((n + 3) / 4) + total;
}

__global__ void __launch_bounds__(1024) svb_offsets_kernel(const SvbParams p) {
    __shared__ uint64_t s_warp[32];
    const int tid = threadIdx.x;
    const int per = (p.n_reads + 1023) / 1024;
    const int lo = min(p.n_reads, tid * per), hi = min(p.n_reads, lo + per);
    uint64_t part = 0;
    for (int r = lo; r < hi; r++) part += ((uint64_t)p.svb_len[r] + 15) & ~15ull;
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
        if (tid == 31) p.svb_off[p.n_reads] = (int64_t)winc;
    }
    __syncthreads();
    uint64_t base = s_warp[tid >> 5] + inc - part;
    for (int r = lo; r < hi; r++) {
        p.svb_off[r] = (int64_t)base;
        base += ((uint64_t)p.svb_len[r] + 15) & ~15ull;
    }
}

__global__ void __launch_bounds__(SVB_THREADS) svb_encode_kernel(const SvbParams p) {
    __shared__ uint32_t s_warp[SVB_THREADS / 32];
    __shared__ uint32_t s_carry;
    const int r = blockIdx.x;
    const uint32_t n = p.read_siglen[r];
    const int16_t *x = p.sig + p.read_sigoff[r];
    uint8_t *out = p.svb + p.svb_off[r];
    uint8_t *keys = out + 4;
    uint8_t *data = keys + (n + 3) / 4;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        *reinterpret_cast<uint32_t *>(out) = n;   // slow5_press.c:1047: the original length, needed for depress
        s_carry = 0;
    }
    __syncthreads();
    for (uint32_t base = 0; base < n; base += SVB_STRIP) {
        const uint32_t i0 = base + threadIdx.x * 8;
        uint32_t v[8], nb[8], mine = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) { v[j] = 0; nb[j] = 0; }
        if (i0 < n) svb_load8(x, i0, n, v, nb);
#pragma unroll
        for (int j = 0; j < 8; j++) mine += nb[j];
        // exclusive scan of the threads' byte counts within the strip
        uint32_t inc = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) s_warp[warp] = inc;
        __syncthreads();
        uint32_t before = s_carry, strip_total = 0;
#pragma unroll
        for (int w = 0; w < SVB_THREADS / 32; w++) {
            if (w < warp) before += s_warp[w];
            strip_total += s_warp[w];
        }
        if (i0 < n) {
            // two key bytes: four 2-bit codes each, first value in the low bits (a partial last byte keeps zeros)
            uint32_t k0 = 0, k1 = 0;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                k0 |= (nb[j] ? nb[j] - 1 : 0u) << (2 * j);
                k1 |= (nb[4 + j] ? nb[4 + j] - 1 : 0u) << (2 * j);
            }
            keys[i0 >> 2] = (uint8_t)k0;
            if (i0 + 4 < n) keys[(i0 >> 2) + 1] = (uint8_t)k1;
            uint8_t *d = data + before + (inc - mine);
#pragma unroll
            for (int j = 0; j < 8; j++) {
                if (nb[j] > 0) d[0] = (uint8_t)v[j];
                if (nb[j] > 1) d[1] = (uint8_t)(v[j] >> 8);
                if (nb[j] > 2) d[2] = (uint8_t)(v[j] >> 16);
                d += nb[j];
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) s_carry += strip_total;
        __syncthreads();
    }
}

}  // namespace sqg
