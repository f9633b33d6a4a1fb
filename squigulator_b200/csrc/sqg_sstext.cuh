// sqg_sstext.cuh — the `ss:Z:` dwell string of PAF/SAM records on the GPU.
//
// The reference appends one "%d," per k-mer with vsnprintf (src/format.c:69-75 for PAF, :114-118 for SAM; the k-mers
// in reverse order for RNA, where t_st > t_end).  Here the text of every read is produced in HBM - per read
// "d0,d1,...,dn-1," without terminator - and copied out instead of the int32 array (~3.3 bytes per k-mer instead of 4,
// and no per-k-mer formatting on the host).
//   T1 sstext_size_kernel    one CTA per read: characters of its string        -> len
//   T2 sstext_offsets_kernel one CTA        : exclusive scan                    -> off, total
//   T3 sstext_write_kernel   one CTA per read: digits and commas
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace sqg {

struct SsTextParams {
    const int32_t *ss;       // dwell per k-mer, read r at ss[ss_off[r] .. ss_off[r+1])
    const int64_t *ss_off;   // n_reads + 1
    int64_t *len;            // per read: characters
    int64_t *off;            // n_reads + 1
    char *text;
    int32_t n_reads;
    int32_t reversed;        // RNA: last k-mer first (src/format.c:70-73)
};

constexpr int SST_THREADS = 256;

__device__ __forceinline__ uint32_t dec_digits(uint32_t v) {
    return 1u + (v >= 10u) + (v >= 100u) + (v >= 1000u) + (v >= 10000u) + (v >= 100000u) + (v >= 1000000u) + (v >= 10000000u) +
           (v >= 100000000u) + (v >= 1000000000u);
}

__global__ void __launch_bounds__(SST_THREADS) sstext_size_kernel(const SsTextParams p) {
    __shared__ uint32_t s_warp[SST_THREADS / 32];
    const int r = blockIdx.x;
    const int64_t a = p.ss_off[r], n = p.ss_off[r + 1] - a;
    uint32_t mine = 0;
    for (int64_t i = threadIdx.x; i < n; i += SST_THREADS) mine += dec_digits((uint32_t)p.ss[a + i]) + 1u;   // dwells are >= 1
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
    if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = mine;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint64_t t = 0;
        for (int w = 0; w < SST_THREADS / 32; w++) t += s_warp[w];
        p.len[r] = (int64_t)t;
    }
}

__global__ void __launch_bounds__(1024) sstext_offsets_kernel(const SsTextParams p) {
    __shared__ uint64_t s_warp[32];
    const int tid = threadIdx.x;
    const int per = (p.n_reads + 1023) / 1024;
    const int lo = min(p.n_reads, tid * per), hi = min(p.n_reads, lo + per);
    uint64_t part = 0;
    for (int r = lo; r < hi; r++) part += (uint64_t)p.len[r];
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
        if (tid == 31) p.off[p.n_reads] = (int64_t)winc;
    }
    __syncthreads();
    uint64_t base = s_warp[tid >> 5] + inc - part;
    for (int r = lo; r < hi; r++) {
        p.off[r] = (int64_t)base;
        base += (uint64_t)p.len[r];
    }
}

__global__ void __launch_bounds__(SST_THREADS) sstext_write_kernel(const SsTextParams p) {
    __shared__ uint32_t s_warp[SST_THREADS / 32];
    __shared__ uint64_t s_carry;
    const int r = blockIdx.x;
    const int64_t a = p.ss_off[r], n = p.ss_off[r + 1] - a;
    char *out = p.text + p.off[r];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int64_t base = 0; base < n; base += SST_THREADS * 4) {
        // thread = 4 consecutive k-mers of the OUTPUT order
        uint32_t v[4], len[4], mine = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int64_t i = base + threadIdx.x * 4 + j;
            v[j] = 0; len[j] = 0;
            if (i < n) {
                v[j] = (uint32_t)p.ss[a + (p.reversed ? n - 1 - i : i)];
                len[j] = dec_digits(v[j]) + 1u;
            }
            mine += len[j];
        }
        uint32_t inc = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) s_warp[warp] = inc;
        __syncthreads();
        uint64_t before = s_carry;
        uint32_t strip_total = 0;
#pragma unroll
        for (int w = 0; w < SST_THREADS / 32; w++) {
            if (w < warp) before += s_warp[w];
            strip_total += s_warp[w];
        }
        char *d = out + before + (inc - mine);
#pragma unroll
        for (int j = 0; j < 4; j++) {
            if (len[j]) {
                uint32_t x = v[j];
                d[len[j] - 1] = ',';
                for (int c = (int)len[j] - 2; c >= 0; c--) { d[c] = (char)('0' + x % 10u); x /= 10u; }
                d += len[j];
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) s_carry += strip_total;
        __syncthreads();
    }
}

}  // namespace sqg
