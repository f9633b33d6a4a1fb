// sqg_legacy.cuh — SQG_RNG_LEGACY: the reference's own random streams, reproduced on the GPU.
//
// The reference draws every random number from minstd Lehmer streams (src/rand.h:79-94) laid out per thread in
// init_rand() (src/sim.c:215-258): for thread 0, dwell = seed+2, offset = seed+4, median_before = seed+5, and ONE
// stream per k-mer rank seeded seed+rank whose state persists across reads.  The n-th output of a stream is
// seed*16807^n mod (2^31-1), so any draw can be computed directly once its POSITION in its stream is known:
//   dwell of the g-th k-mer generated since start-up      -> steps 2g+1, 2g+2 of the dwell stream
//   offset / median_before of the r-th read               -> steps 2r+1, 2r+2 of their streams
//   j-th sample of a k-mer of rank q                      -> steps 2(c+j)+1, 2(c+j)+2 of stream q, where c = samples
//                                                            drawn from stream q by all earlier k-mers (in read
//                                                            order, then k-mer order) since start-up
// c is obtained per batch with a stable radix sort of the k-mers by rank and a segmented exclusive scan of their
// dwells (CUB), plus a per-rank carry kept in device memory between batches.  With that, `squigulator -t1` output is
// reproduced bit for bit for the same read order (tests/test_gpu_golden.py runs the reference's golden files through
// this path).  Parity mode: simple kernels, not tuned; the Philox mode is the fast path.
#pragma once
#include <cub/cub.cuh>

#include "sqg_kernels.cuh"

namespace sqg {

constexpr uint32_t LEHMER_M = 2147483647u;
constexpr uint32_t LEHMER_A = 16807u;

__host__ __device__ inline uint32_t mulmod31(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) % LEHMER_M); }
__host__ __device__ inline uint32_t powmod31(uint32_t a, uint64_t e) {
    uint32_t r = 1;
    while (e) {
        if (e & 1) r = mulmod31(r, a);
        a = mulmod31(a, a);
        e >>= 1;
    }
    return r;
}
// stream seed -> residue (the reference's first Schrage step maps any non-negative int64 seed to seed*16807 mod m)
__host__ __device__ inline uint32_t seed_residue(int64_t s) { return (uint32_t)((uint64_t)s % LEHMER_M); }

// one nrng() draw (src/rand.h:87-94) from residue r = state BEFORE the draw; advances r by two steps
__device__ __forceinline__ double legacy_normal(uint32_t &r, double mean, double sd) {
    r = mulmod31(r, LEHMER_A);
    const double u = (double)(r ? r : LEHMER_M) / 2147483647;
    r = mulmod31(r, LEHMER_A);
    const double t = 2.0 * 3.14159265 * ((double)(r ? r : LEHMER_M) / 2147483647);
    const double x = __dmul_rn(sqrt(__dmul_rn(-2.0, log(u))), cos(t));
    return __dadd_rn(__dmul_rn(x, sd), mean);
}

struct LegacyParams {
    int64_t seed;
    uint64_t dwell_pos0;   // k-mers whose dwell was drawn before this batch
    uint64_t read_pos0;    // reads generated before this batch
    uint64_t *cnt_kmer;    // per rank: samples drawn so far (carry between batches)
    uint32_t *kmer_rank;   // per k-mer of the batch
    uint64_t *kmer_cpos;   // per k-mer of the batch: position of its first sample in its rank's stream
    int32_t noisy, rand_dwell, meth, rev;
    double dwell_mean, dwell_std;
};

// dwell per k-mer (src/gensig.c:255-256) + tile sums.  One warp per tile, 8 consecutive k-mers per lane.
__global__ void __launch_bounds__(128) legacy_dwell_kernel(const __grid_constant__ GenParams p, const __grid_constant__ LegacyParams q) {
    const int tile = blockIdx.x * 4 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (tile >= p.n_tiles) return;
    const TileDesc td = p.tiles[tile];
    uint32_t sum = 0;
    const int m0 = lane * 8;
    if (m0 < td.nk) {
        const uint64_t g0 = q.dwell_pos0 + (uint64_t)td.ss_pos + m0;
        uint32_t r = mulmod31(seed_residue(q.seed + 2), powmod31(LEHMER_A, 2 * g0));
        for (int j = 0; j < 8 && m0 + j < td.nk; j++) {
            int d = p.sps_fixed;
            if (q.rand_dwell) {
                d = (int)round(legacy_normal(r, q.dwell_mean, q.dwell_std));
                if (d < 1) d = -d + 1;
            }
            p.ss[td.ss_pos + m0 + j] = d;
            sum += (uint32_t)d;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (lane == 0) p.tile_sum[tile] = sum;
}

// per-read offset / median_before (src/gensig.c:312-318)
__global__ void __launch_bounds__(256) legacy_read_draws_kernel(const __grid_constant__ GenParams p, const __grid_constant__ LegacyParams q) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= p.n_reads || p.ideal) return;
    const uint64_t n = q.read_pos0 + (uint64_t)r;
    uint32_t a = mulmod31(seed_residue(q.seed + 4), powmod31(LEHMER_A, 2 * n));
    uint32_t b = mulmod31(seed_residue(q.seed + 5), powmod31(LEHMER_A, 2 * n));
    p.read_offset[r] = legacy_normal(a, p.offset_mean, p.offset_std);
    p.read_median[r] = legacy_normal(b, p.median_mean, p.median_std);
}

// rank of every k-mer (src/seq.h:31-42 / :62-74).  One thread per k-mer.
__global__ void __launch_bounds__(256) legacy_rank_kernel(const __grid_constant__ GenParams p, const __grid_constant__ LegacyParams q) {
    const int tile = blockIdx.x;
    const TileDesc td = p.tiles[tile];
    for (int m = threadIdx.x; m < td.nk; m += blockDim.x) {
        uint32_t rank = 0;
        for (int i = 0; i < p.k; i++) {
            const int pos = m + i;
            const uint8_t c = base_code(p.bases[(pos < td.a_rem ? td.a_off : td.b_off) + pos]);
            rank = q.meth ? rank * 5 + (c >> 4) : (rank << 2) | (c & 3);
        }
        q.kmer_rank[td.ss_pos + m] = rank;
    }
}

__global__ void legacy_iota_kernel(uint32_t *idx, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) idx[i] = (uint32_t)i;
}
__global__ void legacy_gather_dwell_kernel(const int32_t *ss, const uint32_t *idx_sorted, uint64_t *dsorted, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dsorted[i] = (uint64_t)ss[idx_sorted[i]];
}
__global__ void legacy_heads_kernel(const uint32_t *rank_sorted, const uint64_t *excl, uint64_t *heads, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) heads[i] = (i == 0 || rank_sorted[i] != rank_sorted[i - 1]) ? excl[i] : 0;
}
__global__ void legacy_cpos_kernel(const uint32_t *rank_sorted, const uint32_t *idx_sorted, const uint64_t *excl,
                                   const uint64_t *seg_start, const uint64_t *cnt_kmer, uint64_t *cpos, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) cpos[idx_sorted[i]] = cnt_kmer[rank_sorted[i]] + excl[i] - seg_start[i];
}
__global__ void legacy_carry_kernel(const uint32_t *rank_sorted, const uint64_t *excl, const uint64_t *seg_start,
                                    const uint64_t *dsorted, uint64_t *cnt_kmer, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && (i == n - 1 || rank_sorted[i + 1] != rank_sorted[i])) cnt_kmer[rank_sorted[i]] += excl[i] + dsorted[i] - seg_start[i];
}

// samples (src/gensig.c:259-272, :348-354).  One warp per tile, 8 consecutive k-mers per lane, scalar int16 stores.
__global__ void __launch_bounds__(128) legacy_signal_kernel(const __grid_constant__ GenParams p, const __grid_constant__ LegacyParams q) {
    const int tile = blockIdx.x * 4 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (tile >= p.n_tiles) return;
    const TileDesc td = p.tiles[tile];
    const uint32_t L = p.read_siglen[td.read];
    const double offset = p.read_offset[td.read];
    int16_t *out = p.sig + p.read_sigoff[td.read];
    const int m0 = lane * 8;
    uint32_t local = 0;
    for (int j = 0; j < 8 && m0 + j < td.nk; j++) local += (uint32_t)p.ss[td.ss_pos + m0 + j];
    uint32_t inc = local;
#pragma unroll
    for (int sh = 1; sh < 32; sh <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, inc, sh);
        if (lane >= sh) inc += v;
    }
    uint32_t n = td.B + inc - local;  // logical sample index within the read
    for (int j = 0; j < 8 && m0 + j < td.nk; j++) {
        const int64_t ks = td.ss_pos + m0 + j;
        const int d = p.ss[ks];
        const uint32_t rank = q.kmer_rank[ks];
        const float2 mv = p.model[rank];
        if (!q.noisy) {
            // src/gensig.c:266,270
            const int16_t v = (int16_t)to_i16_bits(__dsub_rn(__ddiv_rn(__dmul_rn((double)mv.x, p.digitisation), p.range), offset));
            for (int s = 0; s < d; s++, n++) out[q.rev ? L - 1 - n : n] = v;
        } else {
            const double sd = (double)__fmul_rn(mv.y, p.amp_noise);  // float product widened, src/sim.c:249
            uint32_t r = mulmod31(seed_residue(q.seed + (int64_t)rank), powmod31(LEHMER_A, 2 * q.kmer_cpos[ks]));
            for (int s = 0; s < d; s++, n++) {
                const float smp = (float)legacy_normal(r, (double)mv.x, sd);  // src/gensig.c:268: narrowed to float
                out[q.rev ? L - 1 - n : n] =
                    (int16_t)to_i16_bits(__dsub_rn(__ddiv_rn(__dmul_rn((double)smp, p.digitisation), p.range), offset));
            }
        }
    }
}

}  // namespace sqg
