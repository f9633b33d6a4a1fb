// sqg_extract.cuh — reads named by coordinates, cut out of a genome that lives in HBM.
//
// What the reference does per accepted read on the host (src/genread.c:243-281, gen_read_dna; the RNA path uses the
// first two steps only):
//   1. copy `len` characters of the contig from `pos`                               (gen_read_common, :149-153)
//   2. replace every 'N' by "ACGT"[round(3*u)], u from a minstd stream that starts   (is_bad_read, :132-140)
//      at 100 for every read
//   3. '-' strand: reverse, complement (anything that is not ACGTacgt becomes 'T')   (src/seq.h:77-112)
//   4. methylation: for every forward position i with contig[pos+i] == 'C' and       (methylate_dna, :207-241)
//      contig[pos+i+1] == 'G' (i+1 < len), take the next value u of the rand_meth
//      stream; if (int)(254*u) <= meth[pos+i] the C of that site in the READ becomes
//      'M' (forward: index i; reverse: index len-i-2)
// minstd's n-th value is seed*16807^n mod (2^31-1), so both streams are addressed by ordinal: the N's of a read are
// numbered by ballot/popcount as a warp walks the read, the CpG sites of a batch by a per-read count, an exclusive
// scan over the reads, and the same ballot numbering inside the read.  One warp per read, 32 positions per step;
// every output byte is written by the lane that owns its forward position.
#pragma once
#include <cub/cub.cuh>

#include "sqg_legacy.cuh"

namespace sqg {

struct Coord {  // == sqg_coord_t
    int32_t contig, len;
    int64_t pos;
    int32_t strand, reserved;
};

struct ExtractParams {
    const uint8_t *genome;
    const int64_t *contig_off;
    const uint8_t *meth;             // nullable
    const uint8_t *contig_has_meth;  // nullable (= all)
    const Coord *coords;
    const int64_t *out_off;  // n_reads + 1, relative to `out`
    uint8_t *out;
    uint64_t *cg_count;      // per read (do_meth)
    uint64_t *cg_off;        // exclusive scan of cg_count
    uint32_t meth_residue;   // (seed + 6) mod m
    uint64_t meth_draw_base;
    int32_t n_reads, do_meth;
    uint32_t pw[33];         // 16807^j mod m, j = 0..32
};

constexpr int EX_THREADS = 256;

__device__ __forceinline__ bool read_has_meth(const ExtractParams &q, const Coord &c) {
    return q.do_meth && q.meth && (!q.contig_has_meth || q.contig_has_meth[c.contig]);
}

// CpG sites per read (the draws methylate_dna will take)
__global__ void __launch_bounds__(EX_THREADS) extract_count_kernel(const __grid_constant__ ExtractParams q) {
    const int r = blockIdx.x * (EX_THREADS / 32) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= q.n_reads) return;
    const Coord c = q.coords[r];
    uint32_t n = 0;
    if (read_has_meth(q, c)) {
        const uint8_t *g = q.genome + q.contig_off[c.contig] + c.pos;
        for (int i = lane; i + 1 < c.len; i += 32) n += (g[i] == 'C' && g[i + 1] == 'G');
    }
    n = __reduce_add_sync(0xFFFFFFFFu, n);
    if (lane == 0) q.cg_count[r] = n;
}

__device__ __forceinline__ uint8_t complement_base(uint8_t b) {  // src/seq.h:77-101
    switch (b) {
        case 'A': case 'a': return 'T';
        case 'C': case 'c': return 'G';
        case 'G': case 'g': return 'C';
        case 'T': case 't': return 'A';
        default: return 'T';
    }
}

__device__ __forceinline__ double lehmer_unit(uint32_t x) { return (double)(x ? x : LEHMER_M) / 2147483647; }

__global__ void __launch_bounds__(EX_THREADS) extract_reads_kernel(const __grid_constant__ ExtractParams q) {
    __shared__ uint32_t pw[33];
    if (threadIdx.x < 33) pw[threadIdx.x] = q.pw[threadIdx.x];
    __syncthreads();
    const int r = blockIdx.x * (EX_THREADS / 32) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= q.n_reads) return;
    const Coord c = q.coords[r];
    const int len = c.len;
    const uint8_t *g = q.genome + q.contig_off[c.contig] + c.pos;
    const uint8_t *mt = q.meth ? q.meth + q.contig_off[c.contig] + c.pos : nullptr;
    uint8_t *out = q.out + q.out_off[r];
    const bool neg = c.strand == '-';
    const bool meth = read_has_meth(q, c);
    uint32_t xn = 100;  // state of the read's N stream (residue after the N's seen so far)
    uint32_t xm = 0;    // state of rand_meth before this read's first site
    if (meth) xm = mulmod31(q.meth_residue, powmod31(LEHMER_A, q.meth_draw_base + q.cg_off[r]));
    const uint32_t lt = (1u << lane) - 1;
    bool carry = false;  // '-' reads: the site at the previous step's last position was marked
    for (int i0 = 0; i0 < len; i0 += 32) {
        const int i = i0 + lane;
        const bool valid = i < len;
        const uint8_t raw = valid ? g[i] : 0;
        uint32_t nxt = __shfl_down_sync(0xFFFFFFFFu, (uint32_t)raw, 1);
        if (lane == 31) nxt = (i + 1 < len) ? g[i + 1] : 0;
        uint8_t b = raw;
        const bool is_n = raw == 'N';
        const uint32_t mn = __ballot_sync(0xFFFFFFFFu, is_n);
        if (mn) {
            if (is_n) {
                const uint32_t x = mulmod31(xn, pw[__popc(mn & lt) + 1]);
                const int n = (int)round(lehmer_unit(x) * 3);
                b = n == 0 ? 'A' : n == 1 ? 'C' : n == 2 ? 'G' : 'T';
            }
            xn = mulmod31(xn, pw[__popc(mn)]);
        }
        bool mark = false;
        if (meth) {
            const bool site = valid && i + 1 < len && raw == 'C' && nxt == 'G';
            const uint32_t mc = __ballot_sync(0xFFFFFFFFu, site);
            if (mc) {
                if (site) {
                    const uint32_t x = mulmod31(xm, pw[__popc(mc & lt) + 1]);
                    const int methr = (int)(lehmer_unit(x) * 254);
                    mark = methr <= (int)mt[i];
                }
                xm = mulmod31(xm, pw[__popc(mc)]);
            }
        }
        if (!neg) {
            if (valid) out[i] = mark ? 'M' : b;
        } else {
            // the marked C of a site at forward position i is the complement of the G at i+1: owned by the next lane
            uint32_t prev = __shfl_up_sync(0xFFFFFFFFu, (uint32_t)mark, 1);
            if (lane == 0) prev = carry;
            carry = __shfl_sync(0xFFFFFFFFu, (uint32_t)mark, 31) != 0;
            if (valid) out[len - 1 - i] = prev ? 'M' : complement_base(b);
        }
    }
}

}  // namespace sqg
