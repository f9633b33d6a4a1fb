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
// minstd's n-th value is seed*16807^n mod (2^31-1), so both streams are addressed by ordinal.  Reads are cut into
// pieces of XSEG forward positions, one warp per piece: a count kernel gives every piece its number of N's and of CpG
// sites, one exclusive scan over the pieces of the batch turns the counts into ordinals (N's restart at every read,
// CpG sites run through the batch), and inside a piece ballot/popcount numbers the hits of each 32-position row.
// Every output byte is written by the lane that owns its forward position; four rows of loads are in flight per lane.
#pragma once
#include <cub/cub.cuh>

#include "sqg_legacy.cuh"

namespace sqg {

struct Coord {  // == sqg_coord_t
    int32_t contig, len;
    int64_t pos;
    int32_t strand, reserved;
};

struct PieceRef {  // piece k of read `read`: forward positions [k*XSEG, min(len, (k+1)*XSEG))
    int32_t read, k;
};

constexpr int XSEG = 2048;
constexpr int EX_THREADS = 256;
constexpr int EX_ROWS = 4;  // 32-position rows per loop step

struct ExtractParams {
    const uint8_t *genome;
    const int64_t *contig_off;
    const uint8_t *meth;             // nullable
    const uint8_t *contig_has_meth;  // nullable (= all)
    const Coord *coords;
    const PieceRef *pieces;
    const int64_t *out_off;  // n_reads + 1, relative to `out`
    uint8_t *out;
    uint64_t *cnt;           // per piece: (N's << 32) | CpG sites
    uint64_t *cnt_off;       // exclusive scan of cnt over the batch's pieces
    uint32_t meth_residue;   // (seed + 6) mod m
    uint64_t meth_draw_base;
    int32_t n_pieces, do_meth;
    uint32_t pw[33];         // 16807^j mod m, j = 0..32
};

__device__ __forceinline__ bool read_has_meth(const ExtractParams &q, const Coord &c) {
    return q.do_meth && q.meth && (!q.contig_has_meth || q.contig_has_meth[c.contig]);
}

// N's and CpG sites per piece (the draws is_bad_read / methylate_dna will take there)
__global__ void __launch_bounds__(EX_THREADS) extract_count_kernel(const __grid_constant__ ExtractParams q) {
    const int s = blockIdx.x * (EX_THREADS / 32) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (s >= q.n_pieces) return;
    const PieceRef pr = q.pieces[s];
    const Coord c = q.coords[pr.read];
    const int lo = pr.k * XSEG, hi = min(c.len, lo + XSEG);
    const uint8_t *g = q.genome + q.contig_off[c.contig] + c.pos;
    const bool meth = read_has_meth(q, c);
    uint32_t nn = 0, nc = 0;
    for (int i0 = lo; i0 < hi; i0 += 32 * EX_ROWS) {
        uint8_t raw[EX_ROWS];
#pragma unroll
        for (int j = 0; j < EX_ROWS; j++) {
            const int i = i0 + 32 * j + lane;
            raw[j] = i < hi ? g[i] : 0;
        }
#pragma unroll
        for (int j = 0; j < EX_ROWS; j++) {
            const int i = i0 + 32 * j + lane;
            nn += raw[j] == 'N';
            if (meth && raw[j] == 'C' && i + 1 < c.len && g[i + 1] == 'G') nc++;
        }
    }
    nn = __reduce_add_sync(0xFFFFFFFFu, nn);
    nc = __reduce_add_sync(0xFFFFFFFFu, nc);
    if (lane == 0) q.cnt[s] = ((uint64_t)nn << 32) | nc;
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
    const int s = blockIdx.x * (EX_THREADS / 32) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (s >= q.n_pieces) return;
    const PieceRef pr = q.pieces[s];
    const Coord c = q.coords[pr.read];
    const int len = c.len;
    const int lo = pr.k * XSEG, hi = min(len, lo + XSEG);
    const int ld_hi = min(len, hi + 1);  // the base after the piece decides whether its last position is a CpG site
    const uint8_t *g = q.genome + q.contig_off[c.contig] + c.pos;
    const uint8_t *mt = q.meth ? q.meth + q.contig_off[c.contig] + c.pos : nullptr;
    uint8_t *out = q.out + q.out_off[pr.read];
    const bool neg = c.strand == '-';
    const bool meth = read_has_meth(q, c);
    const uint32_t lt = (1u << lane) - 1;
    // stream positions at the start of the piece; the states themselves are computed at the first hit
    const uint64_t n_before = (q.cnt_off[s] >> 32) - (q.cnt_off[s - pr.k] >> 32);
    const uint64_t m_before = q.meth_draw_base + (q.cnt_off[s] & 0xFFFFFFFFu);
    uint32_t xn = 0, xm = 0;
    bool have_xn = false, have_xm = false;
    bool carry = false;  // '-' reads: the site at the previous position (previous row / previous piece) was marked
    if (meth && neg && lo > 0 && g[lo - 1] == 'C' && g[lo] == 'G') {
        const uint32_t x = mulmod31(q.meth_residue, powmod31(LEHMER_A, m_before));  // that site's own draw: ordinal m_before - 1
        carry = (int)(lehmer_unit(x) * 254) <= (int)mt[lo - 1];
    }
    for (int i0 = lo; i0 < hi; i0 += 32 * EX_ROWS) {
        uint32_t raw[EX_ROWS + 1];
#pragma unroll
        for (int j = 0; j < EX_ROWS; j++) {
            const int i = i0 + 32 * j + lane;
            raw[j] = i < ld_hi ? g[i] : 0;
        }
        raw[EX_ROWS] = (lane == 0 && i0 + 32 * EX_ROWS < ld_hi) ? g[i0 + 32 * EX_ROWS] : 0;
#pragma unroll
        for (int j = 0; j < EX_ROWS; j++) {
            const int i = i0 + 32 * j + lane;
            const bool valid = i < hi;
            if (__all_sync(0xFFFFFFFFu, !valid)) break;
            uint32_t nxt = __shfl_down_sync(0xFFFFFFFFu, raw[j], 1);
            const uint32_t nxt_row = __shfl_sync(0xFFFFFFFFu, raw[j + 1], 0);
            if (lane == 31) nxt = nxt_row;
            uint8_t b = (uint8_t)raw[j];
            const bool is_n = valid && raw[j] == 'N';
            const uint32_t mn = __ballot_sync(0xFFFFFFFFu, is_n);
            if (mn) {
                if (!have_xn) {
                    xn = mulmod31(100u, powmod31(LEHMER_A, n_before));
                    have_xn = true;
                }
                if (is_n) {
                    const uint32_t x = mulmod31(xn, pw[__popc(mn & lt) + 1]);
                    const int n = (int)round(lehmer_unit(x) * 3);
                    b = n == 0 ? 'A' : n == 1 ? 'C' : n == 2 ? 'G' : 'T';
                }
                xn = mulmod31(xn, pw[__popc(mn)]);
            }
            bool mark = false;
            if (meth) {
                const bool site = valid && i + 1 < len && raw[j] == 'C' && nxt == 'G';
                const uint32_t mc = __ballot_sync(0xFFFFFFFFu, site);
                if (mc) {
                    if (!have_xm) {
                        xm = mulmod31(q.meth_residue, powmod31(LEHMER_A, m_before));
                        have_xm = true;
                    }
                    if (site) {
                        const uint32_t x = mulmod31(xm, pw[__popc(mc & lt) + 1]);
                        mark = (int)(lehmer_unit(x) * 254) <= (int)mt[i];
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
}

}  // namespace sqg
