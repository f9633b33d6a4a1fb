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
// A piece without N's of a read without methylation - nearly every piece of a real genome - needs no stream at all: it
// is a copy or a reverse complement, done 16 output bytes per lane and step (aligned 16-byte stores, the source read
// as aligned 16-byte words through a byte funnel, the complement as SWAR arithmetic on four bases per register).
#pragma once

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
    uint64_t *cnt_off;       // exclusive scan of cnt over the batch's pieces (methylation only)
    uint32_t meth_residue;   // (seed + 6) mod m
    uint64_t meth_draw_base;
    int32_t n_pieces, do_meth;
    uint32_t pw[33];         // 16807^j mod m, j = 0..32
};

__device__ __forceinline__ bool read_has_meth(const ExtractParams &q, const Coord &c) {
    return q.do_meth && q.meth && (!q.contig_has_meth || q.contig_has_meth[c.contig]);
}

// 16 bytes from any address: the two aligned 16-byte words around them (the second only touched when the address is not a
// multiple of 16: it then holds bytes of the window or of the 64 bytes of padding the genome buffers end in) through a
// byte funnel.  The misalignment is the same for every chunk of a piece, i.e. warp-uniform.  x[0] = lowest address.
template <int Q>
__device__ __forceinline__ void ex_pick(const uint4 &A, const uint4 &B, uint32_t sel, uint32_t (&x)[4]) {
    const uint32_t w[8] = {A.x, A.y, A.z, A.w, B.x, B.y, B.z, B.w};
#pragma unroll
    for (int j = 0; j < 4; j++) x[j] = __byte_perm(w[Q + j], w[Q + j + 1], sel);
}
__device__ __forceinline__ void ex_load16(const uint8_t *src, uint32_t (&x)[4]) {
    const uint32_t sh = (uint32_t)(reinterpret_cast<uintptr_t>(src) & 15u);
    const uint4 *a = reinterpret_cast<const uint4 *>(src - sh);
    const uint4 A = __ldg(a), B = sh ? __ldg(a + 1) : make_uint4(0u, 0u, 0u, 0u);
    const uint32_t sel = 0x3210u + 0x1111u * (sh & 3u);
    switch (sh >> 2) {
        case 0: ex_pick<0>(A, B, sel, x); break;
        case 1: ex_pick<1>(A, B, sel, x); break;
        case 2: ex_pick<2>(A, B, sel, x); break;
        default: ex_pick<3>(A, B, sel, x); break;
    }
}
// 0x80 in every byte of x that is zero (exact: no borrow travels between bytes)
__device__ __forceinline__ uint32_t ex_zero_bytes(uint32_t x) { return ~(((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x) & 0x80808080u; }
// four bases -> their complements in reverse order (src/seq.h:77-112: ACGTacgt -> TGCA, anything else -> T).  The digit
// ((c>>1) ^ (c>>2)) & 3 of ACGT is 0 1 2 3 in either case; the digits, as PRMT selectors, pick the complement out of
// "TGCA" and - to tell the eight letters from every other byte - the letter itself out of "ACGT".
__device__ __forceinline__ uint32_t ex_revcomp4(uint32_t w) {
    const uint32_t d = ((w >> 1) ^ (w >> 2)) & 0x03030303u;
    const uint32_t u = (d | (d >> 4)) & 0x00FF00FFu;
    const uint32_t sel = (u | (u >> 8)) & 0xFFFFu;
    const uint32_t other = __byte_perm(0x54474341u /* "ACGT" */, 0u, sel) ^ (w & 0xDFDFDFDFu);   // non-zero bytes: not one of the eight
    uint32_t comp = __byte_perm(0x41434754u /* "TGCA" */, 0u, sel);
    if (other) {
        const uint32_t m = (0x80808080u ^ ex_zero_bytes(other)) >> 7;   // 0x01 in those bytes
        comp = (comp & ~(m * 0xFFu)) | (m * 0x54u);
    }
    return __byte_perm(comp, 0u, 0x0123);
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
    if (!meth) {
        // only the N's are wanted: 16 bytes per lane and step, all of a piece's loads in flight at once
        uint32_t x[XSEG / 512][4];
#pragma unroll
        for (int t = 0; t < XSEG / 512; t++) {
            const int i = lo + 512 * t + 16 * lane;
            x[t][0] = x[t][1] = x[t][2] = x[t][3] = 0u;
            if (i < hi) ex_load16(g + i, x[t]);
        }
#pragma unroll
        for (int t = 0; t < XSEG / 512; t++) {
            const int left = hi - (lo + 512 * t + 16 * lane);   // bytes of this chunk inside the piece
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int nb = min(max(left - 4 * j, 0), 4);
                const uint32_t valid = nb >= 4 ? 0xFFFFFFFFu : ((1u << (8 * nb)) - 1u);
                nn += __popc(ex_zero_bytes(x[t][j] ^ 0x4E4E4E4Eu) & valid);
            }
        }
        nn = __reduce_add_sync(0xFFFFFFFFu, nn);
        if (lane == 0) q.cnt[s] = (uint64_t)nn << 32;
        return;
    }
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

// exclusive scan of the pieces' packed counts (N's << 32 | CpG sites; both totals stay below 2^32: a batch has fewer
// than 2^32 bases): one CTA, sixteen consecutive pieces per thread and round (eight 16-byte loads in flight per thread;
// the kernel is one CTA: nothing else hides the latency)
constexpr int EX_SCAN_ITEMS = 16;
__global__ void __launch_bounds__(1024) extract_scan_kernel(const __grid_constant__ ExtractParams q) {
    __shared__ uint64_t s_w[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint64_t carry = 0;
    for (int i0 = 0; i0 < q.n_pieces; i0 += 1024 * EX_SCAN_ITEMS) {
        const int i = i0 + EX_SCAN_ITEMS * tid;
        uint64_t v[EX_SCAN_ITEMS];
        if (i + EX_SCAN_ITEMS <= q.n_pieces) {
#pragma unroll
            for (int j = 0; j < EX_SCAN_ITEMS; j += 2) {   // (cudaMalloc'd, i a multiple of 16)
                const ulonglong2 t = *reinterpret_cast<const ulonglong2 *>(q.cnt + i + j);
                v[j] = t.x; v[j + 1] = t.y;
            }
        } else {
#pragma unroll
            for (int j = 0; j < EX_SCAN_ITEMS; j++) v[j] = i + j < q.n_pieces ? q.cnt[i + j] : 0ull;
        }
        uint64_t mine = 0;
#pragma unroll
        for (int j = 0; j < EX_SCAN_ITEMS; j++) mine += v[j];
        uint64_t inc = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint64_t t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        __syncthreads();   // (the totals of the previous round have been read)
        if (lane == 31) s_w[warp] = inc;
        __syncthreads();
        uint64_t before = 0, all = 0;
#pragma unroll 8
        for (int w = 0; w < 32; w++) {
            if (w < warp) before += s_w[w];
            all += s_w[w];
        }
        uint64_t run = carry + before + inc - mine;
        if (i + EX_SCAN_ITEMS <= q.n_pieces) {
#pragma unroll
            for (int j = 0; j < EX_SCAN_ITEMS; j += 2) {
                *reinterpret_cast<ulonglong2 *>(q.cnt_off + i + j) = make_ulonglong2(run, run + v[j]);
                run += v[j] + v[j + 1];
            }
        } else {
#pragma unroll
            for (int j = 0; j < EX_SCAN_ITEMS; j++) {
                if (i + j < q.n_pieces) q.cnt_off[i + j] = run;
                run += v[j];
            }
        }
        carry += all;
    }
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
    if (!meth && q.cnt[s] == 0) {
        // ---- no N, no methylation: out[lo..hi) = g[lo..hi), or out[len-hi..len-lo) = the reverse complement of it ----
        const int o_lo = neg ? len - hi : lo, n = hi - lo;
        uint8_t *dst = out + o_lo;
        const int head = min(n, (int)((16 - (reinterpret_cast<uintptr_t>(dst) & 15)) & 15));
        const int body = (n - head) >> 4;
        auto one = [&](int o) {   // output byte o of the read, by itself
            if (!neg) out[o] = g[o];
            else out[o] = complement_base(g[len - 1 - o]);
        };
        if (lane < head) one(o_lo + lane);
        for (int c0 = 0; c0 < body; c0 += 128) {
            uint32_t x[4][4];
#pragma unroll
            for (int t = 0; t < 4; t++) {
                const int ci = c0 + 32 * t + lane;
                const int o = o_lo + head + 16 * ci;
                if (ci < body) ex_load16(neg ? g + (len - 16 - o) : g + o, x[t]);
            }
#pragma unroll
            for (int t = 0; t < 4; t++) {
                const int ci = c0 + 32 * t + lane;
                if (ci < body) {
                    uint4 v;
                    if (!neg) v = make_uint4(x[t][0], x[t][1], x[t][2], x[t][3]);
                    else v = make_uint4(ex_revcomp4(x[t][3]), ex_revcomp4(x[t][2]), ex_revcomp4(x[t][1]), ex_revcomp4(x[t][0]));
                    *reinterpret_cast<uint4 *>(dst + head + 16 * ci) = v;
                }
            }
        }
        const int done = head + 16 * body;
        if (done + lane < n) one(o_lo + done + lane);
        return;
    }
    const uint32_t lt = (1u << lane) - 1;
    // stream positions at the start of the piece; the states themselves are computed at the first hit
    // (the N's restart at every read: without methylation the batch-wide scan is not run at all, and a piece that does
    //  hold N's adds up the counts of its read's earlier pieces itself)
    uint64_t n_before = 0, m_before = 0;
    if (q.do_meth) {
        n_before = (q.cnt_off[s] >> 32) - (q.cnt_off[s - pr.k] >> 32);
        m_before = q.meth_draw_base + (q.cnt_off[s] & 0xFFFFFFFFu);
    } else {
        uint32_t nb = 0;
        for (int j = lane; j < pr.k; j += 32) nb += (uint32_t)(q.cnt[s - pr.k + j] >> 32);
        n_before = __reduce_add_sync(0xFFFFFFFFu, nb);
    }
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
