// sqg_svb.cuh — svb-zd streams and finished BLOW5 records on the GPU.
//
// Every read's int16 signal is re-coded, in HBM, as the byte stream slow5lib's ptr_compress_svb_zd produces
// (slow5lib/src/slow5_press.c:1055-1087: uint32 sample count, then StreamVByte of the zig-zag deltas - ceil(n/4) key
// bytes with 2 bits per value, then 1-4 little-endian data bytes per value, thirdparty/streamvbyte/src/
// streamvbyte_encode.c:31-79, streamvbyte_zigzag.c:15-27).  With SQG_WANT_RECORDS the stream sits inside the exact bytes
// slow5_rec_to_mem() (slow5lib/src/slow5.c:3815-4010) makes of the read for a BLOW5 file with record compression NONE and
// signal compression SVB_ZD - record size, read id, primary fields, stream, auxiliary fields - all records back to back,
// so that the host's part of writing them is one fwrite.
//
// ONE pass over the signal, every sample coded ONCE.  The reads are cut into SEGMENTS of 8192 samples; a CTA takes
// segments in ticket order (atomic counter, so that a segment's predecessors are always running or done), codes the segment
// into shared memory (each warp its own 1024 samples: keys and data bytes, a running offset from a warp scan per 256
// samples), learns where its data bytes go from a decoupled look-back over the segments before it (status | value words,
// as in single-pass prefix scans; the whole CTA inspects 256 predecessors at a time), and copies keys and data out with
// 16-byte stores (the staging area is read through a byte-funnel, so that the stores are aligned whatever the position).
// What does not depend on the data - record header, keys, auxiliary fields - has a position known up front (L1), so:
//   byte position of a segment's data = fixed[read] + (data bytes of ALL segments before it).
//
//   L1 svb_layout_kernel  one CTA : per read - segments, fixed bytes before its data; scans              -> seg0, fixed
//      svb_segmap_kernel  warp per read: which read a segment belongs to                                 -> seg_read
//   L2 svb_encode_kernel  persistent-by-ticket: keys + data of every segment; per segment the data bytes before it
//                                                                                                         -> out, seg_excl
//   L3 svb_finish_kernel  warp per read: stream header / record header and auxiliary fields, offsets     -> out, off, len
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace sqg {

struct SvbParams {
    const int16_t *sig;          // signal arena
    const int64_t *read_sigoff;  // start of read r in the arena (multiple of 64 samples)
    const uint32_t *read_siglen;
    int64_t *svb_len;            // per read: bytes of its stream (or record)
    int64_t *svb_off;            // n_reads + 1: start of read r's stream (or record) in out; [n_reads] = total bytes
    uint8_t *out;
    int32_t n_reads;
    // layout (L1)
    int64_t *seg0;               // n_reads + 1: first segment of read r; [n_reads] = number of segments
    int32_t *seg_read;           // per segment a 32-byte record: {read, samples of the read, segment within the read, -, arena offset of the read (8 bytes), -, -} (svb_segmap_kernel)
    int64_t *fixed;              // n_reads + 1: bytes that do not depend on the data, up to the start of read r's data area
    unsigned long long *seg_state;   // per segment: look-back word (zeroed before L2)
    unsigned long long *seg_excl;    // per segment (+1): data bytes of all segments before it; [segments] = all data bytes
    unsigned long long *read_d0;     // per read: data bytes of all reads before it (all-ones until its first segment knows)
    unsigned int *ticket;            // zeroed before L2
    int64_t *totals;                 // [0] total bytes
    // records (SQG_WANT_RECORDS; all NULL / 0 for plain streams)
    int32_t records;
    const char *ids;             // read ids back to back
    const int64_t *id_off;       // n_reads + 1
    const double *read_offset, *read_median;
    double digitisation, range, sample_rate;
    int64_t read_number0;        // aux read_number of read 0
    uint64_t start_time0;        // aux start_time of read 0 (samples before the batch)
    int32_t ont_friendly;        // an end_reason byte follows (src/gensig.c:158-166, :211-217)
};

#ifndef SQG_SVB_THREADS
#define SQG_SVB_THREADS 256
#endif
constexpr int SVB_THREADS = SQG_SVB_THREADS;
constexpr int SVB_WARPS = SVB_THREADS / 32;
constexpr int SVB_WCHUNK = 1024;                  // samples of a segment coded by one warp: 4 steps of 256 (8 per lane)
constexpr int SVB_WSTEPS = SVB_WCHUNK / 256;
constexpr int SVB_SEG = SVB_WCHUNK * SVB_WARPS;   // samples per segment (one look-back per segment): 8192
constexpr int SVB_WDATA = SVB_WCHUNK * 3;         // data bytes of a warp's chunk at most
constexpr uint32_t SVB_REC_HEAD = 8 + 2 + 4 + 8 * 4 + 8;   // record size, id length, read_group, 4 doubles, signal bytes (+ the id itself)
// auxiliary fields as set_record_aux_fields writes them (src/gensig.c:185-217): channel_number (uint64 length + "0",
// slow5lib/src/slow5.c:4005), median_before f64, read_number i32, start_mux u8, start_time u64 (+ end_reason u8 when
// ont-friendly)
__host__ __device__ inline uint32_t svb_rec_tail(int ont) { return 8 + 1 + 8 + 4 + 1 + 8 + (ont ? 1 : 0); }
constexpr uint32_t SVB_TAIL_START_TIME = 8 + 1 + 8 + 4 + 1;   // offset of start_time within the tail

// the halves of a word as sign-extended int16 (prmt with the sign-replication bit of the selector; the __byte_perm
// intrinsic does not pass that bit on)
__device__ __forceinline__ int32_t svb_lo16(uint32_t w) {
    int32_t r;
    asm("prmt.b32 %0, %1, 0, 0x9910;" : "=r"(r) : "r"(w));
    return r;
}
__device__ __forceinline__ int32_t svb_hi16(uint32_t w) {
    int32_t r;
    asm("prmt.b32 %0, %1, 0, 0xBB32;" : "=r"(r) : "r"(w));
    return r;
}

// zig-zag deltas of the lane's 8 consecutive samples (q: the 8 int16, prev: the sample before them) and their byte counts
// (|delta| < 2^16 -> zig-zag < 2^17: 1 + [v >= 2^8] + [v >= 2^16] bytes, in add/shift arithmetic)
__device__ __forceinline__ void svb_code8(const uint4 &q, int32_t prev, uint32_t (&v)[8], uint32_t (&nb)[8]) {
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const int32_t s = (j & 1) ? svb_hi16(w[j >> 1]) : svb_lo16(w[j >> 1]);
        const int32_t d = s - prev;
        prev = s;
        v[j] = ((uint32_t)d + (uint32_t)d) ^ (uint32_t)(d >> 31);
        nb[j] = 1u + ((v[j] + 0xFFFF00u) >> 24) + (v[j] >> 16);
    }
}
__device__ __forceinline__ int32_t svb_last(const uint4 &q) { return svb_hi16(q.w); }

// L1: per read the number of segments and the fixed bytes in front of its data area; exclusive scans of both.
// Streams: fixed = 4 (sample count) + keys.  Records: + record head, id, and the tails of the records before.
__global__ void __launch_bounds__(1024) svb_layout_kernel(const SvbParams p) {
    __shared__ uint64_t s_a[32], s_b[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint64_t tail = p.records ? svb_rec_tail(p.ont_friendly) : 0;
    uint64_t ca = 0, cb = 0;   // carries: segments, fixed bytes (incl. the tails of finished records)
    for (int r0 = 0; r0 < p.n_reads; r0 += 1024) {
        const int r = r0 + tid;
        uint64_t nseg = 0, head = 0;
        if (r < p.n_reads) {
            const uint64_t n = p.read_siglen[r];
            nseg = (n + SVB_SEG - 1) / SVB_SEG;
            head = 4 + (n + 3) / 4;
            if (p.records) head += SVB_REC_HEAD + (uint64_t)(p.id_off[r + 1] - p.id_off[r]);
        }
        uint64_t ia = nseg, ib = head + (r < p.n_reads ? tail : 0);
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint64_t va = __shfl_up_sync(0xffffffffu, ia, o), vb = __shfl_up_sync(0xffffffffu, ib, o);
            if (lane >= o) { ia += va; ib += vb; }
        }
        __syncthreads();
        if (lane == 31) { s_a[warp] = ia; s_b[warp] = ib; }
        __syncthreads();
        uint64_t ba = 0, bb = 0, ta = 0, tb = 0;
#pragma unroll 8
        for (int w = 0; w < 32; w++) {
            if (w < warp) { ba += s_a[w]; bb += s_b[w]; }
            ta += s_a[w]; tb += s_b[w];
        }
        if (r < p.n_reads) {
            p.seg0[r] = (int64_t)(ca + ba + ia - nseg);
            // fixed[r]: everything fixed before read r's DATA: heads and tails of the reads before + this read's head
            p.fixed[r] = (int64_t)(cb + bb + ib - tail);
        }
        ca += ta; cb += tb;
    }
    if (tid == 0) {
        p.seg0[p.n_reads] = (int64_t)ca;
        p.fixed[p.n_reads] = (int64_t)cb;   // all fixed bytes of the batch
    }
}

// what a CTA needs to know about a segment, in one record: one warp per read, lanes over its segments
__global__ void __launch_bounds__(256) svb_segmap_kernel(const SvbParams p) {
    const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= p.n_reads) return;
    const int64_t s0 = p.seg0[r], s1 = p.seg0[r + 1];
    const uint32_t n = p.read_siglen[r];
    const unsigned long long x = (unsigned long long)p.read_sigoff[r];
    uint4 *rec = reinterpret_cast<uint4 *>(p.seg_read);
    for (int64_t sg = s0 + lane; sg < s1; sg += 32) {
        rec[2 * sg] = make_uint4((uint32_t)r, n, (uint32_t)(sg - s0), 0u);
        rec[2 * sg + 1] = make_uint4((uint32_t)x, (uint32_t)(x >> 32), 0u, 0u);
    }
}

constexpr unsigned long long SVB_FLAG_AGG = 1ull << 62, SVB_FLAG_PRE = 2ull << 62, SVB_VAL_MASK = (1ull << 62) - 1;

// bytes hb .. hb+15 of the 32 bytes (a : b), hb in 0..15 (warp-uniform)
__device__ __forceinline__ uint4 svb_shift16(const uint4 &a, const uint4 &b, uint32_t hb) {
    const uint32_t q = hb >> 2, sel = 0x3210u + 0x1111u * (hb & 3u);
    uint32_t w0, w1, w2, w3, w4;
    if (q == 0) { w0 = a.x; w1 = a.y; w2 = a.z; w3 = a.w; w4 = b.x; }
    else if (q == 1) { w0 = a.y; w1 = a.z; w2 = a.w; w3 = b.x; w4 = b.y; }
    else if (q == 2) { w0 = a.z; w1 = a.w; w2 = b.x; w3 = b.y; w4 = b.z; }
    else { w0 = a.w; w1 = b.x; w2 = b.y; w3 = b.z; w4 = b.w; }
    return make_uint4(__byte_perm(w0, w1, sel), __byte_perm(w1, w2, sel), __byte_perm(w2, w3, sel), __byte_perm(w3, w4, sel));
}

// a warp's staged bytes (src: 16-byte aligned shared memory, readable 16 bytes past n) -> global, at any alignment of dst:
// 16-byte stores for the aligned body of dst - the source is read through a byte funnel -, head and tail bytes one by one
__device__ __forceinline__ void svb_warp_copy(uint8_t *dst, const uint8_t *src, uint32_t n, int lane) {
    const uint32_t head = min(n, (uint32_t)((16 - (reinterpret_cast<uintptr_t>(dst) & 15)) & 15));
    if ((uint32_t)lane < head) dst[lane] = src[lane];
    const uint32_t body = (n - head) >> 4;
    const uint4 *s4 = reinterpret_cast<const uint4 *>(src);
    uint4 *d4 = reinterpret_cast<uint4 *>(dst + head);
    for (uint32_t i = lane; i < body; i += 32) d4[i] = head ? svb_shift16(s4[i], s4[i + 1], head) : s4[i];
    const uint32_t done = head + (body << 4);
    if (done + lane < n) dst[done + lane] = src[done + lane];
}

// L2: one segment per CTA iteration, in ticket order
__global__ void __launch_bounds__(SVB_THREADS) svb_encode_kernel(const SvbParams p) {
    __shared__ __align__(16) uint8_t s_data[SVB_WARPS][SVB_WDATA + 32];
    __shared__ __align__(16) uint8_t s_keys[SVB_WARPS][SVB_WCHUNK / 4 + 32];
    __shared__ uint32_t s_warp[SVB_WARPS];
    __shared__ unsigned long long s_excl;
    __shared__ int s_seg;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t nseg_all = p.seg0[p.n_reads];
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_seg = (int)atomicAdd(p.ticket, 1u);   // (not taken ahead of time: a ticket held back stalls every look-back behind it)
        __syncthreads();
        const int seg = s_seg;
        if ((int64_t)seg >= nseg_all) return;
        // the segment's read, its length, the segment's place in it and the read's place in the arena: one record, one round trip
        const uint4 m0 = reinterpret_cast<const uint4 *>(p.seg_read)[2 * (size_t)seg], m1 = reinterpret_cast<const uint4 *>(p.seg_read)[2 * (size_t)seg + 1];
        const int r = (int)m0.x;
        const uint32_t n = m0.y;
        const uint32_t sl = m0.z;                                      // segment within the read
        const int16_t *x = p.sig + (long long)(((unsigned long long)m1.y << 32) | m1.x);
        const uint32_t w_i0 = sl * SVB_SEG + warp * SVB_WCHUNK;      // first sample of this warp's chunk
        uint8_t *stage = s_data[warp];
        // ---- the warp codes its chunk into shared memory: 256 samples per step, all four loads issued first ----
        uint32_t woff = 0;   // data bytes staged so far (warp-uniform)
        if (w_i0 < n) {
            int32_t prev = w_i0 ? (int32_t)x[w_i0 - 1] : 0;
            uint4 qq[SVB_WSTEPS];
#pragma unroll
            for (int st = 0; st < SVB_WSTEPS; st++) {
                const uint32_t i0 = w_i0 + st * 256 + lane * 8;
                qq[st] = make_uint4(0, 0, 0, 0);
                if (i0 < n) qq[st] = *reinterpret_cast<const uint4 *>(x + i0);
            }
#pragma unroll
            for (int st = 0; st < SVB_WSTEPS; st++) {
                const uint32_t step_i0 = w_i0 + st * 256;
                if (step_i0 >= n) break;
                const uint32_t i0 = step_i0 + lane * 8;
                const uint4 q = qq[st];
                int32_t pl = __shfl_up_sync(0xffffffffu, svb_last(q), 1);
                if (lane == 0) pl = prev;
                prev = __shfl_sync(0xffffffffu, svb_last(q), 31);
                uint32_t v[8], nb[8];
                svb_code8(q, pl, v, nb);
                if (i0 + 8 > n) {       // the read ends inside (or before) this lane's samples
#pragma unroll
                    for (int j = 0; j < 8; j++)
                        if (i0 + j >= n) { nb[j] = 0; v[j] = 0; }
                }
                const uint32_t mine = ((nb[0] + nb[1]) + (nb[2] + nb[3])) + ((nb[4] + nb[5]) + (nb[6] + nb[7]));
                uint32_t inc = mine;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
                    if (lane >= o) inc += t;
                }
                const uint32_t step_total = __shfl_sync(0xffffffffu, inc, 31);
                // two key bytes: four 2-bit codes (bytes - 1) each, first value in the low bits; a value behind the end codes 0
                const uint32_t c0 = nb[0] ? nb[0] - 1 : 0u, c1 = nb[1] ? nb[1] - 1 : 0u, c2 = nb[2] ? nb[2] - 1 : 0u, c3 = nb[3] ? nb[3] - 1 : 0u;
                const uint32_t c4 = nb[4] ? nb[4] - 1 : 0u, c5 = nb[5] ? nb[5] - 1 : 0u, c6 = nb[6] ? nb[6] - 1 : 0u, c7 = nb[7] ? nb[7] - 1 : 0u;
                const uint32_t k0 = c0 + 4 * c1 + 16 * c2 + 64 * c3, k1 = c4 + 4 * c5 + 16 * c6 + 64 * c7;
                *reinterpret_cast<uint16_t *>(&s_keys[warp][st * 64 + 2 * lane]) = (uint16_t)(k0 | (k1 << 8));
                // data: every value but the lane's last stores two bytes unconditionally (the second is overwritten by the
                // next value when it has one byte only); third bytes (|delta| >= 2^15: none in any sane signal) by a
                // branch of their own
                uint8_t *d = stage + woff + (inc - mine);
                if (i0 + 8 <= n) {
#pragma unroll
                    for (int j = 0; j < 7; j++) {
                        d[0] = (uint8_t)v[j];
                        d[1] = (uint8_t)(v[j] >> 8);
                        d += nb[j];
                    }
                    d[0] = (uint8_t)v[7];
                    if (nb[7] > 1) d[1] = (uint8_t)(v[7] >> 8);
                    if (((v[0] | v[1] | v[2] | v[3]) | (v[4] | v[5] | v[6] | v[7])) >> 16) {   // some value has three bytes
                        d = stage + woff + (inc - mine);
#pragma unroll
                        for (int j = 0; j < 8; j++) {
                            if (nb[j] > 2) d[2] = (uint8_t)(v[j] >> 16);
                            d += nb[j];
                        }
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        if (nb[j] > 0) d[0] = (uint8_t)v[j];
                        if (nb[j] > 1) d[1] = (uint8_t)(v[j] >> 8);
                        if (nb[j] > 2) d[2] = (uint8_t)(v[j] >> 16);
                        d += nb[j];
                    }
                }
                woff += step_total;
            }
        }
        if (lane == 0) s_warp[warp] = woff;
        __syncthreads();
        // ---- decoupled look-back by warp 0: data bytes of all segments before this one.  Lane l looks at segment j - l; the
        // window ends at the nearest published PREFIX (everything nearer must at least have published its AGGREGATE, else
        // the window is read again after a short sleep); the other warps wait at the barrier below.  The status words are
        // published by reductions without a return value (RED: sent and forgotten, yet performed at L2 at once - plain stores
        // measured slower to become visible) and without a fence: a word carries its value ----
        uint32_t seg_total = 0, warp_before = 0;
#pragma unroll
        for (int w = 0; w < SVB_WARPS; w++) {
            if (w < warp) warp_before += s_warp[w];
            seg_total += s_warp[w];
        }
        if (warp == 0) {
            unsigned long long excl = 0;
            volatile unsigned long long *state = reinterpret_cast<volatile unsigned long long *>(p.seg_state);
            if (seg > 0) {
                if (lane == 0) atomicAdd(p.seg_state + seg, SVB_FLAG_AGG | (unsigned long long)seg_total);   // (the word was zero; no value returned: a reduction, sent and forgotten)
                int j = seg - 1;        // window: segments j, j-1, .., j-31
                for (;;) {
                    const int mine = j - lane;
                    unsigned long long w = SVB_FLAG_PRE;       // (before segment 0: a prefix of nothing)
                    if (mine >= 0) w = state[mine];
                    const unsigned int st = (unsigned int)(w >> 62);
                    const unsigned int pre_b = __ballot_sync(0xffffffffu, st == 2), rdy_b = __ballot_sync(0xffffffffu, st != 0);
                    const unsigned int nearer = pre_b ? ((pre_b & (0u - pre_b)) - 1u) : 0xFFFFFFFFu;   // lanes in front of the nearest prefix
                    if (~rdy_b & nearer) {   // one of them has not published yet
                        __nanosleep(40);
                        continue;
                    }
                    unsigned long long part = (((nearer << 1) | 1u) >> lane) & 1u ? (w & SVB_VAL_MASK) : 0ull;   // lanes up to and incl. the prefix
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
                    excl += part;
                    if (pre_b) break;
                    j -= 32;
                }
            }
            if (lane == 0) {
                s_excl = excl;
                // AGGREGATE | total  ->  PREFIX | excl + total, by one more reduction
                atomicAdd(p.seg_state + seg, seg > 0 ? (SVB_FLAG_PRE - SVB_FLAG_AGG) + excl : SVB_FLAG_PRE | (unsigned long long)seg_total);
                p.seg_excl[seg] = excl;
                if ((int64_t)seg == nseg_all - 1) p.seg_excl[nseg_all] = excl + seg_total;
                if (sl == 0) *reinterpret_cast<volatile unsigned long long *>(p.read_d0 + r) = excl;
            }
        }
        __syncthreads();
        const unsigned long long excl = s_excl;
        if (w_i0 >= n) continue;   // (this warp's chunk lies behind the read's end)
        // keys of the segment lie at (start of the read's key area) + 2048 sl.  The key area ends where the read's data area
        // starts: fixed[r] + (data bytes of the READS before) - a value its first segment publishes (that segment holds
        // an earlier ticket: it is running or done)
        unsigned long long d0 = excl;
        if (sl != 0) {
            do { d0 = *reinterpret_cast<volatile unsigned long long *>(p.read_d0 + r); } while (d0 == ~0ull);
        }
        // ---- copy out: this warp's keys and data bytes ----
        const uint32_t nkeys_w = (min(n - w_i0, (uint32_t)SVB_WCHUNK) + 3) / 4;
        svb_warp_copy(p.out + p.fixed[r] + d0 - (uint64_t)((n + 3) / 4) + (uint64_t)(w_i0 / 4), s_keys[warp], nkeys_w, lane);
        svb_warp_copy(p.out + p.fixed[r] + excl + warp_before, stage, woff, lane);
    }
}

__device__ __forceinline__ void svb_put(uint8_t *&q, const void *src, int n) {
    const uint8_t *s = reinterpret_cast<const uint8_t *>(src);
    for (int i = 0; i < n; i++) q[i] = s[i];
    q += n;
}

// L3: what frames the keys and data of a read - warp per read (lane 0 writes the few scalar fields, the warp the id)
__global__ void __launch_bounds__(256) svb_finish_kernel(const SvbParams p) {
    const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= p.n_reads) return;
    const uint64_t n = p.read_siglen[r];
    const uint64_t nkeys = (n + 3) / 4;
    // data bytes of the reads before this one, and of this one
    const int64_t s0 = p.seg0[r], s1 = p.seg0[r + 1];
    const unsigned long long d0 = p.seg_excl[s0], d1 = p.seg_excl[s1];
    const uint64_t data_bytes = d1 - d0;
    const uint64_t stream_bytes = 4 + nkeys + data_bytes;
    const uint64_t idlen = p.records ? (uint64_t)(p.id_off[r + 1] - p.id_off[r]) : 0;
    const uint64_t head = 4 + nkeys + (p.records ? SVB_REC_HEAD + idlen : 0);
    const uint64_t tail = p.records ? svb_rec_tail(p.ont_friendly) : 0;
    uint8_t *start = p.out + p.fixed[r] + d0 - head;   // first byte of the read's stream / record
    if (p.records) {
        for (uint64_t i = lane; i < idlen; i += 32) start[8 + 2 + i] = (uint8_t)p.ids[p.id_off[r] + i];
    }
    if (lane == 0) {
        uint8_t *q = start;
        if (p.records) {
            const uint64_t rec_size = head - 8 + data_bytes + tail;       // bytes behind the size field
            const uint16_t idl = (uint16_t)idlen;
            const uint32_t group = 0;
            const double off = p.read_offset[r];
            svb_put(q, &rec_size, 8);
            svb_put(q, &idl, 2);
            q += idlen;
            svb_put(q, &group, 4);
            svb_put(q, &p.digitisation, 8);
            svb_put(q, &off, 8);
            svb_put(q, &p.range, 8);
            svb_put(q, &p.sample_rate, 8);
            svb_put(q, &stream_bytes, 8);                                   // len_raw_signal = bytes of the compressed signal
        }
        const uint32_t n32 = (uint32_t)n;
        svb_put(q, &n32, 4);                                                // slow5_press.c:1047: the original length
        if (p.records) {
            q = start + head + data_bytes;
            const uint64_t chlen = 1;
            const char ch = '0';
            const double med = p.read_median[r];
            const int32_t rn = (int32_t)(p.read_number0 + r);
            const uint8_t mux = 0;
            // start_time: samples before the read = start_time0 + (samples of the batch's reads before it)
            svb_put(q, &chlen, 8);
            svb_put(q, &ch, 1);
            svb_put(q, &med, 8);
            svb_put(q, &rn, 4);
            svb_put(q, &mux, 1);
            q += 8;   // start_time: written by svb_start_time_kernel (needs the prefix of the lengths)
            if (p.ont_friendly) { const uint8_t er = 0; svb_put(q, &er, 1); }
        }
        p.svb_off[r] = (int64_t)(start - p.out);
        p.svb_len[r] = (int64_t)(head + data_bytes + tail);
        if (r == p.n_reads - 1) {
            p.svb_off[p.n_reads] = (int64_t)(start - p.out) + (int64_t)(head + data_bytes + tail);
            p.totals[0] = p.svb_off[p.n_reads];
        }
    }
}

// records: aux start_time = exclusive prefix of the reads' lengths (src/sim.c:602), one CTA, 1024 reads per round
__global__ void __launch_bounds__(1024) svb_start_time_kernel(const SvbParams p) {
    __shared__ uint64_t s_w[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint64_t carry = p.start_time0;
    const uint32_t tail = svb_rec_tail(p.ont_friendly);
    for (int r0 = 0; r0 < p.n_reads; r0 += 1024) {
        const int r = r0 + tid;
        const uint64_t l = r < p.n_reads ? (uint64_t)p.read_siglen[r] : 0ull;
        uint64_t inc = l;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint64_t v = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += v;
        }
        __syncthreads();
        if (lane == 31) s_w[warp] = inc;
        __syncthreads();
        uint64_t before = 0, all = 0;
#pragma unroll 8
        for (int w = 0; w < 32; w++) {
            if (w < warp) before += s_w[w];
            all += s_w[w];
        }
        if (r < p.n_reads) {
            const uint64_t st = carry + before + inc - l;
            uint8_t *q = p.out + p.svb_off[r] + p.svb_len[r] - tail + SVB_TAIL_START_TIME;
            for (int i = 0; i < 8; i++) q[i] = (uint8_t)(st >> (8 * i));
        }
        carry += all;
    }
}

}  // namespace sqg
