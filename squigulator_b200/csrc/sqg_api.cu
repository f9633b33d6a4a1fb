// sqg_api.cu — C ABI (include/sqg.h) over the kernels in sqg_kernels.cuh.
//
// Host side of the drop-in: what the reference does per batch in process_db()/work_db()
// (src/sim.c:622, src/thread.c:119) around gen_sig() is done here per SLOT: a CUDA stream with its own
// device buffers and pinned staging.  sqg_submit hands batches to slot worker threads, so the H2D
// copy, the kernels and the D2H copy of consecutive batches overlap each other and the caller's
// host-side record encoding.
//
// There is deliberately NO CPU implementation in this library: without a CUDA device every entry point
// fails with SQG_ERR_NODEVICE.
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/sqg.h"
#include <cuda/functional>
#include <nvtx3/nvToolsExt.h>

#include "sqg_kernels.cuh"
#include "sqg_signal.cuh"
#include "sqg_legacy.cuh"
#include "sqg_svb.cuh"
#include "sqg_sstext.cuh"
#include "sqg_extract.cuh"

extern "C" const unsigned char sqg_ztable_blob[];  // Z32 ++ Z2 (binary32), embedded from data/ztable_v3.bin (ztable_blob.S)

namespace {

using namespace sqg;

thread_local std::string g_init_error;
// where CU()/fail() put their message on this thread: a slot worker points it at its job, so that two batches failing at
// once never write the same std::string and sqg_wait reports the message of ITS job
thread_local std::string *tl_err_sink = nullptr;

// NVTX range for the host-side stages of a batch (visible in nsys / ncu timelines)
struct NvtxRange {
    explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};

// constant sequences of --prefix (src/genread.c:37-39, :88, :113) and of the short-read rule
// (src/gensig.c:242-245), kept in front of every uploaded base buffer
constexpr int CONST_REGION = 512;
constexpr int C_HACK = 0, C_DNA_PREFIX = 32, C_RNA_SUFFIX = 128, C_RNA_STALL = 384;
const char STALL_DNA[] = "TTTTTTTTTTTTTTTTTTAATCAA";
const char ADAPTOR_DNA[] = "GGCGTCTGCTTGGGTGTTTAACCTTTTTTTTTTAATGTACTTCGTTCAGTTACGTATTGCT";
const char ADAPTOR_RNA[] = "TGATGATGAGGGATAGACGATGGTTGTTTCTGTTGGTGCTGATATTGCTTTTTTTTTTTTTATGATGCAAGATACGCAC";
const char STALL_RNA[] = "AAAAAGAAAAAACCCCCCCCCCCCCCCCCC";
constexpr int POLYA_LEN = 158;

template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t n, bool keep = false, cudaStream_t st = 0) {
        if (n <= cap) return cudaSuccess;
        size_t ncap = std::max(n, cap + cap / 2);
        T *np = nullptr;
        cudaError_t e = cudaMalloc(&np, ncap * sizeof(T));
        if (e != cudaSuccess) return e;
        if (keep && p && cap) cudaMemcpyAsync(np, p, cap * sizeof(T), cudaMemcpyDeviceToDevice, st);
        if (p) {
            cudaStreamSynchronize(st);
            cudaFree(p);
        }
        p = np;
        cap = ncap;
        return cudaSuccess;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

template <typename T>
struct PinBuf {
    T *p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t n) {
        if (n <= cap) return cudaSuccess;
        size_t ncap = std::max(n, cap + cap / 2);
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
        cudaError_t e = cudaHostAlloc(&p, ncap * sizeof(T), cudaHostAllocDefault);
        if (e != cudaSuccess) return e;
        cap = ncap;
        return cudaSuccess;
    }
    void release() {
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
};

// One in-flight batch: stream + device buffers + pinned host staging/result buffers
struct Slot {
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::vector<cudaEvent_t> kev;  // per-step event pairs around the signal kernel
    DevBuf<uint8_t> d_bases;
    DevBuf<SegDesc> d_segs;
    DevBuf<ReadDesc> d_reads;
    DevBuf<TileDesc> d_tiles;
    DevBuf<uint32_t> d_tile_sum, d_siglen, d_n0;
    DevBuf<uint4> d_kpos;  // TK uint16 per tile: prefix of the dwells within the tile (K1 -> K4)
    DevBuf<int64_t> d_sigoff, d_meta;
    DevBuf<ReadRec> d_read_rec;
    DevBuf<double> d_offset, d_median;
    DevBuf<int16_t> d_sig;
    DevBuf<int32_t> d_ss;
    DevBuf<int64_t> d_svb_len, d_svb_off;   // SQG_WANT_SVB / SQG_WANT_RECORDS (sqg_svb.cuh)
    DevBuf<uint8_t> d_svb;
    DevBuf<int64_t> d_seg0, d_fixed, d_svb_tot, d_id_off;
    DevBuf<unsigned long long> d_seg_state, d_seg_excl, d_read_d0;
    DevBuf<int32_t> d_seg_read;
    DevBuf<unsigned int> d_ticket;
    DevBuf<char> d_ids;
    PinBuf<int64_t> h_svb_tot;
    const sqg_record_info_t *rec_info = nullptr;   // SQG_WANT_RECORDS: ids and aux constants of the batch in flight
    PinBuf<int64_t> h_svb_len, h_svb_off;
    PinBuf<uint8_t> h_svb;
    int64_t svb_bytes = 0;
    DevBuf<int64_t> d_sst_len, d_sst_off, d_ss_off;   // SQG_WANT_SS_TEXT
    DevBuf<char> d_sst;
    PinBuf<int64_t> h_sst_off;
    PinBuf<char> h_sst;
    int64_t sst_bytes = 0;
    // coordinate batches (sqg_extract.cuh)
    DevBuf<Coord> d_coords;
    DevBuf<PieceRef> d_pieces;
    PinBuf<PieceRef> h_pieces;
    DevBuf<int64_t> d_out_off;
    DevBuf<uint64_t> d_cg_count, d_cg_off;
    PinBuf<int64_t> h_base_off;
    PinBuf<uint64_t> h_cg;
    PinBuf<char> h_bases;
    bool from_coords = false;
    int64_t meth_draws = 0;
    // SQG_RNG_LEGACY scratch (per k-mer of the batch)
    DevBuf<uint32_t> d_rank, d_rank_sorted, d_idx, d_idx_sorted;
    DevBuf<uint64_t> d_dsorted, d_excl, d_heads, d_segstart, d_cpos;
    DevBuf<unsigned char> d_cub;
    PinBuf<SegDesc> h_segs;
    PinBuf<ReadDesc> h_reads;
    PinBuf<int64_t> h_meta, h_sigoff, h_len64, h_ss_off;
    PinBuf<uint32_t> h_siglen;
    PinBuf<double> h_offset, h_median;
    PinBuf<int16_t> h_sig;
    PinBuf<int32_t> h_ss;
    // batch geometry
    int64_t n_reads = 0, n_segs = 0, n_tiles = 0, total_kmers = 0, total_bases = 0, first_read = 0;
    uint32_t want = 0;
    int64_t arena_need = 0, total_samples = 0;
    bool const_written = false;
    void release() {
        d_bases.release(); d_segs.release(); d_reads.release(); d_tiles.release(); d_tile_sum.release(); d_kpos.release();
        d_siglen.release(); d_n0.release(); d_sigoff.release(); d_meta.release(); d_read_rec.release();
        d_offset.release(); d_median.release(); d_sig.release(); d_ss.release();
        d_rank.release(); d_rank_sorted.release(); d_idx.release(); d_idx_sorted.release(); d_dsorted.release();
        d_excl.release(); d_heads.release(); d_segstart.release(); d_cpos.release(); d_cub.release();
        h_segs.release(); h_reads.release(); h_meta.release(); h_sigoff.release(); h_len64.release();
        h_ss_off.release(); h_siglen.release(); h_offset.release(); h_median.release(); h_sig.release();
        h_ss.release();
        d_svb_len.release(); d_svb_off.release(); d_svb.release(); h_svb_len.release(); h_svb_off.release(); h_svb.release();
        d_seg0.release(); d_fixed.release(); d_svb_tot.release(); d_id_off.release(); d_seg_state.release(); d_seg_excl.release(); d_seg_read.release();
        d_read_d0.release(); d_ticket.release(); d_ids.release(); h_svb_tot.release();
        d_sst_len.release(); d_sst_off.release(); d_ss_off.release(); d_sst.release(); h_sst_off.release(); h_sst.release();
        d_pieces.release(); h_pieces.release();
        d_coords.release(); d_out_off.release(); d_cg_count.release(); d_cg_off.release();
        h_base_off.release(); h_cg.release(); h_bases.release();
        for (auto e : kev) cudaEventDestroy(e);
        kev.clear();
        if (ev0) cudaEventDestroy(ev0);
        if (ev1) cudaEventDestroy(ev1);
        if (stream) cudaStreamDestroy(stream);
        stream = nullptr; ev0 = ev1 = nullptr;
    }
};

struct Job {
    sqg_ticket_t ticket;
    int slot;
    int64_t n_reads;
    const char *bases;
    const int64_t *base_off;
    int64_t first_read;
    uint32_t want;
    int status = 1;  // 1 = running, <=0 = done with that code
    const sqg_coord_t *coords = nullptr;  // coordinate batch (bases/base_off unused)
    int64_t meth_draw_base = 0;
    std::string err;  // message of this job's failure (sqg_wait copies it to the context)
    const sqg_record_info_t *rec = nullptr;   // SQG_WANT_RECORDS
};

}  // namespace

struct sqg_dev_batch {
    Slot slot;
    bool planned = false;
};

struct sqg_ctx {
    sqg_config_t cfg;
    int device = 0;
    int num_sms = 0;
    std::string err;
    DevBuf<float2> d_model;
    DevBuf<float2> d_model_am;    // (A', M) by rank (model_am_kernel)
    DevBuf<float4> d_pair_model;  // by (k+1)-mer: the parameters of both of its k-mers (base-4 models)
    DevBuf<unsigned char> d_z;  // Z32 ++ Z2
    // device-resident genome (sqg_genome_load)
    DevBuf<uint8_t> d_genome, d_gmeth, d_has_meth;
    DevBuf<int64_t> d_contig_off;
    std::vector<int64_t> contig_off;
    bool genome_meth = false, genome_has_flags = false;
    GenParams base;     // configuration-derived part of the kernel parameters
    bool noisy = false, rand_dwell = false, meth = false, rev = false, prefix = false;
    int k4_grid_per_sm = 1;
    // SQG_RNG_LEGACY: positions reached in the reference's streams (src/sim.c:215-258, thread 0)
    bool legacy = false;
    uint64_t leg_dwell_pos = 0, leg_read_pos = 0;
    DevBuf<uint64_t> d_cnt_kmer;
    std::atomic<int64_t> launches{0};
    Slot sync_slot;  // used by sqg_gen_batch / sqg_gen_sig
    // dispatcher
    std::vector<Slot> slots;
    std::vector<std::thread> workers;
    std::vector<std::deque<Job *>> queues;
    std::mutex mu;
    std::condition_variable cv_work, cv_done;
    std::map<sqg_ticket_t, Job *> jobs;
    std::vector<int> slot_busy;
    sqg_ticket_t next_ticket = 1;
    bool stopping = false;
};

namespace {

void set_error(sqg_ctx *ctx, const char *msg) {
    if (tl_err_sink) *tl_err_sink = msg;
    else if (ctx) ctx->err = msg;
    else g_init_error = msg;
}

#define CU(call)                                                                                     \
    do {                                                                                             \
        cudaError_t e__ = (call);                                                                    \
        if (e__ != cudaSuccess) {                                                                    \
            char b__[512];                                                                           \
            snprintf(b__, sizeof b__, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
            set_error(ctx, b__);                                                                     \
            return e__ == cudaErrorMemoryAllocation ? SQG_ERR_NOMEM : SQG_ERR_CUDA;                  \
        }                                                                                            \
    } while (0)

int fail(sqg_ctx *ctx, int code, const char *msg) {
    set_error(ctx, msg);
    return code;
}

typedef void (*k4_fn)(const GenParams);

template <int I>
struct K4Table {
    static void fill(k4_fn *t) {
        t[I] = (k4_fn)signal_kernel<(I >> 3) & 1, (I >> 2) & 1, (I >> 1) & 1, I & 1>;
        K4Table<I - 1>::fill(t);
    }
};
template <>
struct K4Table<-1> {
    static void fill(k4_fn *) {}
};

// the instantiations of the signal kernel: <NOISY, RAND_DWELL, METH, REV>
k4_fn pick_k4(bool noisy, bool rnd, bool meth, bool rev) {
    static k4_fn tab[16];
    static std::once_flag once;
    std::call_once(once, [] { K4Table<15>::fill(tab); });
    return tab[(noisy << 3) | (rnd << 2) | (meth << 1) | (int)rev];
}

// The model table the signal kernel gathers from (16 MB for 9-mers) is pinned in L2: an access-policy window on the slot's
// stream makes hits on it persisting lines while everything else the stream touches (bases, plan, the signal itself,
// written with evict-first stores) streams through the rest of the 126 MB.  The north-star's "pinned in L2 for the
// 4^9-entry R10 9-mer" (measured: +0.5 % - with evict-first stores the table stays resident anyway).  SQG_L2_PERSIST=0
// switches it off (A/B measurements).
int slot_pin_model(sqg_ctx *ctx, Slot &s) {
    if (ctx->legacy || !ctx->noisy || !ctx->d_pair_model.p) return SQG_OK;
    if (const char *e = getenv("SQG_L2_PERSIST")) if (atoi(e) == 0) return SQG_OK;
    int dev = 0, max_win = 0, max_persist = 0;
    CU(cudaGetDevice(&dev));
    CU(cudaDeviceGetAttribute(&max_win, cudaDevAttrMaxAccessPolicyWindowSize, dev));
    CU(cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, dev));
    const size_t bytes = ctx->d_pair_model.cap * sizeof(float4);
    if (max_win <= 0 || max_persist <= 0) return SQG_OK;
    // a table larger than the carve-out (base-5 pair table of a 9-mer CpG model: 156 MB) is left to the normal policy: a
    // partial window turns most of its lines into streaming misses (measured: 0.37 -> 0.20 of the roofline)
    if (bytes > (size_t)max_persist || bytes > (size_t)max_win) return SQG_OK;
    const size_t carve = std::min<size_t>((size_t)max_persist, std::max<size_t>(bytes, (size_t)1 << 20));
    CU(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, carve));
    cudaStreamAttrValue av;
    memset(&av, 0, sizeof av);
    av.accessPolicyWindow.base_ptr = (void *)ctx->d_pair_model.p;
    av.accessPolicyWindow.num_bytes = std::min<size_t>(bytes, (size_t)max_win);
    av.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)carve / (double)std::max<size_t>(av.accessPolicyWindow.num_bytes, 1));
    av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    CU(cudaStreamSetAttribute(s.stream, cudaStreamAttributeAccessPolicyWindow, &av));
    return SQG_OK;
}

int slot_init(sqg_ctx *ctx, Slot &s) {
    CU(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
    if (int rc = slot_pin_model(ctx, s)) return rc;
    CU(cudaEventCreate(&s.ev0));
    CU(cudaEventCreate(&s.ev1));
    return SQG_OK;
}

// ---- host-side batch preparation: reads -> segments -> tiles (replaces the per-read bookkeeping at the
// top of gen_sig_core / gen_sig_core_seq, src/gensig.c:240-245, :302-309, and attach_prefix) ----
int slot_extract(sqg_ctx *ctx, Slot &s, const sqg_coord_t *coords, int64_t meth_draw_base);

int slot_prepare(sqg_ctx *ctx, Slot &s, int64_t n_reads, const char *bases, const int64_t *base_off,
                 int64_t first_read, uint32_t want, const sqg_coord_t *coords = nullptr, int64_t meth_draw_base = 0) {
    if (n_reads < 0) return fail(ctx, SQG_ERR_ARG, "negative read count");
    if (n_reads > 0x3FFFFFFF) return fail(ctx, SQG_ERR_ARG, "too many reads in one batch");
    s.from_coords = coords != nullptr;
    s.meth_draws = 0;
    if (coords) {  // lengths come with the coordinates; the bases are cut out on the device (slot_extract)
        if (ctx->contig_off.empty()) return fail(ctx, SQG_ERR_STATE, "no genome loaded (sqg_genome_load)");
        if (meth_draw_base < 0) return fail(ctx, SQG_ERR_ARG, "negative meth_draw_base");
        if (ctx->meth && ctx->genome_meth && (ctx->cfg.seed + 6 < 0 || ctx->cfg.seed + 6 > 2147483647))
            return fail(ctx, SQG_ERR_ARG, "CpG marking needs 0 <= seed + 6 <= 2^31-1 (minstd stream seed + 6, src/sim.c:253; "
                                          "larger seeds leave the stream's modular form for its first draws)");
        CU(s.h_base_off.ensure((size_t)n_reads + 1));
        const int64_t nc = (int64_t)ctx->contig_off.size() - 1;
        int64_t acc = 0;
        for (int64_t r = 0; r < n_reads; r++) {
            const sqg_coord_t &c = coords[r];
            if (c.contig < 0 || c.contig >= nc || c.len < 0 || c.pos < 0 ||
                c.pos + c.len > ctx->contig_off[c.contig + 1] - ctx->contig_off[c.contig])
                return fail(ctx, SQG_ERR_ARG, "read coordinates outside the loaded genome");
            if (c.strand != '+' && c.strand != '-') return fail(ctx, SQG_ERR_ARG, "strand must be '+' or '-'");
            s.h_base_off.p[r] = acc;
            acc += c.len;
        }
        s.h_base_off.p[n_reads] = acc;
        base_off = s.h_base_off.p;
    } else if (n_reads > 0 && (!bases || !base_off)) {
        return fail(ctx, SQG_ERR_ARG, "null input");
    }
    const int k = (int)ctx->cfg.kmer_size;
    const int T = ctx->base.T;
    const bool prefix = ctx->prefix, rna = ctx->rev;
    const int64_t nseg_max = n_reads * (prefix && rna ? 2 : 1);
    CU(s.h_segs.ensure((size_t)std::max<int64_t>(nseg_max, 1)));
    CU(s.h_reads.ensure((size_t)std::max<int64_t>(n_reads, 1)));
    CU(s.h_ss_off.ensure((size_t)n_reads + 1));
    const int64_t user0 = base_off ? base_off[0] : 0;
    const int64_t total_bases = n_reads ? base_off[n_reads] - user0 : 0;
    int64_t nseg = 0, ntile = 0, nk_total = 0;
    const int dna_prefix_len = (int)(strlen(STALL_DNA) + strlen(ADAPTOR_DNA));
    const int rna_suffix_len = POLYA_LEN + (int)strlen(ADAPTOR_RNA);
    const int stall_rna_len = (int)strlen(STALL_RNA);
    for (int64_t r = 0; r < n_reads; r++) {
        const int64_t len64 = base_off[r + 1] - base_off[r];
        if (len64 < 0 || len64 > 0x7FFFFFF0 - 512)   // (- 512: prefix/suffix sequences are added to it in int32)
            return fail(ctx, SQG_ERR_ARG, "read length out of range (int32, as in the reference)");
        const int len = (int)len64;
        const int64_t uoff = CONST_REGION + (base_off[r] - user0);
        ReadDesc &rd = s.h_reads.p[r];
        rd.seg0 = (int32_t)nseg;
        rd.nseg = 1;
        rd.ss_off = nk_total;
        rd.shift_len = 0;
        rd.pad = 0;
        s.h_ss_off.p[r] = nk_total;
        SegDesc &sg = s.h_segs.p[nseg];
        int full_len;
        if (prefix && !rna) {  // stall + adaptor in front (src/genread.c:112-121)
            sg.off_a = C_DNA_PREFIX; sg.len_a = dna_prefix_len; sg.off_b = uoff;
            full_len = dna_prefix_len + len;
        } else if (prefix && rna) {  // polyA + adaptor behind (src/genread.c:99-110)
            sg.off_a = uoff; sg.len_a = len; sg.off_b = C_RNA_SUFFIX;
            full_len = len + rna_suffix_len;
        } else {
            sg.off_a = uoff; sg.len_a = len; sg.off_b = uoff;
            full_len = len;
        }
        if (full_len < k) {  // src/gensig.c:242-245: 5 k-mers of "ACGTACGTACGT"
            sg.off_a = C_HACK; sg.len_a = 16; sg.off_b = C_HACK;
            sg.nk = 5;
        } else {
            sg.nk = full_len - k + 1;
        }
        sg.read = (int32_t)r;
        sg.k0 = 0;
        sg.k0_rng = 0;
        sg.tile0 = (int32_t)ntile;
        ntile += (sg.nk + T - 1) / T;
        nk_total += sg.nk;
        nseg++;
        if (prefix && rna) {  // the stall appended by gen_prefix_rna (src/genread.c:88-89)
            const int nk0 = sg.nk;
            SegDesc &s2 = s.h_segs.p[nseg];
            s2.off_a = C_RNA_STALL; s2.len_a = stall_rna_len; s2.off_b = C_RNA_STALL;
            s2.nk = stall_rna_len < k ? 5 : stall_rna_len - k + 1;
            if (stall_rna_len < k) { s2.off_a = C_HACK; s2.len_a = 16; s2.off_b = C_HACK; }
            s2.read = (int32_t)r;
            s2.k0 = nk0;
            s2.k0_rng = (nk0 + 7) & ~7;
            s2.tile0 = (int32_t)ntile;
            ntile += (s2.nk + T - 1) / T;
            nk_total += s2.nk;
            nseg++;
            rd.nseg = 2;
            rd.shift_len = (int)strlen(ADAPTOR_RNA) * (int)ctx->cfg.profile.dwell_mean;
        }
        if (ntile > 0x7FFFFFF0) return fail(ctx, SQG_ERR_ARG, "batch too large (tile count)");
    }
    s.h_ss_off.p[n_reads] = nk_total;
    s.n_reads = n_reads; s.n_segs = nseg; s.n_tiles = ntile; s.total_kmers = nk_total;
    s.total_bases = total_bases; s.first_read = first_read; s.want = want;

    // device buffers + uploads
    const bool fresh = s.d_bases.cap < (size_t)(CONST_REGION + total_bases + 64);
    CU(s.d_bases.ensure((size_t)(CONST_REGION + total_bases + 64), false, s.stream));
    if (fresh || !s.const_written) {
        unsigned char c[CONST_REGION];
        memset(c, 0, sizeof c);
        memcpy(c + C_HACK, "ACGTACGTACGT", 12);
        memcpy(c + C_DNA_PREFIX, STALL_DNA, strlen(STALL_DNA));
        memcpy(c + C_DNA_PREFIX + strlen(STALL_DNA), ADAPTOR_DNA, strlen(ADAPTOR_DNA));
        memset(c + C_RNA_SUFFIX, 'A', POLYA_LEN);
        memcpy(c + C_RNA_SUFFIX + POLYA_LEN, ADAPTOR_RNA, strlen(ADAPTOR_RNA));
        memcpy(c + C_RNA_STALL, STALL_RNA, strlen(STALL_RNA));
        CU(cudaMemcpyAsync(s.d_bases.p, c, CONST_REGION, cudaMemcpyHostToDevice, s.stream));
        CU(cudaStreamSynchronize(s.stream));  // c is a stack buffer
        s.const_written = true;
    }
    if (coords) {
        int rc = slot_extract(ctx, s, coords, meth_draw_base);
        if (rc != SQG_OK) return rc;
    } else if (total_bases) {
        CU(cudaMemcpyAsync(s.d_bases.p + CONST_REGION, bases + user0, (size_t)total_bases, cudaMemcpyHostToDevice, s.stream));
    }
    CU(s.d_segs.ensure((size_t)std::max<int64_t>(nseg, 1), false, s.stream));
    CU(s.d_reads.ensure((size_t)std::max<int64_t>(n_reads, 1), false, s.stream));
    if (nseg) CU(cudaMemcpyAsync(s.d_segs.p, s.h_segs.p, (size_t)nseg * sizeof(SegDesc), cudaMemcpyHostToDevice, s.stream));
    if (n_reads) CU(cudaMemcpyAsync(s.d_reads.p, s.h_reads.p, (size_t)n_reads * sizeof(ReadDesc), cudaMemcpyHostToDevice, s.stream));
    const size_t nt = (size_t)std::max<int64_t>(ntile, 1), nr = (size_t)std::max<int64_t>(n_reads, 1);
    CU(s.d_tiles.ensure(nt, false, s.stream));
    CU(s.d_tile_sum.ensure(nt, false, s.stream));
    if (ctx->rand_dwell && !ctx->legacy) CU(s.d_kpos.ensure(nt * (TK / 8), false, s.stream));
    CU(s.d_siglen.ensure(nr, false, s.stream));
    CU(s.d_n0.ensure(nr, false, s.stream));
    CU(s.d_sigoff.ensure(nr, false, s.stream));
    CU(s.d_read_rec.ensure(nr, false, s.stream));
    CU(s.d_offset.ensure(nr, false, s.stream));
    CU(s.d_median.ensure(nr, false, s.stream));
    CU(s.d_meta.ensure(4, false, s.stream));
    CU(s.h_meta.ensure(4));
    if ((want & (SQG_WANT_SS | SQG_WANT_SS_TEXT)) || ctx->legacy) CU(s.d_ss.ensure((size_t)std::max<int64_t>(nk_total, 1), false, s.stream));
    return SQG_OK;
}

// coordinate batches: the reads' bases are written into d_bases by the extraction kernels (sqg_extract.cuh)
int slot_extract(sqg_ctx *ctx, Slot &s, const sqg_coord_t *coords, int64_t meth_draw_base) {
    const size_t n = (size_t)s.n_reads;
    CU(s.h_cg.ensure(2));
    s.h_cg.p[0] = s.h_cg.p[1] = 0;
    if (!n || !s.total_bases) return SQG_OK;
    if (s.total_bases > 0xFFFFFFFFll) return fail(ctx, SQG_ERR_ARG, "coordinate batch too large (>= 2^32 bases)");
    static_assert(sizeof(Coord) == sizeof(sqg_coord_t), "Coord mirrors sqg_coord_t");
    // pieces of XSEG positions, one warp each
    size_t np = 0;
    for (size_t r = 0; r < n; r++) np += ((size_t)coords[r].len + XSEG - 1) / XSEG;
    CU(s.h_pieces.ensure(np));
    np = 0;
    for (size_t r = 0; r < n; r++)
        for (int32_t k = 0; (int64_t)k * XSEG < coords[r].len; k++) s.h_pieces.p[np++] = PieceRef{(int32_t)r, k};
    if (np > 0x7FFFFFF0) return fail(ctx, SQG_ERR_ARG, "coordinate batch too large (piece count)");
    CU(s.d_coords.ensure(n, false, s.stream));
    CU(s.d_out_off.ensure(n + 1, false, s.stream));
    CU(s.d_pieces.ensure(np, false, s.stream));
    CU(s.d_cg_count.ensure(np, false, s.stream));
    CU(s.d_cg_off.ensure(np, false, s.stream));
    CU(cudaMemcpyAsync(s.d_coords.p, coords, n * sizeof(Coord), cudaMemcpyHostToDevice, s.stream));
    CU(cudaMemcpyAsync(s.d_out_off.p, s.h_base_off.p, (n + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, s.stream));
    CU(cudaMemcpyAsync(s.d_pieces.p, s.h_pieces.p, np * sizeof(PieceRef), cudaMemcpyHostToDevice, s.stream));
    ExtractParams q;
    memset(&q, 0, sizeof q);
    q.genome = ctx->d_genome.p;
    q.contig_off = ctx->d_contig_off.p;
    q.meth = ctx->genome_meth ? ctx->d_gmeth.p : nullptr;
    q.contig_has_meth = ctx->genome_has_flags ? ctx->d_has_meth.p : nullptr;
    q.coords = s.d_coords.p;
    q.pieces = s.d_pieces.p;
    q.out_off = s.d_out_off.p;
    q.out = s.d_bases.p + CONST_REGION;
    q.cnt = s.d_cg_count.p;
    q.cnt_off = s.d_cg_off.p;
    q.n_pieces = (int32_t)np;
    q.do_meth = ctx->meth && ctx->genome_meth;
    q.meth_residue = seed_residue(ctx->cfg.seed + 6);  // rand_meth of thread 0, src/sim.c:253
    q.meth_draw_base = (uint64_t)meth_draw_base;
    q.pw[0] = 1;
    for (int j = 1; j <= 32; j++) q.pw[j] = mulmod31(q.pw[j - 1], LEHMER_A);
    const int grid = (int)((np + EX_THREADS / 32 - 1) / (EX_THREADS / 32));
    extract_count_kernel<<<grid, EX_THREADS, 0, s.stream>>>(q);
    if (q.do_meth) {   // CpG ordinals run through the batch (rand_meth is one stream); N ordinals restart at every read
        extract_scan_kernel<<<1, 1024, 0, s.stream>>>(q);
        publish_kernel<<<1, 32, 0, s.stream>>>(reinterpret_cast<const int64_t *>(q.cnt_off + (np - 1)), 1,
                                               reinterpret_cast<const int64_t *>(q.cnt + (np - 1)), 1,
                                               reinterpret_cast<int64_t *>(s.h_cg.p));
        ctx->launches += 2;
    }
    extract_reads_kernel<<<grid, EX_THREADS, 0, s.stream>>>(q);
    ctx->launches += 2;
    CU(cudaGetLastError());
    return SQG_OK;
}

GenParams slot_params(sqg_ctx *ctx, Slot &s) {
    GenParams p = ctx->base;
    p.bases = s.d_bases.p; p.segs = s.d_segs.p; p.reads = s.d_reads.p;
    p.tiles = s.d_tiles.p; p.tile_sum = s.d_tile_sum.p; p.kpos = s.d_kpos.p;
    p.read_siglen = s.d_siglen.p; p.read_n0 = s.d_n0.p; p.read_sigoff = s.d_sigoff.p; p.read_rec = s.d_read_rec.p;
    p.read_offset = s.d_offset.p; p.read_median = s.d_median.p; p.meta = s.d_meta.p;
    p.sig = s.d_sig.p; p.ss = s.d_ss.p;
    p.n_reads = (int32_t)s.n_reads; p.n_segs = (int32_t)s.n_segs; p.n_tiles = (int32_t)s.n_tiles;
    p.first_read = s.first_read;
    p.want_ss = (s.want & (SQG_WANT_SS | SQG_WANT_SS_TEXT)) ? 1 : 0;
    return p;
}

LegacyParams legacy_params(sqg_ctx *ctx, Slot &s) {
    LegacyParams q;
    memset(&q, 0, sizeof q);
    q.seed = ctx->cfg.seed;
    q.dwell_pos0 = ctx->leg_dwell_pos;
    q.read_pos0 = ctx->leg_read_pos;
    q.cnt_kmer = ctx->d_cnt_kmer.p;
    q.kmer_rank = s.d_rank.p;
    q.kmer_cpos = s.d_cpos.p;
    q.noisy = ctx->noisy; q.rand_dwell = ctx->rand_dwell; q.meth = ctx->meth; q.rev = ctx->rev;
    q.dwell_mean = ctx->cfg.profile.dwell_mean; q.dwell_std = ctx->cfg.profile.dwell_std;
    return q;
}

// K0-K3: tile descriptors, lengths and offsets.  Asynchronous on the slot's stream.
int slot_plan(sqg_ctx *ctx, Slot &s) {
    if (s.n_reads == 0) return SQG_OK;
    const GenParams p = slot_params(ctx, s);
    CU(cudaMemsetAsync(s.d_meta.p, 0, 4 * sizeof(int64_t), s.stream));
    tile_desc_kernel<<<(int)((s.n_segs + 7) / 8), 256, 0, s.stream>>>(p);
    ctx->launches++;
    const int g2 = (int)((s.n_reads + 7) / 8);  // one warp per read
    const int g1 = (int)((s.n_tiles + 3) / 4);  // legacy dwell kernel: 4 tiles (warps) per CTA
    if (ctx->legacy) {
        const LegacyParams q = legacy_params(ctx, s);
        legacy_dwell_kernel<<<g1, 128, 0, s.stream>>>(p, q);
        read_plan_kernel<true><<<g2, 256, 0, s.stream>>>(p);
        legacy_read_draws_kernel<<<g2, 256, 0, s.stream>>>(p, q);
        ctx->launches += 3;
    } else {
        if (ctx->rand_dwell) {
            dwell_kernel<<<std::min<int>(ctx->num_sms, (int)((s.n_tiles + K1_THREADS / 32 - 1) / (K1_THREADS / 32))), K1_THREADS, K1_SMEM, s.stream>>>(p);
            ctx->launches++;
            read_plan_kernel<true><<<g2, 256, 0, s.stream>>>(p);
        } else {
            read_plan_kernel<false><<<g2, 256, 0, s.stream>>>(p);
            if (p.want_ss && s.total_kmers > 0) {  // aln->ss of the fixed-dwell modes (src/gensig.c:273-281)
                fixed_ss_kernel<<<(int)((s.total_kmers + 255) / 256), 256, 0, s.stream>>>(p, s.total_kmers);
                ctx->launches++;
            }
        }
        ctx->launches++;
    }
    read_offsets_kernel<<<1, 1024, 0, s.stream>>>(p);
    ctx->launches++;
    CU(cudaGetLastError());
    return SQG_OK;
}

// read the totals back (one small D2H + sync) and make sure the signal arena is large enough
int slot_size_arena(sqg_ctx *ctx, Slot &s) {
    if (s.n_reads == 0) { s.arena_need = s.total_samples = 0; return SQG_OK; }
    publish_kernel<<<1, 32, 0, s.stream>>>(s.d_meta.p, 4, nullptr, 0, s.h_meta.p);  // not a D2H copy: see publish_kernel
    ctx->launches += 1;
    CU(cudaStreamSynchronize(s.stream));
    if (s.h_meta.p[2]) return fail(ctx, SQG_ERR_RANGE, "a read has >= UINT32_MAX samples (reference: src/sim.c:559-562)");
    s.arena_need = s.h_meta.p[0];
    s.total_samples = s.h_meta.p[1];
    CU(s.d_sig.ensure((size_t)std::max<int64_t>(s.arena_need, 64), false, s.stream));
    return SQG_OK;
}

// SQG_RNG_LEGACY signal generation: stream positions by sort + segmented scan, then the legacy sample kernel
int slot_generate_legacy(sqg_ctx *ctx, Slot &s) {
    const int64_t n = s.total_kmers;
    const size_t nn = (size_t)std::max<int64_t>(n, 1);
    CU(s.d_rank.ensure(nn, false, s.stream)); CU(s.d_rank_sorted.ensure(nn, false, s.stream));
    CU(s.d_idx.ensure(nn, false, s.stream)); CU(s.d_idx_sorted.ensure(nn, false, s.stream));
    CU(s.d_dsorted.ensure(nn, false, s.stream)); CU(s.d_excl.ensure(nn, false, s.stream));
    CU(s.d_heads.ensure(nn, false, s.stream)); CU(s.d_segstart.ensure(nn, false, s.stream));
    CU(s.d_cpos.ensure(nn, false, s.stream));
    const GenParams p = slot_params(ctx, s);
    const LegacyParams q = legacy_params(ctx, s);
    const int gb = (int)((n + 255) / 256);
    legacy_rank_kernel<<<(int)s.n_tiles, 256, 0, s.stream>>>(p, q);
    ctx->launches++;
    if (ctx->noisy && n > 0) {
        if (n > 0xFFFFFFF0ll) return fail(ctx, SQG_ERR_ARG, "legacy mode: batch too large");
        legacy_iota_kernel<<<gb, 256, 0, s.stream>>>(s.d_idx.p, n);
        int bits = 1;
        while ((1ull << bits) < ctx->cfg.num_kmer) bits++;
        size_t t1 = 0, t2 = 0, t3 = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, t1, s.d_rank.p, s.d_rank_sorted.p, s.d_idx.p, s.d_idx_sorted.p, (int)n, 0, bits, s.stream);
        cub::DeviceScan::ExclusiveSum(nullptr, t2, s.d_dsorted.p, s.d_excl.p, (int)n, s.stream);
        cub::DeviceScan::InclusiveScan(nullptr, t3, s.d_heads.p, s.d_segstart.p, cuda::maximum<uint64_t>{}, (int)n, s.stream);
        size_t tb = std::max(t1, std::max(t2, t3));
        CU(s.d_cub.ensure(tb + 16, false, s.stream));
        CU(cub::DeviceRadixSort::SortPairs(s.d_cub.p, tb, s.d_rank.p, s.d_rank_sorted.p, s.d_idx.p, s.d_idx_sorted.p, (int)n, 0, bits, s.stream));
        legacy_gather_dwell_kernel<<<gb, 256, 0, s.stream>>>(s.d_ss.p, s.d_idx_sorted.p, s.d_dsorted.p, n);
        CU(cub::DeviceScan::ExclusiveSum(s.d_cub.p, tb, s.d_dsorted.p, s.d_excl.p, (int)n, s.stream));
        legacy_heads_kernel<<<gb, 256, 0, s.stream>>>(s.d_rank_sorted.p, s.d_excl.p, s.d_heads.p, n);
        CU(cub::DeviceScan::InclusiveScan(s.d_cub.p, tb, s.d_heads.p, s.d_segstart.p, cuda::maximum<uint64_t>{}, (int)n, s.stream));
        legacy_cpos_kernel<<<gb, 256, 0, s.stream>>>(s.d_rank_sorted.p, s.d_idx_sorted.p, s.d_excl.p, s.d_segstart.p, ctx->d_cnt_kmer.p, s.d_cpos.p, n);
        legacy_carry_kernel<<<gb, 256, 0, s.stream>>>(s.d_rank_sorted.p, s.d_excl.p, s.d_segstart.p, s.d_dsorted.p, ctx->d_cnt_kmer.p, n);
        ctx->launches += 9;
    }
    const int g1 = (int)((s.n_tiles + 3) / 4);
    legacy_signal_kernel<<<g1, 128, 0, s.stream>>>(p, q);
    ctx->launches++;
    if (ctx->prefix && ctx->rev) {
        prefix_shift_kernel<<<(int)s.n_reads, 256, 0, s.stream>>>(p);
        ctx->launches++;
    }
    CU(cudaGetLastError());
    // the streams have moved on (reference: state persists across reads, src/sim.c:215-258)
    if (ctx->rand_dwell) ctx->leg_dwell_pos += (uint64_t)n;
    if (!(ctx->cfg.flags & SQG_IDEAL)) ctx->leg_read_pos += (uint64_t)s.n_reads;
    return SQG_OK;
}

// K4 (+ the RNA prefix shift).  Asynchronous.
int slot_generate(sqg_ctx *ctx, Slot &s, cudaEvent_t before = nullptr, cudaEvent_t after = nullptr) {
    if (s.n_reads == 0) return SQG_OK;
    if (ctx->legacy) return slot_generate_legacy(ctx, s);
    const GenParams p = slot_params(ctx, s);
    const int grid = (int)std::min<int64_t>((s.n_tiles + K4_WARPS - 1) / K4_WARPS, (int64_t)ctx->num_sms * ctx->k4_grid_per_sm);
    k4_fn fn = pick_k4(ctx->noisy, ctx->rand_dwell, ctx->meth, ctx->rev);
    if (before) CU(cudaEventRecord(before, s.stream));
    void *args[] = {(void *)&p};
    CU(cudaLaunchKernel((const void *)fn, dim3(grid), dim3(K4_THREADS), args, SM_TOTAL, s.stream));
    if (after) CU(cudaEventRecord(after, s.stream));
    ctx->launches++;
    if (ctx->prefix && ctx->rev) {
        prefix_shift_kernel<<<(int)s.n_reads, 256, 0, s.stream>>>(p);
        ctx->launches++;
    }
    CU(cudaGetLastError());
    return SQG_OK;
}

// SQG_WANT_SVB / SQG_WANT_RECORDS: svb-zd streams or finished BLOW5 records of the batch's reads, in HBM (sqg_svb.cuh).
// One pass over the signal; the output buffer is sized from an upper bound (3 data bytes per sample), so nothing has to
// come back to the host before the encoder runs - the total is published behind it.
int slot_compress(sqg_ctx *ctx, Slot &s) {
    s.svb_bytes = 0;
    if (s.n_reads == 0) return SQG_OK;
    const size_t n = (size_t)s.n_reads;
    const bool records = (s.want & SQG_WANT_RECORDS) != 0;
    const sqg_record_info_t *ri = s.rec_info;
    if (records && (!ri || !ri->read_ids || !ri->id_off)) return fail(ctx, SQG_ERR_ARG, "SQG_WANT_RECORDS needs the reads' ids (sqg_record_info_t)");
    const size_t nseg_max = (size_t)(s.total_samples / SVB_SEG) + n + 1;
    int64_t id_bytes = 0;
    if (records) {
        id_bytes = ri->id_off[n] - ri->id_off[0];
        for (size_t r = 0; r < n; r++) {
            const int64_t l = ri->id_off[r + 1] - ri->id_off[r];
            if (l < 0 || l > 65535) return fail(ctx, SQG_ERR_ARG, "read id length out of range (uint16, slow5lib)");
        }
    }
    const size_t cap = (size_t)s.total_samples * 3 + (size_t)s.total_samples / 4 + n * (8 + (records ? SVB_REC_HEAD + svb_rec_tail(1) : 0)) +
                       (size_t)id_bytes + 64;
    CU(s.d_svb_len.ensure(n, false, s.stream));
    CU(s.d_svb_off.ensure(n + 1, false, s.stream));
    CU(s.d_seg0.ensure(n + 1, false, s.stream));
    CU(s.d_fixed.ensure(n + 1, false, s.stream));
    CU(s.d_read_d0.ensure(n, false, s.stream));
    CU(s.d_seg_state.ensure(nseg_max, false, s.stream));
    CU(s.d_seg_read.ensure(nseg_max * 8, false, s.stream));   // (32-byte records)
    CU(s.d_seg_excl.ensure(nseg_max + 1, false, s.stream));
    CU(s.d_ticket.ensure(4, false, s.stream));
    CU(s.d_svb_tot.ensure(4, false, s.stream));
    CU(s.h_svb_tot.ensure(4));
    CU(s.d_svb.ensure(cap, false, s.stream));
    SvbParams q;
    memset(&q, 0, sizeof q);
    q.sig = s.d_sig.p; q.read_sigoff = s.d_sigoff.p; q.read_siglen = s.d_siglen.p;
    q.svb_len = s.d_svb_len.p; q.svb_off = s.d_svb_off.p; q.out = s.d_svb.p; q.n_reads = (int32_t)s.n_reads;
    q.seg0 = s.d_seg0.p; q.fixed = s.d_fixed.p; q.seg_state = s.d_seg_state.p; q.seg_excl = s.d_seg_excl.p;
    q.read_d0 = s.d_read_d0.p; q.ticket = s.d_ticket.p; q.totals = s.d_svb_tot.p; q.seg_read = s.d_seg_read.p;
    if (records) {
        CU(s.d_ids.ensure((size_t)std::max<int64_t>(id_bytes, 1), false, s.stream));
        CU(s.d_id_off.ensure(n + 1, false, s.stream));
        CU(cudaMemcpyAsync(s.d_ids.p, ri->read_ids + ri->id_off[0], (size_t)id_bytes, cudaMemcpyHostToDevice, s.stream));
        CU(cudaMemcpyAsync(s.d_id_off.p, ri->id_off, (n + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, s.stream));
        q.records = 1;
        q.ids = s.d_ids.p - ri->id_off[0];   // (id_off values are used as they are)
        q.id_off = s.d_id_off.p;
        q.read_offset = s.d_offset.p; q.read_median = s.d_median.p;
        const sqg_profile_t &pr = ctx->cfg.profile;
        q.digitisation = pr.digitisation; q.range = pr.range; q.sample_rate = pr.sample_rate;
        q.read_number0 = s.first_read;
        q.start_time0 = ri->start_time0;
        q.ont_friendly = ri->ont_friendly ? 1 : 0;
    }
    CU(cudaMemsetAsync(s.d_seg_state.p, 0, nseg_max * sizeof(unsigned long long), s.stream));
    CU(cudaMemsetAsync(s.d_ticket.p, 0, 4 * sizeof(unsigned int), s.stream));
    CU(cudaMemsetAsync(s.d_read_d0.p, 0xFF, n * sizeof(unsigned long long), s.stream));
    svb_layout_kernel<<<1, 1024, 0, s.stream>>>(q);
    svb_segmap_kernel<<<(int)((n + 7) / 8), 256, 0, s.stream>>>(q);
    int svb_ctas = 8;   // CTAs per SM launched (persistent by ticket; five are resident at 48 registers: measured 2/3/4/5 resident = 1.36/1.06/0.92/0.85 ms, six with spills 0.90)
    if (const char *e = getenv("SQG_SVB_CTAS")) svb_ctas = std::max(1, atoi(e));   // (experiments)
    svb_encode_kernel<<<ctx->num_sms * svb_ctas, SVB_THREADS, 0, s.stream>>>(q);
    svb_finish_kernel<<<(int)((n + 7) / 8), 256, 0, s.stream>>>(q);
    if (records) svb_start_time_kernel<<<1, 1024, 0, s.stream>>>(q);
    publish_kernel<<<1, 32, 0, s.stream>>>(s.d_svb_tot.p, 1, nullptr, 0, s.h_svb_tot.p);
    ctx->launches += records ? 6 : 5;
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(s.stream));
    s.svb_bytes = s.h_svb_tot.p[0];
    if (s.svb_bytes < 0 || (size_t)s.svb_bytes > cap) return fail(ctx, SQG_ERR_CUDA, "svb-zd encoder: output beyond its bound");
    return SQG_OK;
}

// SQG_WANT_SS_TEXT: the dwell strings of the batch's reads, in HBM (sqg_sstext.cuh).
int slot_sstext(sqg_ctx *ctx, Slot &s) {
    s.sst_bytes = 0;
    if (s.n_reads == 0) return SQG_OK;
    const size_t n = (size_t)s.n_reads;
    CU(s.d_sst_len.ensure(n, false, s.stream));
    CU(s.d_sst_off.ensure(n + 1, false, s.stream));
    CU(s.d_ss_off.ensure(n + 1, false, s.stream));
    CU(s.h_sst_off.ensure(n + 1));
    CU(cudaMemcpyAsync(s.d_ss_off.p, s.h_ss_off.p, (n + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, s.stream));
    SsTextParams q;
    q.ss = s.d_ss.p; q.ss_off = s.d_ss_off.p; q.len = s.d_sst_len.p; q.off = s.d_sst_off.p; q.text = nullptr;
    q.n_reads = (int32_t)s.n_reads; q.reversed = ctx->rev ? 1 : 0;
    sstext_size_kernel<<<(int)s.n_reads, SST_THREADS, 0, s.stream>>>(q);
    sstext_offsets_kernel<<<1, 1024, 0, s.stream>>>(q);
    publish_kernel<<<1, 32, 0, s.stream>>>(s.d_sst_off.p + n, 1, nullptr, 0, s.h_sst_off.p + n);
    CU(cudaStreamSynchronize(s.stream));
    s.sst_bytes = s.h_sst_off.p[n];
    CU(s.d_sst.ensure((size_t)std::max<int64_t>(s.sst_bytes, 16), false, s.stream));
    CU(s.h_sst.ensure((size_t)std::max<int64_t>(s.sst_bytes, 16)));
    q.text = s.d_sst.p;
    sstext_write_kernel<<<(int)s.n_reads, SST_THREADS, 0, s.stream>>>(q);
    CU(cudaMemcpyAsync(s.h_sst.p, s.d_sst.p, (size_t)s.sst_bytes, cudaMemcpyDeviceToHost, s.stream));
    CU(cudaMemcpyAsync(s.h_sst_off.p, s.d_sst_off.p, (n + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost, s.stream));
    ctx->launches += 4;
    CU(cudaGetLastError());
    return SQG_OK;
}

void fill_result(Slot &s, sqg_result_t *res) {
    const bool svb = (s.want & (SQG_WANT_SVB | SQG_WANT_RECORDS)) != 0;
    const bool want_bases = s.from_coords && (s.want & SQG_WANT_BASES);
    res->n_reads = s.n_reads;
    res->total_samples = s.total_samples;
    res->signal = svb ? nullptr : s.h_sig.p;
    res->sig_off = s.h_sigoff.p;
    res->len_raw_signal = s.h_len64.p;
    res->offset = s.h_offset.p;
    res->median_before = s.h_median.p;
    res->ss = (s.want & SQG_WANT_SS) ? s.h_ss.p : nullptr;
    res->ss_off = s.h_ss_off.p;
    res->svb = svb ? s.h_svb.p : nullptr;
    res->svb_off = svb ? s.h_svb_off.p : nullptr;
    res->svb_len = svb ? s.h_svb_len.p : nullptr;
    res->ss_text = (s.want & SQG_WANT_SS_TEXT) ? s.h_sst.p : nullptr;
    res->ss_text_off = (s.want & SQG_WANT_SS_TEXT) ? s.h_sst_off.p : nullptr;
    res->bases = want_bases ? s.h_bases.p : nullptr;
    res->bases_off = want_bases ? s.h_base_off.p : nullptr;
    res->meth_draws = s.from_coords ? s.meth_draws : 0;
}

// D2H of everything the caller gets back; fills *res.  Synchronises the slot's stream.
int slot_fetch(sqg_ctx *ctx, Slot &s, sqg_result_t *res) {
    const size_t n = (size_t)s.n_reads;
    const bool svb = (s.want & (SQG_WANT_SVB | SQG_WANT_RECORDS)) != 0;
    if (svb) {
        CU(s.h_svb_len.ensure(n + 1));
        CU(s.h_svb_off.ensure(n + 1));
        CU(s.h_svb.ensure((size_t)std::max<int64_t>(s.svb_bytes, 16)));
        if (n) {
            CU(cudaMemcpyAsync(s.h_svb.p, s.d_svb.p, (size_t)s.svb_bytes, cudaMemcpyDeviceToHost, s.stream));
            CU(cudaMemcpyAsync(s.h_svb_len.p, s.d_svb_len.p, n * sizeof(int64_t), cudaMemcpyDeviceToHost, s.stream));
            CU(cudaMemcpyAsync(s.h_svb_off.p, s.d_svb_off.p, (n + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost, s.stream));
        } else {
            s.h_svb_off.p[0] = 0;
        }
    }
    CU(s.h_siglen.ensure(n + 1));
    CU(s.h_sigoff.ensure(n + 1));
    CU(s.h_len64.ensure(n + 1));
    CU(s.h_offset.ensure(n + 1));
    CU(s.h_median.ensure(n + 1));
    CU(s.h_sig.ensure(svb ? 64 : (size_t)std::max<int64_t>(s.arena_need, 64)));   // (svb-zd: the raw signal stays in HBM)
    if (n) {
        if (!svb) CU(cudaMemcpyAsync(s.h_sig.p, s.d_sig.p, (size_t)s.arena_need * sizeof(int16_t), cudaMemcpyDeviceToHost, s.stream));
        CU(cudaMemcpyAsync(s.h_siglen.p, s.d_siglen.p, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, s.stream));
        CU(cudaMemcpyAsync(s.h_sigoff.p, s.d_sigoff.p, n * sizeof(int64_t), cudaMemcpyDeviceToHost, s.stream));
        CU(cudaMemcpyAsync(s.h_offset.p, s.d_offset.p, n * sizeof(double), cudaMemcpyDeviceToHost, s.stream));
        CU(cudaMemcpyAsync(s.h_median.p, s.d_median.p, n * sizeof(double), cudaMemcpyDeviceToHost, s.stream));
        if (s.want & SQG_WANT_SS) {
            CU(s.h_ss.ensure((size_t)std::max<int64_t>(s.total_kmers, 1)));
            CU(cudaMemcpyAsync(s.h_ss.p, s.d_ss.p, (size_t)s.total_kmers * sizeof(int32_t), cudaMemcpyDeviceToHost, s.stream));
        }
    }
    const bool want_bases = s.from_coords && (s.want & SQG_WANT_BASES);
    if (want_bases) {
        CU(s.h_bases.ensure((size_t)std::max<int64_t>(s.total_bases, 1)));
        if (s.total_bases)
            CU(cudaMemcpyAsync(s.h_bases.p, s.d_bases.p + CONST_REGION, (size_t)s.total_bases, cudaMemcpyDeviceToHost, s.stream));
    }
    CU(cudaStreamSynchronize(s.stream));
    for (size_t i = 0; i < n; i++) s.h_len64.p[i] = (int64_t)s.h_siglen.p[i];
    if (s.from_coords) s.meth_draws = (int64_t)((s.h_cg.p[0] + s.h_cg.p[1]) & 0xFFFFFFFFull);  // low half: CpG sites
    if (res) fill_result(s, res);
    return SQG_OK;
}

int slot_run_all(sqg_ctx *ctx, Slot &s, int64_t n_reads, const char *bases, const int64_t *base_off,
                 int64_t first_read, uint32_t want, sqg_result_t *res, const sqg_coord_t *coords = nullptr,
                 int64_t meth_draw_base = 0) {
    int rc;
    {
        NvtxRange r("sqg:prepare (segments, H2D)");
        if ((rc = slot_prepare(ctx, s, n_reads, bases, base_off, first_read, want, coords, meth_draw_base)) != SQG_OK) return rc;
    }
    {
        NvtxRange r("sqg:plan (tiles, dwells, offsets)");
        if ((rc = slot_plan(ctx, s)) != SQG_OK) return rc;
        if ((rc = slot_size_arena(ctx, s)) != SQG_OK) return rc;
    }
    {
        NvtxRange r("sqg:generate (signal kernel)");
        if ((rc = slot_generate(ctx, s)) != SQG_OK) return rc;
    }
    if (want & (SQG_WANT_SVB | SQG_WANT_RECORDS | SQG_WANT_SS_TEXT)) {
        NvtxRange r("sqg:compress (svb-zd / records, ss text)");
        if ((want & (SQG_WANT_SVB | SQG_WANT_RECORDS)) && (rc = slot_compress(ctx, s)) != SQG_OK) return rc;
        if ((want & SQG_WANT_SS_TEXT) && (rc = slot_sstext(ctx, s)) != SQG_OK) return rc;
    }
    NvtxRange r("sqg:fetch (D2H)");
    return slot_fetch(ctx, s, res);
}

void worker_main(sqg_ctx *ctx, int slot_idx) {
    cudaSetDevice(ctx->device);
    for (;;) {
        Job *job = nullptr;
        {
            std::unique_lock<std::mutex> lk(ctx->mu);
            ctx->cv_work.wait(lk, [&] { return ctx->stopping || !ctx->queues[slot_idx].empty(); });
            if (ctx->stopping && ctx->queues[slot_idx].empty()) return;
            job = ctx->queues[slot_idx].front();
            ctx->queues[slot_idx].pop_front();
        }
        tl_err_sink = &job->err;   // (this thread's CU()/fail() messages belong to the job)
        ctx->slots[slot_idx].rec_info = job->rec;
        int rc = slot_run_all(ctx, ctx->slots[slot_idx], job->n_reads, job->bases, job->base_off, job->first_read,
                              job->want, nullptr, job->coords, job->meth_draw_base);
        tl_err_sink = nullptr;
        {
            std::lock_guard<std::mutex> lk(ctx->mu);
            job->status = rc;
        }
        ctx->cv_done.notify_all();
    }
}

int ctx_setup(sqg_ctx *ctx, const sqg_config_t *cfg) {
    ctx->cfg = *cfg;
    const sqg_profile_t &pr = cfg->profile;
    if (cfg->kmer_size < 1 || cfg->kmer_size > 9) return fail(ctx, SQG_ERR_ARG, "kmer_size must be 1..9 (MAX_KMER_SIZE, src/sq.h:17)");
    uint64_t expect = 1;
    for (uint32_t i = 0; i < cfg->kmer_size; i++) expect *= cfg->meth ? 5 : 4;
    if (cfg->num_kmer != expect) return fail(ctx, SQG_ERR_ARG, "num_kmer must be 4^k (or 5^k with meth)");
    if (!(pr.range > 0) || !(pr.digitisation > 0) || !(pr.dwell_mean >= 1) || pr.dwell_std < 0)
        return fail(ctx, SQG_ERR_ARG, "profile: need range>0, digitisation>0, dwell_mean>=1, dwell_std>=0");
    if (cfg->rng_mode != SQG_RNG_PHILOX && cfg->rng_mode != SQG_RNG_LEGACY) return fail(ctx, SQG_ERR_ARG, "unknown rng_mode");
    ctx->legacy = cfg->rng_mode == SQG_RNG_LEGACY;
    if (ctx->legacy && cfg->seed < 1) return fail(ctx, SQG_ERR_ARG, "SQG_RNG_LEGACY needs seed >= 1");

    const bool ideal = cfg->flags & SQG_IDEAL;
    ctx->noisy = !(ideal || (cfg->flags & SQG_IDEAL_AMP));
    // dwell_std == 0 makes round(N(mean,0)) a constant (rna004 presets, src/sim.c:136-137)
    const bool fixed = ideal || (cfg->flags & SQG_IDEAL_TIME);
    ctx->rand_dwell = !fixed;
    ctx->meth = cfg->meth != 0;
    ctx->rev = cfg->flags & SQG_RNA;
    ctx->prefix = cfg->flags & SQG_PREFIX;

    GenParams &b = ctx->base;
    memset(&b, 0, sizeof b);
    b.k = (int32_t)cfg->kmer_size;
    b.num_kmer = cfg->num_kmer;
    if (cfg->meth) {
        uint32_t p5 = 1;
        for (uint32_t i = 0; i + 1 < cfg->kmer_size; i++) p5 *= 5;
        b.kmask = p5;
    } else {
        b.kmask = (uint32_t)((1ull << (2 * cfg->kmer_size)) - 1);
    }
    b.digitisation = pr.digitisation; b.range = pr.range; b.scale = pr.digitisation / pr.range;
    b.offset_mean = pr.offset_mean; b.offset_std = pr.offset_std;
    b.median_mean = pr.median_before_mean; b.median_std = pr.median_before_std;
    b.dwell_mean = (float)pr.dwell_mean; b.dwell_std = (float)pr.dwell_std;
    b.sps_fixed = (int32_t)pr.dwell_mean;
    b.ideal = ideal ? 1 : 0;
    b.amp_noise = cfg->amp_noise;
    b.key0 = (uint32_t)((uint64_t)cfg->seed & 0xFFFFFFFFu);
    b.key1 = (uint32_t)((uint64_t)cfg->seed >> 32);
    b.shift_val = (int32_t)(int16_t)(30 * pr.digitisation / pr.range);  // src/genread.c:82

    if (ctx->rand_dwell) {
        if (pr.dwell_std == 0.0) {
            // constant dwell: no draw needed, but it is round(dwell_mean), not (int)dwell_mean
            int d = (int)std::round(pr.dwell_mean);
            if (d < 1) d = -d + 1;
            b.sps_fixed = d;
            ctx->rand_dwell = false;
        }
    }
    b.par_cap = PAR_N;
    b.tile_s_cap = 0xFFFFFFFFu;
    if (cfg->meth) b.pow5k = b.kmask * 5u;
    if (ctx->rand_dwell) {
        // largest possible dwell: |z| <= Z_MAX, folded values included
        const double mx = std::floor((double)b.dwell_mean + (double)Z_MAX * (double)b.dwell_std + 0.5) + 2.0;
        if (!(mx < 1200.0)) return fail(ctx, SQG_ERR_ARG, "dwell_mean + 6.06*dwell_std too large for one tile");
        // k-mers per tile: as many as keep a tile's samples inside the signal kernel's window even for a six-sigma run of
        // long dwells; a tile beyond TILE_S_CAP would still be generated
        // correctly, by that kernel's slow path.  T * mx < 2^16: the dwell kernel stores 16-bit prefixes.
        int T = TK;
        // moments of the folded normal |N(mean, std)| (the fold of draws below 1 moves them by one sample at most)
        const double mu = (double)b.dwell_mean, sg = (double)b.dwell_std;
        const double dm = sg * std::sqrt(2.0 / M_PI) * std::exp(-mu * mu / (2 * sg * sg)) + mu * std::erf(mu / (sg * std::sqrt(2.0))) + 1.0;
        const double ds = std::sqrt(std::max(mu * mu + sg * sg - (dm - 1.0) * (dm - 1.0), 0.0));
        while (T > 8 && (T * dm + 6.0 * ds * std::sqrt((double)T) > (double)TILE_S_CAP || T * mx >= 65000.0)) T -= 8;
        b.T = T;
        b.tile_s_cap = TILE_S_CAP;
        // a chunk of 8 samples holds three k-mers only if one of them has a dwell <= 6: the warp-wide vote pays when that is rare
        b.l2_vote = (mu - 6.0) / std::max(sg, 1e-9) > 1.3 ? 1 : 0;
        if (const char *e = getenv("SQG_L2_VOTE")) b.l2_vote = atoi(e);   // (experiments)
    } else {
        // n / sps == umulhi(n, magic) needs n * sps < 2^32 for every window sample n < par_cap * sps
        const uint64_t sps = (uint64_t)b.sps_fixed;
        if (sps > 16000) return fail(ctx, SQG_ERR_ARG, "dwell_mean too large");
        const uint64_t kcap = std::min<uint64_t>(PAR_N, 0xFFFFFFFFull / (sps * sps));   // k-mers in the window at most
        if (kcap < 12) return fail(ctx, SQG_ERR_ARG, "dwell_mean too large");
        b.par_cap = (int32_t)kcap;
        b.T = (int32_t)std::min<uint64_t>(TK, (kcap - 4) & ~7ull);
        b.sps_magic = sps == 1 ? 0u : (uint32_t)(0x100000000ull / sps) + 1u;  // sps == 1 is special-cased in div_sps()
    }
    if (cfg->reserved > 0) {
        // testing knob: shrink the window (low 16 bits: samples per tile the fast path accepts; high bits: k-mers) so that
        // small inputs reach the cut-run and slow-tile paths
        const uint32_t sc = (uint32_t)cfg->reserved & 0xFFFFu, kc = (uint32_t)cfg->reserved >> 16;
        if (sc && ctx->rand_dwell) b.tile_s_cap = std::min<uint32_t>(b.tile_s_cap, sc);
        if (kc) b.par_cap = std::max<int32_t>(b.T + 2, std::min<int32_t>(b.par_cap, (int32_t)kc));
    }
    for (int r = 0; r < PHILOX_ROUNDS; r++) {
        b.rk[2 * r] = b.key0 + (uint32_t)r * 0x9E3779B9u;
        b.rk[2 * r + 1] = b.key1 + (uint32_t)r * 0xBB67AE85u;
    }
    return SQG_OK;
}

int ctx_device_setup(sqg_ctx *ctx, const sqg_model_t *h_model, const void *d_model_in) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(ctx, SQG_ERR_NODEVICE, "no CUDA device: libsqg has no CPU fallback");
    if (ctx->cfg.device < 0 || ctx->cfg.device >= ndev) return fail(ctx, SQG_ERR_ARG, "bad device ordinal");
    ctx->device = ctx->cfg.device;
    CU(cudaSetDevice(ctx->device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, ctx->device));
    if (prop.major < 10) return fail(ctx, SQG_ERR_NODEVICE, "libsqg is built for sm_100a (B200) only");
    ctx->num_sms = prop.multiProcessorCount;
    const size_t n = ctx->cfg.num_kmer;
    CU(ctx->d_model.ensure(n));
    if (h_model) CU(cudaMemcpy(ctx->d_model.p, h_model, n * sizeof(float2), cudaMemcpyHostToDevice));
    else CU(cudaMemcpy(ctx->d_model.p, d_model_in, n * sizeof(float2), cudaMemcpyDeviceToDevice));
    const size_t zbytes = (size_t)Z32_BYTES + Z2_SUB * sizeof(float);
    CU(ctx->d_z.ensure(zbytes));
    CU(cudaMemcpy(ctx->d_z.p, sqg_ztable_blob, zbytes, cudaMemcpyHostToDevice));
    if (ctx->legacy) {
        CU(ctx->d_cnt_kmer.ensure(n));
        CU(cudaMemset(ctx->d_cnt_kmer.p, 0, n * sizeof(uint64_t)));
    }
    ctx->base.model = ctx->d_model.p;
    if (!ctx->legacy) {
        // the model in the form the sample arithmetic uses: (A', M) by rank (model_am_kernel)
        CU(ctx->d_model_am.ensure(n));
        model_am_kernel<<<(unsigned)((n + 255) / 256), 256>>>(ctx->d_model.p, ctx->d_model_am.p, (uint32_t)n, ctx->cfg.amp_noise, (float)ctx->base.scale);
        CU(cudaGetLastError());
        ctx->base.model_am = ctx->d_model_am.p;
    }
    if (!ctx->legacy && ctx->noisy) {
        // the same, indexed by (k+1)-mer: one 16-byte gather serves two consecutive k-mers (pair_model_kernel)
        const uint64_t n_pair = (ctx->meth ? 5ull : 4ull) * n;
        if (n_pair * sizeof(float4) > (256ull << 20)) return fail(ctx, SQG_ERR_ARG, "k-mer size too large for the paired model table");
        // (+8: a pair whose second k-mer lies past its tile is addressed with one digit of whatever follows the window)
        CU(ctx->d_pair_model.ensure((size_t)n_pair + 8));
        CU(cudaMemset(ctx->d_pair_model.p + n_pair, 0, 8 * sizeof(float4)));
        if (ctx->meth) pair_model5_kernel<<<(unsigned)((n_pair + 255) / 256), 256>>>(ctx->d_model_am.p, ctx->d_pair_model.p, (uint32_t)n_pair, ctx->base.pow5k);
        else pair_model_kernel<<<(unsigned)((n_pair + 255) / 256), 256>>>(ctx->d_model_am.p, ctx->d_pair_model.p, (uint32_t)n_pair, ctx->base.kmask);
        CU(cudaGetLastError());
        ctx->base.pair_model = ctx->d_pair_model.p;
    }
    CU(cudaDeviceSynchronize());
    ctx->base.z32 = reinterpret_cast<const float *>(ctx->d_z.p);
    ctx->base.z2 = reinterpret_cast<const float *>(ctx->d_z.p + (size_t)Z32_BYTES);

    // The sample kernel reads the int16 straight out of the mantissa of fma.rz(z, A', B' + 32768) and relies on bit 23
    // to flag everything outside [0, 32768); that flag is reliable while 16384 <= z*A' + B' + 32768 < 131072 for every
    // k-mer, every |z| <= Z_MAX and every offset a read can draw.  Check it once over the model; otherwise (absurd
    // profiles or models) every sample takes the exact path (GenParams::wide).
    if (ctx->noisy && !ctx->legacy) {
        std::vector<float2> hm(n);
        CU(cudaMemcpy(hm.data(), ctx->d_model.p, n * sizeof(float2), cudaMemcpyDeviceToHost));
        const sqg_profile_t &pr = ctx->cfg.profile;
        const double scale = pr.digitisation / pr.range;
        const double off_span = 12.0 * std::fabs(pr.offset_std);  // |deviate| <= (sum of the four weights) * Z_MAX = 11.6
        const bool ideal = (ctx->cfg.flags & SQG_IDEAL) != 0;
        const double off_lo = pr.offset_mean - (ideal ? 0.0 : off_span), off_hi = pr.offset_mean + (ideal ? 0.0 : off_span);
        double lo = 1e300, hi = -1e300;
        bool finite = std::isfinite(scale) && std::isfinite(off_lo) && std::isfinite(off_hi) && std::isfinite((double)ctx->cfg.amp_noise);
        for (size_t i = 0; i < n && finite; i++) {
            const double a = std::fabs((double)hm[i].y * (double)ctx->cfg.amp_noise * scale) * ((double)Z_MAX + 0.01) + 2.0;
            const double m = (double)hm[i].x * scale;
            if (!std::isfinite(a) || !std::isfinite(m)) { finite = false; break; }
            lo = std::min(lo, m - off_hi - a);
            hi = std::max(hi, m - off_lo + a);
        }
        ctx->base.wide = (finite && lo > -16000.0 && hi < 98000.0) ? 0 : 1;
    }

    CU(cudaFuncSetAttribute((const void *)dwell_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)K1_SMEM));
    // the svb-zd encoder stages a segment per CTA in (static) shared memory: let the SM carve out what 6-8 resident CTAs need
    CU(cudaFuncSetAttribute((const void *)svb_encode_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    k4_fn fn = pick_k4(ctx->noisy, ctx->rand_dwell, ctx->meth, ctx->rev);
    if (SM_TOTAL > prop.sharedMemPerBlockOptin) return fail(ctx, SQG_ERR_CUDA, "signal kernel: shared-memory layout does not fit");
    {
        // the sample loop addresses shared memory absolutely: dynamic array = reserved kilobyte + no static shared memory
        int reserved = 0;
        cudaFuncAttributes fa;
        CU(cudaDeviceGetAttribute(&reserved, cudaDevAttrReservedSharedMemoryPerBlock, ctx->device));
        CU(cudaFuncGetAttributes(&fa, (const void *)fn));
        if ((uint32_t)reserved + (uint32_t)fa.sharedSizeBytes != SMEM_ORIGIN)
            return fail(ctx, SQG_ERR_CUDA, "signal kernel: unexpected origin of dynamic shared memory");
    }
    CU(cudaFuncSetAttribute((const void *)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM_TOTAL));
    int occ = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, (const void *)fn, K4_THREADS, SM_TOTAL));
    if (occ < 1) return fail(ctx, SQG_ERR_CUDA, "signal kernel does not fit on an SM");
    ctx->k4_grid_per_sm = occ;

    int rc = slot_init(ctx, ctx->sync_slot);
    if (rc != SQG_OK) return rc;
    return SQG_OK;
}

int init_common(sqg_ctx **out, const sqg_config_t *cfg, const sqg_model_t *h_model, const void *d_model) {
    if (!out || !cfg || (!h_model && !d_model)) return fail(nullptr, SQG_ERR_ARG, "null argument");
    *out = nullptr;
    sqg_ctx *ctx = new (std::nothrow) sqg_ctx();
    if (!ctx) return fail(nullptr, SQG_ERR_NOMEM, "out of host memory");
    int rc = ctx_setup(ctx, cfg);
    if (rc == SQG_OK) rc = ctx_device_setup(ctx, h_model, d_model);
    if (rc != SQG_OK) {
        g_init_error = ctx->err;
        sqg_destroy(ctx);
        return rc;
    }
    *out = ctx;
    return SQG_OK;
}

int ensure_dispatcher(sqg_ctx *ctx) {
    if (!ctx->slots.empty()) return SQG_OK;
    // legacy streams are consumed in submission order: one slot, one worker
    const int n = ctx->legacy ? 1 : (ctx->cfg.n_slots > 0 ? std::min(ctx->cfg.n_slots, 16) : 3);
    ctx->slots.resize(n);
    ctx->queues.resize(n);
    ctx->slot_busy.assign(n, 0);
    for (int i = 0; i < n; i++) {
        int rc = slot_init(ctx, ctx->slots[i]);
        if (rc != SQG_OK) {   // no half-built dispatcher: the next sqg_submit starts over
            for (auto &s : ctx->slots) s.release();
            ctx->slots.clear();
            ctx->queues.clear();
            ctx->slot_busy.clear();
            return rc;
        }
    }
    for (int i = 0; i < n; i++) ctx->workers.emplace_back(worker_main, ctx, i);
    return SQG_OK;
}

}  // namespace

// =================================================================================================
extern "C" {

const char *sqg_version(void) { return "squigulator-b200 0.1.0 (sm_100a)"; }

const char *sqg_last_error(const sqg_ctx_t *ctx) { return ctx ? ctx->err.c_str() : g_init_error.c_str(); }

int sqg_init(sqg_ctx_t **ctx, const sqg_config_t *cfg, const sqg_model_t *model) {
    return init_common(ctx, cfg, model, nullptr);
}

int sqg_init_device_model(sqg_ctx_t **ctx, const sqg_config_t *cfg, const void *d_model) {
    return init_common(ctx, cfg, nullptr, d_model);
}

void sqg_destroy(sqg_ctx_t *ctx) {
    if (!ctx) return;
    {
        std::lock_guard<std::mutex> lk(ctx->mu);
        ctx->stopping = true;
    }
    ctx->cv_work.notify_all();
    for (auto &t : ctx->workers) t.join();
    cudaSetDevice(ctx->device);
    for (auto &kv : ctx->jobs) delete kv.second;
    for (auto &s : ctx->slots) s.release();
    ctx->sync_slot.release();
    ctx->d_model.release();
    ctx->d_model_am.release();
    if (ctx->d_pair_model.p) cudaCtxResetPersistingL2Cache();   // (slot_pin_model: the table's lines go back to the normal policy)
    ctx->d_pair_model.release();
    ctx->d_z.release();
    ctx->d_cnt_kmer.release();
    ctx->d_genome.release(); ctx->d_gmeth.release(); ctx->d_has_meth.release(); ctx->d_contig_off.release();
    delete ctx;
}

void *sqg_host_alloc(size_t bytes) {
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) return nullptr;
    return p;
}

void sqg_host_free(void *p) {
    if (p) cudaFreeHost(p);
}

int sqg_gen_batch(sqg_ctx_t *ctx, int64_t n_reads, const char *bases, const int64_t *base_off,
                  int64_t first_read_index, uint32_t want, sqg_result_t *res) {
    if (!ctx || !res) return SQG_ERR_ARG;
    CU(cudaSetDevice(ctx->device));
    if (ctx->legacy) {   // the reference's streams are consumed in call order: no batch of the dispatcher may be in flight
        std::lock_guard<std::mutex> lk(ctx->mu);
        for (int b : ctx->slot_busy)
            if (b) return fail(ctx, SQG_ERR_STATE, "SQG_RNG_LEGACY: sqg_gen_batch while a submitted batch is in flight");
    }
    return slot_run_all(ctx, ctx->sync_slot, n_reads, bases, base_off, first_read_index, want, res);
}

static int submit_job(sqg_ctx_t *ctx, int64_t n_reads, const char *bases, const int64_t *base_off,
                      const sqg_coord_t *coords, int64_t meth_draw_base, int64_t first_read_index, uint32_t want,
                      sqg_ticket_t *ticket, const sqg_record_info_t *rec = nullptr) {
    if (!ctx || !ticket) return SQG_ERR_ARG;
    CU(cudaSetDevice(ctx->device));
    int rc = ensure_dispatcher(ctx);
    if (rc != SQG_OK) return rc;
    std::unique_lock<std::mutex> lk(ctx->mu);
    int slot = -1;
    ctx->cv_done.wait(lk, [&] {
        for (size_t i = 0; i < ctx->slot_busy.size(); i++)
            if (!ctx->slot_busy[i]) { slot = (int)i; return true; }
        return false;
    });
    Job *job = new Job{ctx->next_ticket++, slot, n_reads, bases, base_off, first_read_index, want, 1, coords, meth_draw_base};
    job->rec = rec;
    ctx->slot_busy[slot] = 1;
    ctx->jobs[job->ticket] = job;
    ctx->queues[slot].push_back(job);
    *ticket = job->ticket;
    lk.unlock();
    ctx->cv_work.notify_all();
    return SQG_OK;
}

int sqg_submit(sqg_ctx_t *ctx, int64_t n_reads, const char *bases, const int64_t *base_off,
               int64_t first_read_index, uint32_t want, sqg_ticket_t *ticket) {
    return submit_job(ctx, n_reads, bases, base_off, nullptr, 0, first_read_index, want, ticket);
}

int sqg_gen_batch_records(sqg_ctx_t *ctx, int64_t n_reads, const char *bases, const int64_t *base_off,
                          int64_t first_read_index, const sqg_record_info_t *rec, uint32_t want, sqg_result_t *res) {
    if (!ctx || !res) return SQG_ERR_ARG;
    if (n_reads > 0 && !rec) return fail(ctx, SQG_ERR_ARG, "null record info");
    CU(cudaSetDevice(ctx->device));
    ctx->sync_slot.rec_info = rec;
    const int rc = slot_run_all(ctx, ctx->sync_slot, n_reads, bases, base_off, first_read_index, want | SQG_WANT_RECORDS, res);
    ctx->sync_slot.rec_info = nullptr;
    return rc;
}

int sqg_submit_records(sqg_ctx_t *ctx, int64_t n_reads, const char *bases, const int64_t *base_off,
                       int64_t first_read_index, const sqg_record_info_t *rec, uint32_t want, sqg_ticket_t *ticket) {
    if (n_reads > 0 && !rec) return ctx ? fail(ctx, SQG_ERR_ARG, "null record info") : SQG_ERR_ARG;
    return submit_job(ctx, n_reads, bases, base_off, nullptr, 0, first_read_index, want | SQG_WANT_RECORDS, ticket, rec);
}

int sqg_submit_coords(sqg_ctx_t *ctx, int64_t n_reads, const sqg_coord_t *coords, int64_t first_read_index,
                      int64_t meth_draw_base, uint32_t want, sqg_ticket_t *ticket) {
    if (n_reads > 0 && !coords) return ctx ? fail(ctx, SQG_ERR_ARG, "null coordinates") : SQG_ERR_ARG;
    static const sqg_coord_t none = {0, 0, 0, '+', 0};
    return submit_job(ctx, n_reads, nullptr, nullptr, coords ? coords : &none, meth_draw_base, first_read_index, want, ticket);
}

int sqg_gen_batch_coords(sqg_ctx_t *ctx, int64_t n_reads, const sqg_coord_t *coords, int64_t first_read_index,
                         int64_t meth_draw_base, uint32_t want, sqg_result_t *res) {
    if (!ctx || !res) return SQG_ERR_ARG;
    if (n_reads > 0 && !coords) return fail(ctx, SQG_ERR_ARG, "null coordinates");
    static const sqg_coord_t none = {0, 0, 0, '+', 0};
    CU(cudaSetDevice(ctx->device));
    return slot_run_all(ctx, ctx->sync_slot, n_reads, nullptr, nullptr, first_read_index, want, res,
                        coords ? coords : &none, meth_draw_base);
}

int sqg_genome_load(sqg_ctx_t *ctx, int32_t n_contigs, const char *seq, const int64_t *contig_off,
                    const uint8_t *meth, const uint8_t *contig_has_meth) {
    if (!ctx) return SQG_ERR_ARG;
    if (n_contigs < 1 || !seq || !contig_off) return fail(ctx, SQG_ERR_ARG, "genome: need >= 1 contig, seq and contig_off");
    for (int32_t c = 0; c < n_contigs; c++)
        if (contig_off[c + 1] < contig_off[c]) return fail(ctx, SQG_ERR_ARG, "genome: contig_off must be non-decreasing");
    if (contig_off[0] != 0) return fail(ctx, SQG_ERR_ARG, "genome: contig_off[0] must be 0");
    CU(cudaSetDevice(ctx->device));
    {  // no batch may be in flight while the genome is replaced
        std::unique_lock<std::mutex> lk(ctx->mu);
        for (int b : ctx->slot_busy)
            if (b) return fail(ctx, SQG_ERR_STATE, "genome load while batches are in flight");
    }
    const size_t total = (size_t)contig_off[n_contigs];
    CU(cudaDeviceSynchronize());
    CU(ctx->d_genome.ensure(total + 64));
    CU(ctx->d_contig_off.ensure((size_t)n_contigs + 1));
    CU(cudaMemcpy(ctx->d_genome.p, seq, total, cudaMemcpyHostToDevice));
    CU(cudaMemset(ctx->d_genome.p + total, 0, 64));
    CU(cudaMemcpy(ctx->d_contig_off.p, contig_off, ((size_t)n_contigs + 1) * sizeof(int64_t), cudaMemcpyHostToDevice));
    ctx->genome_meth = meth != nullptr;
    ctx->genome_has_flags = meth != nullptr && contig_has_meth != nullptr;
    if (meth) {
        CU(ctx->d_gmeth.ensure(total + 64));
        CU(cudaMemcpy(ctx->d_gmeth.p, meth, total, cudaMemcpyHostToDevice));
        if (contig_has_meth) {
            CU(ctx->d_has_meth.ensure((size_t)n_contigs));
            CU(cudaMemcpy(ctx->d_has_meth.p, contig_has_meth, (size_t)n_contigs, cudaMemcpyHostToDevice));
        }
    }
    ctx->contig_off.assign(contig_off, contig_off + n_contigs + 1);
    return SQG_OK;
}

int sqg_wait(sqg_ctx_t *ctx, sqg_ticket_t ticket, sqg_result_t *res) {
    if (!ctx) return SQG_ERR_ARG;
    std::unique_lock<std::mutex> lk(ctx->mu);
    auto it = ctx->jobs.find(ticket);
    if (it == ctx->jobs.end()) return fail(ctx, SQG_ERR_STATE, "unknown ticket");
    Job *job = it->second;
    ctx->cv_done.wait(lk, [&] { return job->status <= 0; });
    if (job->status != SQG_OK) {
        ctx->err = job->err;   // (under ctx->mu: sqg_last_error on the waiting thread sees this job's message)
        return job->status;
    }
    if (res) fill_result(ctx->slots[job->slot], res);
    return SQG_OK;
}

int sqg_release(sqg_ctx_t *ctx, sqg_ticket_t ticket) {
    if (!ctx) return SQG_ERR_ARG;
    {
        std::unique_lock<std::mutex> lk(ctx->mu);
        auto it = ctx->jobs.find(ticket);
        if (it == ctx->jobs.end()) return fail(ctx, SQG_ERR_STATE, "unknown ticket");
        Job *job = it->second;
        ctx->cv_done.wait(lk, [&] { return job->status <= 0; });
        ctx->slot_busy[job->slot] = 0;
        ctx->jobs.erase(it);
        delete job;
    }
    ctx->cv_done.notify_all();
    return SQG_OK;
}

int16_t *sqg_gen_sig(sqg_ctx_t *ctx, const char *read, int32_t len, double *offset, double *median_before,
                     int64_t *len_raw_signal, int64_t read_index, int32_t **ss, int64_t *ss_n) {
    if (!ctx || !read || len < 0) return nullptr;
    const int64_t off[2] = {0, len};
    sqg_result_t res;
    if (sqg_gen_batch(ctx, 1, read, off, read_index, ss ? SQG_WANT_SS : 0, &res) != SQG_OK) return nullptr;
    const int64_t n = res.len_raw_signal[0];
    int16_t *out = (int16_t *)malloc((size_t)std::max<int64_t>(n, 1) * sizeof(int16_t));
    if (!out) return nullptr;
    memcpy(out, res.signal + res.sig_off[0], (size_t)n * sizeof(int16_t));
    if (offset) *offset = res.offset[0];
    if (median_before) *median_before = res.median_before[0];
    if (len_raw_signal) *len_raw_signal = n;
    if (ss) {
        const int64_t nk = res.ss_off[1] - res.ss_off[0];
        *ss = (int32_t *)malloc((size_t)std::max<int64_t>(nk, 1) * sizeof(int32_t));
        if (*ss) memcpy(*ss, res.ss, (size_t)nk * sizeof(int32_t));
        if (ss_n) *ss_n = nk;
    }
    return out;
}

// ---- device-resident batches ----

int sqg_dev_batch_create(sqg_ctx_t *ctx, int64_t n_reads, const char *bases, const int64_t *base_off,
                         int64_t first_read_index, uint32_t want, sqg_dev_batch_t **batch) {
    if (!ctx || !batch) return SQG_ERR_ARG;
    CU(cudaSetDevice(ctx->device));
    sqg_dev_batch *b = new (std::nothrow) sqg_dev_batch();
    if (!b) return fail(ctx, SQG_ERR_NOMEM, "out of host memory");
    int rc = slot_init(ctx, b->slot);
    if (rc == SQG_OK) rc = slot_prepare(ctx, b->slot, n_reads, bases, base_off, first_read_index, want);
    if (rc == SQG_OK) {
        cudaError_t e = cudaStreamSynchronize(b->slot.stream);
        if (e != cudaSuccess) rc = fail(ctx, SQG_ERR_CUDA, cudaGetErrorString(e));
    }
    if (rc != SQG_OK) {
        b->slot.release();
        delete b;
        return rc;
    }
    *batch = b;
    return SQG_OK;
}

int sqg_dev_batch_run(sqg_ctx_t *ctx, sqg_dev_batch_t *b, int32_t steps, float *ms_total, float *ms_signal_kernel) {
    if (!ctx || !b || steps < 1) return SQG_ERR_ARG;
    CU(cudaSetDevice(ctx->device));
    Slot &s = b->slot;
    int rc;
    if (!b->planned) {  // first (untimed) plan sizes the arena; the plan is deterministic, so it holds for every step
        if ((rc = slot_plan(ctx, s)) != SQG_OK) return rc;
        if ((rc = slot_size_arena(ctx, s)) != SQG_OK) return rc;
        b->planned = true;
    }
    while ((int)s.kev.size() < 2 * steps) {
        cudaEvent_t e;
        CU(cudaEventCreate(&e));
        s.kev.push_back(e);
    }
    CU(cudaEventRecord(s.ev0, s.stream));
    for (int i = 0; i < steps; i++) {
        if ((rc = slot_plan(ctx, s)) != SQG_OK) return rc;
        if ((rc = slot_generate(ctx, s, s.kev[2 * i], s.kev[2 * i + 1])) != SQG_OK) return rc;
    }
    CU(cudaEventRecord(s.ev1, s.stream));
    CU(cudaStreamSynchronize(s.stream));
    if (ms_total) CU(cudaEventElapsedTime(ms_total, s.ev0, s.ev1));
    if (ms_signal_kernel) {
        float acc = 0.f;
        for (int i = 0; i < steps; i++) {
            float t = 0.f;
            CU(cudaEventElapsedTime(&t, s.kev[2 * i], s.kev[2 * i + 1]));
            acc += t;
        }
        *ms_signal_kernel = acc;
    }
    return SQG_OK;
}

int sqg_dev_batch_info(sqg_ctx_t *ctx, sqg_dev_batch_t *b, int64_t *total_samples, int64_t *total_kmers,
                       int64_t *total_bases, int64_t *kernel_launches) {
    if (!ctx || !b) return SQG_ERR_ARG;
    if (total_samples) *total_samples = b->slot.total_samples;
    if (total_kmers) *total_kmers = b->slot.total_kmers;
    if (total_bases) *total_bases = b->slot.total_bases;
    if (kernel_launches) *kernel_launches = ctx->launches.load();
    return SQG_OK;
}

int sqg_dev_batch_fetch(sqg_ctx_t *ctx, sqg_dev_batch_t *b, sqg_result_t *res) {
    if (!ctx || !b || !res) return SQG_ERR_ARG;
    if (!b->planned) return fail(ctx, SQG_ERR_STATE, "batch has not been run");
    CU(cudaSetDevice(ctx->device));
    if (b->slot.want & (SQG_WANT_SVB | SQG_WANT_RECORDS)) {
        int rc = slot_compress(ctx, b->slot);
        if (rc != SQG_OK) return rc;
    }
    if (b->slot.want & SQG_WANT_SS_TEXT) {
        int rc = slot_sstext(ctx, b->slot);
        if (rc != SQG_OK) return rc;
    }
    return slot_fetch(ctx, b->slot, res);
}

void sqg_dev_batch_destroy(sqg_ctx_t *ctx, sqg_dev_batch_t *b) {
    if (!b) return;
    if (ctx) cudaSetDevice(ctx->device);
    b->slot.release();
    delete b;
}

int sqg_bench_store(sqg_ctx_t *ctx, size_t bytes, int32_t steps, float *ms_total) {
    if (!ctx || !ms_total || steps < 1) return SQG_ERR_ARG;
    CU(cudaSetDevice(ctx->device));
    Slot &s = ctx->sync_slot;
    const size_t n16 = bytes / 16;
    CU(s.d_sig.ensure(n16 * 8, false, s.stream));
    CU(cudaEventRecord(s.ev0, s.stream));
    for (int i = 0; i < steps; i++) {
        store_only_kernel<<<ctx->num_sms * 4, 512, 0, s.stream>>>(reinterpret_cast<uint4 *>(s.d_sig.p), n16);
        ctx->launches++;
    }
    CU(cudaEventRecord(s.ev1, s.stream));
    CU(cudaStreamSynchronize(s.stream));
    CU(cudaEventElapsedTime(ms_total, s.ev0, s.ev1));
    return SQG_OK;
}

int64_t sqg_launch_count(const sqg_ctx_t *ctx) { return ctx ? ctx->launches.load() : 0; }

}  // extern "C"
