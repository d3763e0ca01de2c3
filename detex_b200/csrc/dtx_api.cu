// dtx_api.cu -- the C ABI (include/detex_b200.h) over the kernels.
//
// Host-side responsibilities only: packing subspaces into 16-vector basis blocks, laying
// out the K segments, sizing / growing device buffers, building the work-item list and
// launching K0 -> K1 -> K3 on the context's stream.
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <functional>
#include <map>
#include <string>
#include <utility>
#include <vector>

#include "../../include/detex_b200.h"
#include "dtx_kernels.cuh"

using namespace dtx;

namespace {

// Owning device buffer: grows on demand, frees in the destructor (every early return of an entry
// point releases its temporaries).
template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t cap = 0;  // elements
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    DevBuf(DevBuf&& o) noexcept : p(o.p), cap(o.cap) { o.p = nullptr; o.cap = 0; }
    DevBuf& operator=(DevBuf&& o) noexcept {
        if (this != &o) { release(); p = o.p; cap = o.cap; o.p = nullptr; o.cap = 0; }
        return *this;
    }
    ~DevBuf() { release(); }
    cudaError_t reserve(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = n + n / 8 + 256;
        cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&p), want * sizeof(T));
        if (e != cudaSuccess) {
            want = n;
            e = cudaMalloc(reinterpret_cast<void**>(&p), want * sizeof(T));
        }
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

// Pinned host staging buffer (async D2H target), same growth rule.
template <typename T>
struct PinBuf {
    T* p = nullptr;
    size_t cap = 0;
    PinBuf() = default;
    PinBuf(const PinBuf&) = delete;
    PinBuf& operator=(const PinBuf&) = delete;
    ~PinBuf() { release(); }
    cudaError_t reserve(size_t n) {
        if (n <= cap) return cudaSuccess;
        release();
        cudaError_t e = cudaMallocHost(reinterpret_cast<void**>(&p), n * sizeof(T));
        if (e == cudaSuccess) cap = n;
        return e;
    }
    void release() {
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
};

struct BasisSet {
    BasisLayout lay{};
    int R = 0;
    std::vector<int> rank_off;
    bool has_thr = false;
    bool has_split = false;   // some subspace has rank > 16 (its pieces are summed after K1)
    int ds_rows = 0;          // DS rows per chunk: S + scratch rows of the 2nd.. pieces of rank > 16 subspaces
    std::vector<PieceSum> pieces;   // (subspace row, scratch row) in piece order
    DevBuf<PieceSum> d_pieces;
    DevBuf<double> d_U;
    DevBuf<int> d_rank_off, d_slot_row;
    DevBuf<BlockInfo> d_binfo;
    DevBuf<uint8_t> d_Aimg;
    DevBuf<uint8_t> d_Aimg8;  // image of the 8-bit cross-term engine, built on first use
    bool have_img8 = false;
    double nu4 = 1.0;         // max over basis vectors of sum u^4 / (sum u^2)^2 (precision policy)
    DevBuf<float> d_thr;
    DevBuf<unsigned long long> d_hist;
    DevBuf<double> d_fas;
    // magnitude estimation (N1): one blob per subspace [ewf | mags | mean | std | wfu_var]
    std::map<int, DevBuf<double>> ev_blob;
    std::map<int, std::pair<int, int>> ev_meta;  // subspace -> (nev, is_single)
    void release() {
        for (auto& kv : ev_blob) kv.second.release();
        ev_blob.clear(); ev_meta.clear();
        d_pieces.release(); pieces.clear();
        d_U.release(); d_rank_off.release(); d_slot_row.release(); d_binfo.release();
        d_Aimg.release(); d_Aimg8.release(); have_img8 = false; d_thr.release(); d_hist.release(); d_fas.release();
    }
};

}  // namespace

struct dtx_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int num_sms = 0;
    std::string err;
    std::map<int, BasisSet> sets;

    // current batch
    int nchunks = 0;
    int dtype = DTX_F64;
    std::vector<long long> raw_off;  // element offsets in raw buffer
    std::vector<long long> rawL;
    std::vector<long long> core_lo, core_hi;   // per chunk core lag range (dtx_set_core_lags), empty = all lags
    const void* d_raw = nullptr;     // points into raw_own or caller memory
    DevBuf<uint8_t> raw_own;
    // H2D of the next batch overlaps the projection of the current one: the copies run on a second
    // stream, ordered behind the last kernel that READS the raw chunks (K0 of the previous batch)
    // and in front of everything the main stream does next
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_raw_read = nullptr, ev_copy_done = nullptr;
    bool raw_read_pending = false, copy_pending = false;

    // last run
    int run_set = -1;
    int run_S = 0;
    bool ran = false;
    bool have_ds64 = false;
    std::vector<ChunkDesc> h_chunks;
    DevBuf<ChunkDesc> d_chunks;
    DevBuf<int4> d_items;
    std::vector<int> items_key;   // shape signature of the cached work-item list
    int n_items = 0;
    DevBuf<__half> d_xsplit;
    DevBuf<float> d_mu, d_invE, d_DS, d_scale, d_rowmax;
    DevBuf<double> d_DS64, d_sum;
    DevBuf<unsigned> d_maxbits, d_k4bits;
    DevBuf<int> d_chunk_mode;     // per chunk: 1 = 8-bit cross terms in the last run
    int hist_bins = HIST_BINS;    // bins of the device histograms (numBins - 1 of fas._initFAS, fas.py:31)
    int sta_window = 0;           // triggerSTATime in samples (0 = reference default: STA = |DS|)
    double x8_eps = 2e-6;         // adaptive engine: admitted RMS error of a normalised projection
    DevBuf<int> d_rowflags, d_ncand, d_zeroE, d_chunk_bad;
    DevBuf<double> d_lta_acc;     // fused mode: window sums of |DS| per candidate (kept zero between uses)
    bool fused = false;           // dtx_set_fused: K1 does the row reductions itself, DS is never written
    bool last_fused = false;      // the last run had no dense DS
    bool d_lta_acc_zeroed = false;
    DevBuf<Candidate> d_cand;
    int cand_cap = 1 << 20;
    // K1 timing: one CUDA event pair per run, kept until dtx_k1_ms_history collects them
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> k1_events;   // pool
    size_t k1_used = 0;          // pairs recorded since the last collection
    bool k1_timed = false;
    long long launches = 0;

    // results accumulating over the batches of a station (dtx_accumulate_begin): candidate rows and
    // rowmax / rowflags entries are indexed by the chunk's position in the whole sequence
    bool accumulate = false;
    long long acc_chunks = 0;    // chunks of the batches run so far
    long long acc_capacity = 0;  // chunks the row buffers were sized for

    // CCX: persistent workspace (no allocation per call once warm) and its own basis set
    BasisSet ccx_set;
    DevBuf<uint8_t> cx_X;
    DevBuf<double> cx_wa, cx_wb, cx_es, cx_ed, cx_pad, cx_cc, cx_sub, cx_tcc, cx_tsub, cx_pcc, cx_psub;
    DevBuf<int> cx_lag, cx_tlag, cx_plag, cx_rows, cx_nflag, cx_slot;
    DevBuf<int4> cx_karg;
    DevBuf<double> cx_xd;       // events de-multiplexed into float64 [N][Nc][ns] (ring re-scoring)
    PinBuf<int> cx_nflag_h;      // degenerate-pair count of the last CCX call, checked at the next synchronisation
    bool cx_flag_check = false;
    DevBuf<int2> cx_flag;
    int ccx_passes = 1;          // MMAs per K step of the CCX series: 1 = hi*hi screening (default), 3 = fp16x3
    int ccx_max_batch = 512;     // signals per K1 launch (dtx_set_ccx_batch lowers it for tests)
    long long ccx_ds_bytes = 16LL << 30;  // series buffer of one CCX batch (4096 events: 8 launches of 512 signals)
};

namespace {

int fail(dtx_ctx* c, int code, const std::string& msg) {
    if (c) c->err = msg;
    return code;
}

#define DTX_CUDA(call)                                                                     \
    do {                                                                                   \
        cudaError_t e_ = (call);                                                           \
        if (e_ != cudaSuccess) {                                                           \
            char buf_[512];                                                                \
            snprintf(buf_, sizeof(buf_), "%s failed: %s (%s:%d)", #call,                   \
                     cudaGetErrorString(e_), __FILE__, __LINE__);                          \
            return fail(ctx, DTX_ERR_CUDA, buf_);                                          \
        }                                                                                  \
    } while (0)

int round_up(int v, int m) { return (v + m - 1) / m * m; }

// Called after enqueuing a kernel that reads (or writes) the raw chunk buffer on the main stream.
void mark_raw_access(dtx_ctx* ctx) {
    if (!ctx->ev_raw_read) cudaEventCreateWithFlags(&ctx->ev_raw_read, cudaEventDisableTiming);
    if (ctx->ev_raw_read && cudaEventRecord(ctx->ev_raw_read, ctx->stream) == cudaSuccess) ctx->raw_read_pending = true;
}

}  // namespace

extern "C" {

int dtx_version(void) { return 100; }

int dtx_create(int device, void* stream, dtx_ctx** out) {
    if (!out) return DTX_ERR_ARG;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return DTX_ERR_DEVICE;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return DTX_ERR_CUDA;
    if (prop.major != 10) return DTX_ERR_DEVICE;  // sm_100a only, no fallback
    if (cudaSetDevice(device) != cudaSuccess) return DTX_ERR_CUDA;
    dtx_ctx* ctx = new dtx_ctx();
    ctx->device = device;
    ctx->num_sms = prop.multiProcessorCount;
    if (stream) {
        ctx->stream = static_cast<cudaStream_t>(stream);
    } else {
        if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
            delete ctx;
            return DTX_ERR_CUDA;
        }
        ctx->own_stream = true;
    }
    *out = ctx;
    return DTX_OK;
}

void dtx_destroy(dtx_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (auto& ev : ctx->k1_events) {
        cudaEventDestroy(ev.first);
        cudaEventDestroy(ev.second);
    }
    if (ctx->copy_stream) {
        cudaStreamSynchronize(ctx->copy_stream);
        cudaStreamDestroy(ctx->copy_stream);
    }
    if (ctx->ev_raw_read) cudaEventDestroy(ctx->ev_raw_read);
    if (ctx->ev_copy_done) cudaEventDestroy(ctx->ev_copy_done);
    cudaStream_t st = ctx->own_stream ? ctx->stream : nullptr;
    delete ctx;   // every device / pinned buffer is released by its destructor
    if (st) cudaStreamDestroy(st);
}

const char* dtx_last_error(const dtx_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

static int ccx_check_flags(dtx_ctx* ctx);

int dtx_sync(dtx_ctx* ctx) {
    if (!ctx) return DTX_ERR_ARG;
    DTX_CUDA(cudaSetDevice(ctx->device));
    DTX_CUDA(cudaStreamSynchronize(ctx->stream));
    return ccx_check_flags(ctx);
}

// Basis set from vectors already in bs.d_U ([R][n] float64 on the device; `fill_U`, if given, is
// called once the buffer exists and puts them there).  All per-vector statistics come from one
// reduction kernel, so nothing here is O(R*n) on the host.
static int set_bases_device(dtx_ctx* ctx, BasisSet& bs, int set_id, const std::function<int(double*)>& fill_U,
                            const int32_t* rank_off, int S, int n, int Nc, const double* thresholds) {
    const int R = rank_off[S];
    for (int s = 0; s < S; ++s) {
        const int r = rank_off[s + 1] - rank_off[s];
        if (r < 1) return fail(ctx, DTX_ERR_ARG, "dtx_set_bases: subspace rank must be >= 1");
    }
    {   // every argument check comes before the stored vectors are touched
        const int kc = round_up(n / Nc + 7, CHUNK_TAPS);
        if (Nc * ((kc + MAX_SEG_TAPS - 1) / MAX_SEG_TAPS) > MAX_SEGS)
            return fail(ctx, DTX_ERR_ARG, "dtx_set_bases: template too long");
    }
    DTX_CUDA(bs.d_U.reserve(static_cast<size_t>(R) * n));
    {
        const int rc = fill_U(bs.d_U.p);
        if (rc != DTX_OK) return rc;
    }
    // per row: sum, max |u|, sum u^2, sum u^4
    DevBuf<double> d_stats;
    DTX_CUDA(d_stats.reserve(static_cast<size_t>(R) * 4));
    launch_basis_row_stats(bs.d_U.p, R, n, d_stats.p, ctx->stream);
    ctx->launches += 1;
    DTX_CUDA(cudaGetLastError());
    std::vector<double> stats(static_cast<size_t>(R) * 4);
    DTX_CUDA(cudaMemcpyAsync(stats.data(), d_stats.p, sizeof(double) * R * 4, cudaMemcpyDeviceToHost, ctx->stream));
    DTX_CUDA(cudaStreamSynchronize(ctx->stream));
    for (auto& kv : bs.ev_blob) kv.second.release();   // events belong to the previous bases
    bs.ev_blob.clear();
    bs.ev_meta.clear();
    if (ctx->run_set == set_id) ctx->ran = false;      // results of the old bases are stale
    BasisLayout& lay = bs.lay;
    lay = BasisLayout{};
    lay.Nc = Nc;
    lay.n = n;
    lay.ns = n / Nc;
    lay.S = S;
    bs.R = R;
    bs.rank_off.assign(rank_off, rank_off + S + 1);

    // K segments: per channel, taps 0 .. ns+7 (8 phases) rounded to 64, split at MAX_SEG_TAPS
    const int Kc = round_up(lay.ns + 7, CHUNK_TAPS);
    int nseg = 0, chunk0 = 0;
    for (int c = 0; c < Nc; ++c)
        for (int t0 = 0; t0 < Kc; t0 += MAX_SEG_TAPS) {
            if (nseg >= MAX_SEGS) return fail(ctx, DTX_ERR_ARG, "dtx_set_bases: template too long");
            Seg& sg = lay.seg[nseg++];
            sg.chan = c;
            sg.tap0 = t0;
            sg.ntaps = std::min(MAX_SEG_TAPS, Kc - t0);
            sg.chunk0 = chunk0;
            chunk0 += sg.ntaps / CHUNK_TAPS;
        }
    lay.nseg = nseg;
    lay.nchunks = chunk0;

    // pack subspaces into blocks of 16 vector slots, first-fit decreasing by rank.  A subspace
    // of rank > 16 is cut into pieces of <= 16 vectors; DS is additive over the pieces (sum of
    // squared projections): the first piece writes the subspace's DS row, every further piece a
    // scratch row behind the S subspace rows, and launch_sum_pieces adds them up in piece order
    // (bit-reproducible, unlike atomics in the epilogue).
    struct Piece { int s, row0, r, out_row; };
    std::vector<Piece> pieces;
    bs.has_split = false;
    bs.pieces.clear();
    int nscratch = 0;
    for (int s = 0; s < S; ++s) {
        const int r = rank_off[s + 1] - rank_off[s];
        for (int k0 = 0; k0 < r; k0 += VEC_PER_BLOCK) {
            const int out_row = k0 == 0 ? s : S + nscratch++;
            pieces.push_back(Piece{s, rank_off[s] + k0, std::min(VEC_PER_BLOCK, r - k0), out_row});
            if (k0 > 0) bs.pieces.push_back(PieceSum{s, out_row});
        }
        if (r > VEC_PER_BLOCK) bs.has_split = true;
    }
    bs.ds_rows = S + nscratch;
    std::vector<int> order(pieces.size());
    for (size_t i = 0; i < pieces.size(); ++i) order[i] = static_cast<int>(i);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return pieces[a].r > pieces[b].r; });
    std::vector<int> used;                    // slots used per block
    std::vector<std::vector<int>> members;    // pieces per block
    for (int pi : order) {
        const int r = pieces[pi].r;
        int b = -1;
        for (size_t i = 0; i < used.size(); ++i)
            if (used[i] + r <= VEC_PER_BLOCK) { b = static_cast<int>(i); break; }
        if (b < 0) { used.push_back(0); members.emplace_back(); b = static_cast<int>(used.size()) - 1; }
        used[b] += r;
        members[b].push_back(pi);
    }
    lay.nblocks = static_cast<int>(used.size());
    // one all-zero padding block behind an odd number of blocks: CTA pairs (cta_group::2) take blocks two at a time
    const int nblk_pad = lay.nblocks + (lay.nblocks & 1);
    std::vector<int> slot_row(static_cast<size_t>(nblk_pad) * VEC_PER_BLOCK, -1);
    std::vector<BlockInfo> binfo(static_cast<size_t>(nblk_pad) * VEC_PER_BLOCK);
    for (size_t i = static_cast<size_t>(lay.nblocks) * VEC_PER_BLOCK; i < binfo.size(); ++i) {
        binfo[i].sumU = 0.f; binfo[i].out_row = -1; binfo[i].nrows = 0; binfo[i].seg_end = static_cast<int>(i % VEC_PER_BLOCK) + 1;
    }
    double umax = 0.0;
    for (int k = 0; k < R; ++k) umax = std::max(umax, stats[static_cast<size_t>(k) * 4 + 1]);
    int eu = 0;
    if (umax > 0 && std::isfinite(umax)) eu = 13 - std::ilogb(umax);  // max|U| * 2^eu in [2^13, 2^14)
    lay.u_exp = eu;
    bs.nu4 = 0.0;
    for (int k = 0; k < R; ++k) {
        const double s2 = stats[static_cast<size_t>(k) * 4 + 2], s4 = stats[static_cast<size_t>(k) * 4 + 3];
        const double q = s2 > 0 ? s4 / (s2 * s2) : 1.0;
        bs.nu4 = std::max(bs.nu4, std::isfinite(q) ? q : 1.0);
    }
    lay.u_inv_scale = std::ldexp(1.0f, -eu);
    for (int b = 0; b < lay.nblocks; ++b) {
        int slot = 0;
        for (int i = 0; i < VEC_PER_BLOCK; ++i) {
            BlockInfo& bi = binfo[static_cast<size_t>(b) * VEC_PER_BLOCK + i];
            bi.sumU = 0.f; bi.out_row = -1; bi.nrows = 0; bi.seg_end = i + 1;
        }
        for (int pi : members[b]) {
            const Piece& pc = pieces[pi];
            const int r = pc.r, s = pc.s;
            for (int k = 0; k < r; ++k, ++slot) {
                const int row = pc.row0 + k;
                slot_row[static_cast<size_t>(b) * VEC_PER_BLOCK + slot] = row;
                BlockInfo& bi = binfo[static_cast<size_t>(b) * VEC_PER_BLOCK + slot];
                bi.sumU = static_cast<float>(stats[static_cast<size_t>(row) * 4]);   // float64 block reduction
                bi.out_row = pc.out_row;
                bi.nrows = (k == 0) ? r : 0;
                bi.seg_end = slot - k + r;
                (void)s;
            }
        }
    }
    DTX_CUDA(bs.d_rank_off.reserve(S + 1));
    DTX_CUDA(bs.d_slot_row.reserve(slot_row.size()));
    DTX_CUDA(bs.d_binfo.reserve(binfo.size()));
    DTX_CUDA(bs.d_thr.reserve(S));
    DTX_CUDA(bs.d_hist.reserve(static_cast<size_t>(S) * HIST_MAX_BINS));
    DTX_CUDA(bs.d_fas.reserve(static_cast<size_t>(S) * 5));
    const size_t img_bytes = static_cast<size_t>(nblk_pad) * lay.nchunks * 32768;
    DTX_CUDA(bs.d_Aimg.reserve(img_bytes));
    // synchronous copies: the caller's arrays need not outlive this call
    DTX_CUDA(cudaStreamSynchronize(ctx->stream));
    DTX_CUDA(cudaMemcpy(bs.d_rank_off.p, rank_off, sizeof(int) * (S + 1), cudaMemcpyHostToDevice));
    DTX_CUDA(cudaMemcpy(bs.d_slot_row.p, slot_row.data(), sizeof(int) * slot_row.size(), cudaMemcpyHostToDevice));
    DTX_CUDA(cudaMemcpy(bs.d_binfo.p, binfo.data(), sizeof(BlockInfo) * binfo.size(), cudaMemcpyHostToDevice));
    if (!bs.pieces.empty()) {
        DTX_CUDA(bs.d_pieces.reserve(bs.pieces.size()));
        DTX_CUDA(cudaMemcpy(bs.d_pieces.p, bs.pieces.data(), sizeof(PieceSum) * bs.pieces.size(), cudaMemcpyHostToDevice));
    }
    std::vector<float> thr(S, INFINITY);
    bs.has_thr = thresholds != nullptr;
    if (thresholds) for (int s = 0; s < S; ++s) thr[s] = static_cast<float>(thresholds[s]);
    DTX_CUDA(cudaMemcpy(bs.d_thr.p, thr.data(), sizeof(float) * S, cudaMemcpyHostToDevice));
    DTX_CUDA(cudaMemset(bs.d_hist.p, 0, sizeof(unsigned long long) * S * HIST_MAX_BINS));
    DTX_CUDA(cudaMemset(bs.d_fas.p, 0, sizeof(double) * S * 5));
    {
        BasisLayout lp = lay;
        lp.nblocks = nblk_pad;
        launch_basis_image(bs.d_U.p, bs.d_slot_row.p, lp, bs.d_Aimg.p, 0, ctx->stream);
    }
    bs.have_img8 = false;
    ctx->launches += 1;
    DTX_CUDA(cudaGetLastError());
    DTX_CUDA(cudaStreamSynchronize(ctx->stream));
    return DTX_OK;
}

int dtx_set_bases(dtx_ctx* ctx, int set_id, const double* U, const int32_t* rank_off, int S, int n,
                  int Nc, const double* thresholds) {
    if (!ctx) return DTX_ERR_ARG;
    if (!U || !rank_off || S < 1 || Nc < 1 || n < Nc || n % Nc != 0)
        return fail(ctx, DTX_ERR_ARG, "dtx_set_bases: bad shape (need S>=1, n % Nc == 0)");
    DTX_CUDA(cudaSetDevice(ctx->device));
    const size_t count = static_cast<size_t>(rank_off[S]) * n;
    auto fill = [&](double* d_U) -> int {
        // synchronous copy: the caller's array need not outlive this call
        DTX_CUDA(cudaStreamSynchronize(ctx->stream));
        DTX_CUDA(cudaMemcpy(d_U, U, sizeof(double) * count, cudaMemcpyHostToDevice));
        return DTX_OK;
    };
    return set_bases_device(ctx, ctx->sets[set_id], set_id, fill, rank_off, S, n, Nc, thresholds);
}

static int set_chunk_table(dtx_ctx* ctx, int nchunks, const int64_t* L, const int64_t* offs, int dtype) {
    if (nchunks < 1 || !L) return fail(ctx, DTX_ERR_ARG, "need nchunks >= 1");
    if (dtype != DTX_F64 && dtype != DTX_F32) return fail(ctx, DTX_ERR_ARG, "dtype must be DTX_F64 or DTX_F32");
    ctx->nchunks = nchunks;
    ctx->dtype = dtype;
    ctx->raw_off.resize(nchunks);
    ctx->rawL.resize(nchunks);
    ctx->core_lo.clear();
    ctx->core_hi.clear();
    long long off = 0;
    for (int i = 0; i < nchunks; ++i) {
        if (L[i] < 1 || L[i] > (1LL << 30)) return fail(ctx, DTX_ERR_ARG, "chunk length out of range");
        ctx->rawL[i] = L[i];
        ctx->raw_off[i] = offs ? offs[i] : off;
        off += (L[i] + 1) & ~1LL;  // keep chunks 16 B aligned for both dtypes
    }
    ctx->ran = false;
    return DTX_OK;
}

int dtx_load_chunks(dtx_ctx* ctx, int nchunks, const void* const* host_ptrs, const int64_t* L, int dtype) {
    if (!ctx || !host_ptrs) return DTX_ERR_ARG;
    DTX_CUDA(cudaSetDevice(ctx->device));
    int rc = set_chunk_table(ctx, nchunks, L, nullptr, dtype);
    if (rc) return rc;
    const size_t esz = dtype == DTX_F32 ? 4 : 8;
    const long long total = ctx->raw_off.back() + ((ctx->rawL.back() + 1) & ~1LL);
    if (!ctx->copy_stream) {
        DTX_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
        DTX_CUDA(cudaEventCreateWithFlags(&ctx->ev_copy_done, cudaEventDisableTiming));
    }
    // the host arrays of the previous dtx_load_chunks may be released by the caller once this call
    // has started: wait for their copy (long finished unless nothing ran in between)
    if (ctx->copy_pending) DTX_CUDA(cudaEventSynchronize(ctx->ev_copy_done));
    DTX_CUDA(ctx->raw_own.reserve(static_cast<size_t>(total) * esz));
    // the previous batch may still be in flight: overwrite its chunks only behind the last kernel that
    // reads them (K0 / the float64 engine / magnitudes), not behind the whole projection
    if (ctx->raw_read_pending) DTX_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_raw_read, 0));
    for (int i = 0; i < nchunks; ++i)
        DTX_CUDA(cudaMemcpyAsync(ctx->raw_own.p + ctx->raw_off[i] * esz, host_ptrs[i],
                                 static_cast<size_t>(L[i]) * esz, cudaMemcpyHostToDevice, ctx->copy_stream));
    DTX_CUDA(cudaEventRecord(ctx->ev_copy_done, ctx->copy_stream));
    ctx->copy_pending = true;
    DTX_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_copy_done, 0));
    ctx->d_raw = ctx->raw_own.p;
    return DTX_OK;
}

int dtx_set_core_lags(dtx_ctx* ctx, const int64_t* lo, const int64_t* hi) {
    if (!ctx) return DTX_ERR_ARG;
    if (ctx->nchunks < 1) return fail(ctx, DTX_ERR_STATE, "dtx_set_core_lags: no chunks loaded");
    if (!lo || !hi) {
        ctx->core_lo.clear();
        ctx->core_hi.clear();
        return DTX_OK;
    }
    ctx->core_lo.assign(lo, lo + ctx->nchunks);
    ctx->core_hi.assign(hi, hi + ctx->nchunks);
    return DTX_OK;
}

int dtx_attach_device_chunks(dtx_ctx* ctx, int nchunks, const void* dev_base, const int64_t* elem_offsets,
                             const int64_t* L, int dtype) {
    if (!ctx || !dev_base || !elem_offsets) return DTX_ERR_ARG;
    int rc = set_chunk_table(ctx, nchunks, L, elem_offsets, dtype);
    if (rc) return rc;
    ctx->d_raw = dev_base;
    return DTX_OK;
}

// what K1's fused epilogue (MODE 2) needs from dtx_detect_run
struct FusedArgs {
    double hist_lo, hist_hi;
    int want_fas, row_base;
};

// K0 + (K1 | fp64 direct) on the loaded chunks; leaves DS in ctx->d_DS (unless `fused`: then K1 itself
// produces rowmax / flags / histograms / candidates / FAS sums and no DS exists).
static int project_run(dtx_ctx* ctx, BasisSet& bs, int engine, int kblk, int mode, int keep_ds64,
                       const int* blk_hi, int hi_only = 0, const FusedArgs* fused = nullptr) {
    const BasisLayout& lay = bs.lay;
    const int Nc = lay.Nc, n = lay.n, ns = lay.ns, S = lay.S;
    const int Kc = round_up(ns + 7, CHUNK_TAPS);
    const int nchunks = ctx->nchunks;

    ctx->h_chunks.resize(nchunks);
    long long sig = 0, nrm = 0, ds = 0;
    int max_Lpad = 0, max_ntiles = 0, maxT = 0;
    long long nitems = 0;
    for (int i = 0; i < nchunks; ++i) {
        ChunkDesc& cd = ctx->h_chunks[i];
        const long long L = ctx->rawL[i] / Nc * Nc;
        if (L <= n) return fail(ctx, DTX_ERR_SHORT_CHUNK, "chunk not longer than the template (detect.py:262)");
        cd.raw_off = ctx->raw_off[i];
        cd.L = static_cast<int>(L);
        cd.Ls = static_cast<int>(L / Nc);
        cd.T = cd.Ls - ns + 1;
        if (cd.T < 10) return fail(ctx, DTX_ERR_SHORT_CHUNK, "fewer than 10 lags (detect.py:270)");
        cd.ntiles = (cd.T + TILE_T - 1) / TILE_T;
        cd.Tpad = cd.ntiles * TILE_T;
        cd.Lpad = round_up(cd.Tpad + Kc, 64);
        cd.sig_off = sig; cd.norm_off = nrm; cd.ds_off = ds;
        cd.blk_lo = 0;
        cd.blk_hi = blk_hi ? blk_hi[i] : lay.nblocks;
        cd.t_lo = 0;
        cd.t_hi = cd.T;
        if (mode == 0 && !ctx->core_lo.empty()) {
            if (ctx->core_lo[i] < 0 || ctx->core_lo[i] >= ctx->core_hi[i] || ctx->core_hi[i] > cd.T || ctx->core_lo[i] % 4)
                return fail(ctx, DTX_ERR_ARG, "dtx_set_core_lags: need 0 <= lo < hi <= T and lo % 4 == 0 for every chunk");
            cd.t_lo = static_cast<int>(ctx->core_lo[i]);
            cd.t_hi = static_cast<int>(ctx->core_hi[i]);
        }
        sig += 2LL * Nc * cd.Lpad;
        nrm += cd.Tpad;
        ds += static_cast<long long>(bs.ds_rows) * cd.Tpad;
        max_Lpad = std::max(max_Lpad, cd.Lpad);
        max_ntiles = std::max(max_ntiles, cd.ntiles);
        maxT = std::max(maxT, cd.T);
        nitems += cd.ntiles;
    }
    // K1 tile: 2048 lags (N = 256), or 1024 (N = 128) when no chunk has more than 1024 lags
    const int nq = maxT <= TILE_T / 2 ? 128 : 256;
    // Work items (chunk, tile, basis block).  Two orders, both keep the 148 CTAs on the same basis
    // block at any time so that its image (lay.nchunks x 32 KB) is streamed from L2:
    //   super = 1: (chunk group, block, tile)  -- a group's split signal + norm tiles are re-read once
    //              per block pass and should survive it in L2; needs >= 4 waves of items per pass
    //              (with 2.4 waves the CTAs spread over several blocks: 67 GB of DRAM reads);
    //   super = S: (chunk group, superblock of S blocks, wave of num_sms tiles, block, tile) -- a tile's
    //              signal is re-read S times in a row and the group's signal only once per SUPERBLOCK
    //              pass, so the group can be the whole batch and the image is read from HBM once.
    // Pure reordering of the list; which one reads less HBM depends on the shape.  Cost model fitted
    // to the ncu sweep in profiles/r01_k1_traffic_ab.md (48 x 256-subspace detection chunks: group 4 /
    // super 1 = 11.2 GB, whole batch / super 4 = 7.1 GB, super 8 thrashes the image; CCX, with
    // 0.5 MB blocks and a 15 MB group signal, is best left at super 1):
    //   image reads  = image bytes x number of groups
    //   signal reads = signal bytes x (1 + miss x (passes - 1)),  passes = nblocks / super,
    //   miss         = clamp((super x group signal + 2 super x block - 15 MB) / 120 MB, 0, 1)
    const int tiles_per_chunk = std::max(1, (maxT + 8 * nq - 1) / (8 * nq));
    const int g4 = std::max(1, (4 * ctx->num_sms + tiles_per_chunk - 1) / tiles_per_chunk);
    const double blk_bytes = 32768.0 * lay.nchunks, img_bytes = blk_bytes * lay.nblocks;
    const double sig_bytes = 2.0 * sig + 8.0 * nrm;   // fp16 planes + mu / invE
    auto dram_estimate = [&](int G, int S) {
        const int ngroups = (nchunks + G - 1) / G;
        const double miss = std::min(1.0, std::max(0.0, (S * sig_bytes / ngroups + 2.0 * S * blk_bytes - 15e6) / 120e6));
        return img_bytes * ngroups + sig_bytes * (1.0 + miss * (static_cast<double>(lay.nblocks) / S - 1.0));
    };
    int group = std::min(g4, nchunks), super = 1;
    {
        double best = dram_estimate(group, 1);
        const int cand_g[3] = {std::min(g4, nchunks), std::min(2 * g4, nchunks), nchunks};
        for (int S : {1, 4}) {
            if (S > 1 && (2.0 * S * blk_bytes > 40e6 || lay.nblocks < 2 * S)) continue;
            for (int G : cand_g) {
                if (S == 1 && G < std::min(g4, nchunks)) continue;
                const double e = dram_estimate(G, S);
                if (e < 0.8 * best) { best = e; group = G; super = S; }
            }
        }
    }
    if (const char* g = std::getenv("DTX_K1_GROUP"))   // experiment knobs: force the order
        if (std::atoi(g) > 0) group = std::atoi(g);
    if (const char* g = std::getenv("DTX_K1_SUPER"))
        if (std::atoi(g) > 0) super = std::atoi(g);
    // CTA pairs (cta_group::2, DTX_K1_CG2=1): detection only, 2048-lag tiles; one item per cluster and block PAIR
    int cg2 = 0;
    if (const char* g = std::getenv("DTX_K1_CG2")) cg2 = std::atoi(g) != 0;
    cg2 = cg2 && mode != 1 && nq == 256 && engine != DTX_ENGINE_FP64 && ctx->num_sms >= 2;
    const int bstep = cg2 ? 2 : 1;
    if (cg2) super = std::max(2, super & ~1);
    // CCX screening series (mode 1, one MMA per K step, tiles of 1024 lags): chunks are taken in PAIRS that share
    // the stream of the basis image (k1_project.cu, DUAL); DTX_K1_NODUAL=1 keeps one chunk per item
    const bool dual = mode == 1 && hi_only && nq == 128 && !cg2 && engine != DTX_ENGINE_FP64 &&
                      !(std::getenv("DTX_K1_NODUAL") && std::atoi(std::getenv("DTX_K1_NODUAL")) != 0);
    const int wave = cg2 ? ctx->num_sms / 2 : ctx->num_sms;
    // the list only depends on the batch's shape: reuse the device copy when it has not changed
    std::vector<int> sig_key{nq, lay.nblocks, group, nchunks, super, wave, cg2, dual ? 1 : 0};
    for (int i = 0; i < nchunks; ++i) {
        sig_key.push_back(ctx->h_chunks[i].T);
        sig_key.push_back(ctx->h_chunks[i].blk_hi);
    }
    const bool items_cached = sig_key == ctx->items_key && ctx->d_items.p != nullptr;
    std::vector<int4> items;
    if (!items_cached) items.reserve(static_cast<size_t>(nitems) * 2 * lay.nblocks);
    struct TileRef { int chunk, tile, partner, blk_lo, blk_hi; };
    std::vector<TileRef> tl;   // (chunk, tile) of the current group; with `dual`, (chunk pair, tile)
    for (int g0 = 0; g0 < nchunks && !items_cached; g0 += group) {
        const int g1 = std::min(nchunks, g0 + group);
        int max_blk = 0;
        tl.clear();
        for (int i = g0; i < g1; i += dual ? 2 : 1) {
            const int partner = (dual && i + 1 < g1) ? i + 1 : i;   // an odd last chunk pairs with itself
            const ChunkDesc &c0 = ctx->h_chunks[i], &c1 = ctx->h_chunks[partner];
            const int lo = std::min(c0.blk_lo, c1.blk_lo), hi = std::max(c0.blk_hi, c1.blk_hi);
            max_blk = std::max(max_blk, hi);
            const int nt = (std::max(c0.T, c1.T) + 8 * nq - 1) / (8 * nq);
            for (int t = 0; t < nt; ++t) tl.push_back(TileRef{i, t, dual ? partner : 0, lo, hi});
        }
        for (int sb = 0; sb < max_blk; sb += super)
            for (size_t w0 = 0; w0 < tl.size(); w0 += (super > 1 ? wave : tl.size()))
                for (int b = sb; b < std::min(sb + super, max_blk); b += bstep) {
                    const size_t w1 = super > 1 ? std::min(tl.size(), w0 + wave) : tl.size();
                    for (size_t k = w0; k < w1; ++k) {
                        if (b < tl[k].blk_lo || b >= tl[k].blk_hi) continue;
                        items.push_back(make_int4(tl[k].chunk, tl[k].tile, b, tl[k].partner));
                    }
                }
    }

    DTX_CUDA(ctx->d_chunks.reserve(nchunks));
    if (!items_cached) {
        DTX_CUDA(ctx->d_items.reserve(items.size()));
        ctx->items_key = sig_key;
        ctx->n_items = static_cast<int>(items.size());
    }
    DTX_CUDA(ctx->d_xsplit.reserve(sig));
    DTX_CUDA(ctx->d_mu.reserve(nrm));
    DTX_CUDA(ctx->d_invE.reserve(nrm));
    if (!fused) DTX_CUDA(ctx->d_DS.reserve(ds));
    if (keep_ds64) DTX_CUDA(ctx->d_DS64.reserve(ds));
    DTX_CUDA(ctx->d_chunk_bad.reserve(nchunks));
    DTX_CUDA(ctx->d_scale.reserve(nchunks));
    DTX_CUDA(ctx->d_sum.reserve(nchunks));
    DTX_CUDA(ctx->d_maxbits.reserve(nchunks));
    DTX_CUDA(ctx->d_k4bits.reserve(nchunks));
    DTX_CUDA(ctx->d_chunk_mode.reserve(nchunks));
    DTX_CUDA(ctx->d_zeroE.reserve(nchunks));
    if (!ctx->accumulate) {
        DTX_CUDA(ctx->d_rowmax.reserve(static_cast<size_t>(nchunks) * S));
        DTX_CUDA(ctx->d_rowflags.reserve(static_cast<size_t>(nchunks) * S));
    }
    DTX_CUDA(ctx->d_ncand.reserve(2));   // [0] = candidates so far, [1] = count before the current batch
    DTX_CUDA(ctx->d_cand.reserve(ctx->cand_cap));
    cudaStream_t st = ctx->stream;
    DTX_CUDA(cudaMemcpyAsync(ctx->d_chunks.p, ctx->h_chunks.data(), sizeof(ChunkDesc) * nchunks,
                             cudaMemcpyHostToDevice, st));
    if (!items_cached)
        DTX_CUDA(cudaMemcpyAsync(ctx->d_items.p, items.data(), sizeof(int4) * items.size(),
                                 cudaMemcpyHostToDevice, st));
    if (!ctx->accumulate) DTX_CUDA(cudaMemsetAsync(ctx->d_ncand.p, 0, 2 * sizeof(int), st));
    else DTX_CUDA(cudaMemcpyAsync(ctx->d_ncand.p + 1, ctx->d_ncand.p, sizeof(int), cudaMemcpyDeviceToDevice, st));

    const int f32 = ctx->dtype == DTX_F32;
    // 8-bit cross terms: forced, or per chunk where the random-rounding model
    //   rms error of (u . w)/(|u||w|)  <=  X8_C * (K4 * nu4)^(1/4)
    // (Cauchy-Schwarz on sum u^2 x^2; K4, nu4 = fourth-moment concentration of the worst window
    // and of the worst basis vector) stays below ctx->x8_eps.  X8_C = rms relative error of the two
    // 8-bit cross products per tap, relative to |u_j x_j|.
    constexpr double X8_C = 1.6e-5;
    const int x8 = (engine == DTX_ENGINE_TCGEN05_X8) ? X8_FORCE
                 : (engine == DTX_ENGINE_TCGEN05_AUTO && mode == 0) ? X8_AUTO : X8_OFF;
    // CCX (mode 1): k0_norm records, per padded event, the worst ratio of window power to window energy
    // -- the amplification of operand rounding in the normalised series (band of ccx_scan_kernel)
    const int k0_policy = mode == 1 ? X8_RATIO : x8;
    const float k4_limit = static_cast<float>(std::pow(ctx->x8_eps / X8_C, 4.0) / bs.nu4);
    if (x8 && !bs.have_img8) {
        BasisLayout lp = lay;
        lp.nblocks = lay.nblocks + (lay.nblocks & 1);
        DTX_CUDA(bs.d_Aimg8.reserve(static_cast<size_t>(lp.nblocks) * lay.nchunks * 32768));
        launch_basis_image(bs.d_U.p, bs.d_slot_row.p, lp, bs.d_Aimg8.p, 1, st);
        bs.have_img8 = true;
        ctx->launches += 1;
    }
    launch_k0(ctx->d_raw, f32, ctx->d_chunks.p, nchunks, Nc, n, max_Lpad, max_ntiles, ctx->d_sum.p,
              ctx->d_maxbits.p, ctx->d_scale.p, ctx->d_xsplit.p, ctx->d_mu.p, ctx->d_invE.p, k0_policy, k4_limit,
              ctx->d_k4bits.p, ctx->d_chunk_mode.p, mode == 0 ? ctx->d_zeroE.p : nullptr, ctx->d_chunk_bad.p, st);
    DTX_CUDA(cudaGetLastError());
    if (engine != DTX_ENGINE_FP64 && !keep_ds64) mark_raw_access(ctx);   // K1 works on the split planes only
    ctx->launches += 3;  // k0_stats, k0_split, k0_norm
    ctx->have_ds64 = false;
    ctx->k1_timed = false;
    if (engine != DTX_ENGINE_FP64) {
        K1Args a;
        a.Aimg = bs.d_Aimg.p; a.Aimg8 = x8 ? bs.d_Aimg8.p : nullptr;
        a.chunk_mode = x8 ? ctx->d_chunk_mode.p : nullptr;
        a.kblk8 = (3 * kblk + 1) / 2;   // same number of MMAs per TMEM accumulation as the 3-MMA mode
        a.xsplit = ctx->d_xsplit.p; a.mu = ctx->d_mu.p; a.invE = ctx->d_invE.p;
        a.chunk_scale = ctx->d_scale.p; a.chunks = ctx->d_chunks.p; a.items = ctx->d_items.p;
        a.binfo = bs.d_binfo.p; a.DS = ctx->d_DS.p; a.nitems = ctx->n_items;
        a.kblk = kblk; a.num_sms = ctx->num_sms; a.nq = dual ? 256 : nq; a.mode = mode;
        a.hi_only = hi_only;
        a.dual = dual ? 1 : 0;
        a.cg2 = cg2;
        a.fused = fused ? 1 : 0;
        a.thr = bs.d_thr.p; a.rowmax_bits = reinterpret_cast<unsigned*>(ctx->d_rowmax.p); a.rowflags = ctx->d_rowflags.p;
        a.hist = bs.d_hist.p; a.nbins = ctx->hist_bins; a.cand = ctx->d_cand.p; a.cand_cap = ctx->cand_cap;
        a.ncand = ctx->d_ncand.p; a.chunk_bad = ctx->d_chunk_bad.p; a.S = S;
        a.hist_lo = 0.0; a.hist_hi = 1.0; a.fas = nullptr; a.row_base = 0;
        if (fused) {
            a.DS = nullptr;
            a.hist_lo = fused->hist_lo; a.hist_hi = fused->hist_hi; a.row_base = fused->row_base;
            a.fas = fused->want_fas ? bs.d_fas.p : nullptr;
            // the epilogue raises maxima / flags with atomics: start the rows of this batch from zero
            const size_t rows = static_cast<size_t>(nchunks) * S;
            DTX_CUDA(cudaMemsetAsync(ctx->d_rowmax.p + fused->row_base, 0, sizeof(float) * rows, st));
            DTX_CUDA(cudaMemsetAsync(ctx->d_rowflags.p + fused->row_base, 0, sizeof(int) * rows, st));
        }
        if (!ctx->accumulate && mode == 0) ctx->k1_used = 0;   // only the last run's pair is kept (CCX keeps its batches')
        if (ctx->k1_used >= ctx->k1_events.size()) {
            cudaEvent_t e0 = nullptr, e1 = nullptr;
            DTX_CUDA(cudaEventCreate(&e0));
            DTX_CUDA(cudaEventCreate(&e1));
            ctx->k1_events.emplace_back(e0, e1);
        }
        DTX_CUDA(cudaEventRecord(ctx->k1_events[ctx->k1_used].first, st));
        launch_k1(a, lay, st);
        DTX_CUDA(cudaEventRecord(ctx->k1_events[ctx->k1_used].second, st));
        ctx->k1_used += 1;
        ctx->k1_timed = true;
        ctx->launches += 1 + (keep_ds64 ? 1 : 0);
        DTX_CUDA(cudaGetLastError());
        if (bs.has_split) {
            int max_Tpad = 0;
            for (int i = 0; i < nchunks; ++i) max_Tpad = std::max(max_Tpad, ctx->h_chunks[i].Tpad);
            launch_sum_pieces(ctx->d_chunks.p, nchunks, max_Tpad, bs.d_pieces.p, static_cast<int>(bs.pieces.size()),
                              ctx->d_DS.p, st);
            ctx->launches += 1;
            DTX_CUDA(cudaGetLastError());
        }
        if (keep_ds64) {
            launch_direct(ctx->d_raw, f32, ctx->d_chunks.p, nchunks, bs.d_U.p, bs.d_rank_off.p, S, n, Nc,
                          maxT, ctx->d_sum.p, nullptr, ctx->d_DS64.p, st);
            ctx->have_ds64 = true;
        }
    } else {
        launch_direct(ctx->d_raw, f32, ctx->d_chunks.p, nchunks, bs.d_U.p, bs.d_rank_off.p, S, n, Nc, maxT,
                      ctx->d_sum.p, ctx->d_DS.p, keep_ds64 ? ctx->d_DS64.p : nullptr, st);
        ctx->have_ds64 = keep_ds64 != 0;
        ctx->launches += 1;
    }
    DTX_CUDA(cudaGetLastError());
    if (engine == DTX_ENGINE_FP64 || keep_ds64) mark_raw_access(ctx);    // the float64 kernel reads the raw chunks
    if (fused) {       // rows of chunks with non-finite samples: MaxDS = NaN, flag bit 0
        launch_fused_bad_rows(ctx->d_chunk_bad.p, nchunks, S, fused->row_base, reinterpret_cast<unsigned*>(ctx->d_rowmax.p),
                              ctx->d_rowflags.p, st);
        ctx->launches += 1;
        DTX_CUDA(cudaGetLastError());
    } else if (mode == 0) {   // zero-energy windows: DS = +inf in every subspace row (reference: x/0, detect.py:577)
        launch_zero_energy_fix(ctx->d_chunks.p, nchunks, max_ntiles, S, ctx->d_invE.p, ctx->d_zeroE.p, ctx->d_DS.p,
                               ctx->have_ds64 ? ctx->d_DS64.p : nullptr, st);
        ctx->launches += 1;
        DTX_CUDA(cudaGetLastError());
    }
    ctx->run_S = bs.lay.S;
    return DTX_OK;
}

int dtx_detect_run(dtx_ctx* ctx, int set_id, int engine, int kblk, double hist_lo, double hist_hi,
                   int lta_window, int want_fas, int keep_ds64) {
    if (!ctx) return DTX_ERR_ARG;
    DTX_CUDA(cudaSetDevice(ctx->device));
    auto it = ctx->sets.find(set_id);
    if (it == ctx->sets.end()) return fail(ctx, DTX_ERR_STATE, "dtx_detect_run: unknown basis set");
    if (ctx->nchunks < 1 || !ctx->d_raw) return fail(ctx, DTX_ERR_STATE, "dtx_detect_run: no chunks loaded");
    if (engine != DTX_ENGINE_TCGEN05 && engine != DTX_ENGINE_FP64 && engine != DTX_ENGINE_TCGEN05_X8 &&
        engine != DTX_ENGINE_TCGEN05_AUTO)
        return fail(ctx, DTX_ERR_ARG, "bad engine");
    // default: drain every 192 taps (36 MMAs per TMEM accumulation).  Same-box A/B (profiles/r02_kblk_ab.md):
    // kblk 2 / 3 / 4 = 5.995e9 / 6.118e9 / 6.151e9 ts/s with max |DS - float64| 1.5e-6 / 2.4e-6 / 3.4e-6 on
    // a planted chunk (DS = 0.97); 3 keeps a 4x margin to the 1e-5 tolerance.
    if (kblk == 0) kblk = 3;
    if (kblk < 1 || kblk > 64) return fail(ctx, DTX_ERR_ARG, "kblk out of range");
    BasisSet& bs = it->second;
    const int S = bs.lay.S;
    if (ctx->accumulate) {
        if (ctx->acc_chunks + ctx->nchunks > ctx->acc_capacity)
            return fail(ctx, DTX_ERR_CAPACITY, "dtx_detect_run: more chunks than dtx_accumulate_begin announced");
        if (ctx->acc_chunks > 0 && (ctx->run_set != set_id || ctx->run_S != S))
            return fail(ctx, DTX_ERR_STATE, "dtx_detect_run: the basis set changed inside an accumulation");
        DTX_CUDA(ctx->d_rowmax.reserve(static_cast<size_t>(ctx->acc_capacity) * S));   // no-ops once sized
        DTX_CUDA(ctx->d_rowflags.reserve(static_cast<size_t>(ctx->acc_capacity) * S));
    }
    const int row_base = ctx->accumulate ? static_cast<int>(ctx->acc_chunks * S) : 0;
    // fused mode: tensor-core engines, subspaces of rank <= 16 (the pieces of larger ones must be summed
    // before anything can be counted), no float64 copy requested
    const bool fused = ctx->fused && engine != DTX_ENGINE_FP64 && !keep_ds64 && !bs.has_split;
    FusedArgs fa{hist_lo, hist_hi, want_fas, row_base};
    if (fused && !ctx->accumulate) {
        DTX_CUDA(ctx->d_rowmax.reserve(static_cast<size_t>(ctx->nchunks) * S));
        DTX_CUDA(ctx->d_rowflags.reserve(static_cast<size_t>(ctx->nchunks) * S));
    }
    const int rc = project_run(ctx, bs, engine, kblk, 0, keep_ds64, nullptr, 0, fused ? &fa : nullptr);
    if (rc != DTX_OK) return rc;
    const int nchunks = ctx->nchunks;
    cudaStream_t st = ctx->stream;
    ctx->last_fused = fused;
    if (fused) {
        if (bs.has_thr && lta_window > 0) {
            // DS_STALTA of the candidates: the statistic around each from the float64 closed form
            if (lta_window > 65536) return fail(ctx, DTX_ERR_ARG, "dtx_detect_run: LTA window too long for the fused mode");
            DTX_CUDA(ctx->d_lta_acc.reserve(2 * static_cast<size_t>(ctx->cand_cap)));
            if (!ctx->d_lta_acc_zeroed) {
                DTX_CUDA(cudaMemsetAsync(ctx->d_lta_acc.p, 0, sizeof(double) * 2 * ctx->cand_cap, st));
                ctx->d_lta_acc_zeroed = true;
            }
            launch_lta_direct(ctx->d_raw, ctx->dtype == DTX_F32, ctx->d_chunks.p, ctx->d_sum.p, bs.d_U.p, bs.d_rank_off.p,
                              bs.lay.n, bs.lay.Nc, S, ctx->d_cand.p, ctx->d_ncand.p, ctx->d_ncand.p + 1, ctx->cand_cap,
                              row_base, lta_window, ctx->sta_window, ctx->d_lta_acc.p, st);
            DTX_CUDA(cudaGetLastError());
            mark_raw_access(ctx);       // reads the raw chunks
            ctx->launches += 2;
        }
        if (ctx->accumulate) ctx->acc_chunks += nchunks;
        ctx->run_set = set_id;
        ctx->run_S = S;
        ctx->ran = true;
        return DTX_OK;
    }
    launch_k3(ctx->d_DS.p, ctx->d_chunks.p, nchunks, S, bs.d_thr.p, ctx->d_rowmax.p, ctx->d_rowflags.p,
              bs.d_hist.p, hist_lo, hist_hi, ctx->hist_bins, ctx->d_cand.p, ctx->cand_cap, ctx->d_ncand.p,
              want_fas ? bs.d_fas.p : nullptr, row_base, st);
    DTX_CUDA(cudaGetLastError());
    ctx->launches += 2;  // k3_fast_kernel + k3_kernel (flagged rows only)
    if (bs.has_thr && lta_window > 0) {
        launch_lta(ctx->d_DS.p, ctx->d_chunks.p, S, ctx->d_rowflags.p, ctx->d_cand.p, ctx->d_ncand.p,
                   ctx->d_ncand.p + 1, ctx->cand_cap, row_base, lta_window, ctx->sta_window, st);
        DTX_CUDA(cudaGetLastError());
        ctx->launches += 1;
    }
    if (ctx->accumulate) ctx->acc_chunks += nchunks;
    ctx->run_set = set_id;
    ctx->run_S = S;
    ctx->ran = true;
    return DTX_OK;
}

/* Results of several dtx_detect_run calls (the batches of one station) accumulate on the device:
 * candidate rows and rowmax / rowflags entries are numbered by the chunk's position in the whole
 * sequence, and nothing has to be fetched (no host synchronisation) between the batches. */
int dtx_accumulate_begin(dtx_ctx* ctx, int64_t total_chunks) {
    if (!ctx) return DTX_ERR_ARG;
    if (total_chunks < 1 || total_chunks > (1LL << 24)) return fail(ctx, DTX_ERR_ARG, "dtx_accumulate_begin: bad chunk count");
    DTX_CUDA(cudaSetDevice(ctx->device));
    DTX_CUDA(ctx->d_ncand.reserve(2));
    DTX_CUDA(cudaMemsetAsync(ctx->d_ncand.p, 0, 2 * sizeof(int), ctx->stream));
    ctx->accumulate = true;
    ctx->acc_chunks = 0;
    ctx->acc_capacity = total_chunks;
    ctx->k1_used = 0;
    ctx->ran = false;
    return DTX_OK;
}

int dtx_accumulate_end(dtx_ctx* ctx) {
    if (!ctx) return DTX_ERR_ARG;
    ctx->accumulate = false;
    ctx->acc_chunks = 0;
    ctx->acc_capacity = 0;
    ctx->ran = false;
    return DTX_OK;
}

int dtx_set_fused(dtx_ctx* ctx, int on) {
    if (!ctx) return DTX_ERR_ARG;
    ctx->fused = on != 0;
    return DTX_OK;
}

int dtx_set_x8_tolerance(dtx_ctx* ctx, double eps) {
    if (!ctx) return DTX_ERR_ARG;
    if (!(eps >= 0.0) || !std::isfinite(eps)) return fail(ctx, DTX_ERR_ARG, "dtx_set_x8_tolerance: eps must be >= 0");
    ctx->x8_eps = eps;
    return DTX_OK;
}

int dtx_get_chunk_modes(dtx_ctx* ctx, int32_t* modes) {
    if (!ctx || !modes) return DTX_ERR_ARG;
    if (!ctx->ran) return fail(ctx, DTX_ERR_STATE, "dtx_get_chunk_modes: no run");
    DTX_CUDA(cudaSetDevice(ctx->device));
    DTX_CUDA(cudaMemcpyAsync(modes, ctx->d_chunk_mode.p, sizeof(int32_t) * ctx->nchunks,
                             cudaMemcpyDeviceToHost, ctx->stream));
    DTX_CUDA(cudaStreamSynchronize(ctx->stream));
    return DTX_OK;
}

int dtx_num_lags(dtx_ctx* ctx, int chunk, int64_t* T) {
    if (!ctx || !T) return DTX_ERR_ARG;
    if (!ctx->ran || chunk < 0 || chunk >= ctx->nchunks) return fail(ctx, DTX_ERR_STATE, "no run / bad chunk");
    *T = ctx->h_chunks[chunk].T;
    return DTX_OK;
}

int dtx_get_ds(dtx_ctx* ctx, int chunk, int subspace, float* out, int64_t count) {
    if (!ctx || !out) return DTX_ERR_ARG;
    if (!ctx->ran || chunk < 0 || chunk >= ctx->nchunks || subspace < 0 || subspace >= ctx->run_S)
        return fail(ctx, DTX_ERR_STATE, "dtx_get_ds: no run / bad index");
    if (ctx->last_fused) return fail(ctx, DTX_ERR_STATE, "dtx_get_ds: the last run was fused (dtx_set_fused): no dense DS exists");
    const ChunkDesc& cd = ctx->h_chunks[chunk];
    if (count < cd.T) return fail(ctx, DTX_ERR_CAPACITY, "dtx_get_ds: buffer smaller than T");
    DTX_CUDA(cudaSetDevice(ctx->device));
    DTX_CUDA(cudaMemcpyAsync(out, ctx->d_DS.p + cd.ds_off + static_cast<long long>(subspace) * cd.Tpad,
                             sizeof(float) * cd.T, cudaMemcpyDeviceToHost, ctx->stream));
    DTX_CUDA(cudaStreamSynchronize(ctx->stream));
    return DTX_OK;
}

int dtx_get_ds64(dtx_ctx* ctx, int chunk, int subspace, double* out, int64_t count) {
    if (!ctx || !out) return DTX_ERR_ARG;
    if (!ctx->ran || !ctx->have_ds64 || chunk < 0 || chunk >= ctx->nchunks || subspace < 0 ||
        subspace >= ctx->run_S)
        return fail(ctx, DTX_ERR_STATE, "dtx_get_ds64: run with keep_ds64 first");
    const ChunkDesc& cd = ctx->h_chunks[chunk];
    if (count < cd.T) return fail(ctx, DTX_ERR_CAPACITY, "dtx_get_ds64: buffer smaller than T");
    DTX_CUDA(cudaSetDevice(ctx->device));
    DTX_CUDA(cudaMemcpyAsync(out, ctx->d_DS64.p + cd.ds_off + static_cast<long long>(subspace) * cd.Tpad,
                             sizeof(double) * cd.T, cudaMemcpyDeviceToHost, ctx->stream));
    DTX_CUDA(cudaStreamSynchronize(ctx->stream));
    return DTX_OK;
}

int dtx_set_trigger_sta(dtx_ctx* ctx, int sta_window) {
    if (!ctx) return DTX_ERR_ARG;
    if (sta_window < 0) return fail(ctx, DTX_ERR_ARG, "dtx_set_trigger_sta: window must be >= 0");
    ctx->sta_window = sta_window;
    return DTX_OK;
}

int dtx_get_stalta(dtx_ctx* ctx, int chunk, int subspace, int W, float* out, int64_t count) {
    if (!ctx || !out) return DTX_ERR_ARG;
    if (!ctx->ran || chunk < 0 || chunk >= ctx->nchunks || subspace < 0 || subspace >= ctx->run_S || W < 1)
        return fail(ctx, DTX_ERR_STATE, "dtx_get_stalta: no run / bad index");
    if (ctx->last_fused) return fail(ctx, DTX_ERR_STATE, "dtx_get_stalta: the last run was fused (dtx_set_fused): no dense DS exists");
    const ChunkDesc& cd = ctx->h_chunks[chunk];
    if (count < cd.T) return fail(ctx, DTX_ERR_CAPACITY, "dtx_get_stalta: buffer smaller than T");
    DTX_CUDA(cudaSetDevice(ctx->device));
    const int Wsta = ctx->sta_window;
    if (W > stalta_dense_max_window() || Wsta > stalta_dense_max_window())
        return fail(ctx, DTX_ERR_ARG, "dtx_get_stalta: rolling window does not fit in shared memory (max ~13800 samples)");
    if (cd.T < W || cd.T < Wsta) {
        for (int i = 0; i < cd.T; ++i) out[i] = NAN;
        return DTX_OK;
    }
    int flags = 0;
    const long long batch0 = ctx->accumulate ? ctx->acc_chunks - ctx->nchunks : 0;   // first chunk of the last batch
    DTX_CUDA(cudaMemcpyAsync(&flags, ctx->d_rowflags.p + (batch0 + chunk) * ctx->run_S + subspace,
                             sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    DTX_CUDA(cudaStreamSynchronize(ctx->stream));
    DevBuf<float> tmp;
    DTX_CUDA(tmp.reserve(static_cast<size_t>(cd.T) * (Wsta > 0 ? 2 : 1)));
    launch_stalta_dense(ctx->d_DS.p + cd.ds_off + static_cast<long long>(subspace) * cd.Tpad, cd.T, W, Wsta,
                        (flags & 2) ? 1 : 0, tmp.p, tmp.p + cd.T, ctx->stream);
    DTX_CUDA(cudaGetLastError());
    DTX_CUDA(cudaMemcpyAsync(out, tmp.p, sizeof(float) * cd.T, cudaMemcpyDeviceToHost, ctx->stream));
    DTX_CUDA(cudaStreamSynchronize(ctx->stream));
    return DTX_OK;
}

int dtx_get_rowstats(dtx_ctx* ctx, float* maxds, int32_t* flags, int64_t count) {
    if (!ctx) return DTX_ERR_ARG;
    if (!ctx->ran) return fail(ctx, DTX_ERR_STATE, "dtx_get_rowstats: no run");
    const int64_t rows = (ctx->accumulate ? ctx->acc_chunks : static_cast<int64_t>(ctx->nchunks)) * ctx->run_S;
    if (count < rows) return fail(ctx, DTX_ERR_CAPACITY, "dtx_get_rowstats: buffer too small");
    DTX_CUDA(cudaSetDevice(ctx->device));
    if (maxds) DTX_CUDA(cudaMemcpyAsync(maxds, ctx->d_rowmax.p, sizeof(float) * rows, cudaMemcpyDeviceToHost, ctx->stream));
    if (flags) DTX_CUDA(cudaMemcpyAsync(flags, ctx->d_rowflags.p, sizeof(int) * rows, cudaMemcpyDeviceToHost, ctx->stream));
    DTX_CUDA(cudaStreamSynchronize(ctx->stream));
    return DTX_OK;
}

int dtx_set_hist_bins(dtx_ctx* ctx, int nbins) {
    if (!ctx) return DTX_ERR_ARG;
    if (nbins < 1 || nbins > HIST_MAX_BINS) return fail(ctx, DTX_ERR_ARG, "dtx_set_hist_bins: 1 <= nbins <= 1024");
    ctx->hist_bins = nbins;
    return DTX_OK;
}

int dtx_get_hist(dtx_ctx* ctx, int set_id, uint64_t* hist, int64_t count, int reset) {
    if (!ctx) return DTX_ERR_ARG;
    auto it = ctx->sets.find(set_id);
    if (it == ctx->sets.end()) return fail(ctx, DTX_ERR_STATE, "dtx_get_hist: unknown basis set");
    BasisSet& bs = it->second;
    const int nb = ctx->hist_bins;
    const int64_t nel = static_cast<int64_t>(bs.lay.S) * nb;
    if (hist && count < nel) return fail(ctx, DTX_ERR_CAPACITY, "dtx_get_hist: buffer too small");
    DTX_CUDA(cudaSetDevice(ctx->device));
    // device rows have a pitch of HIST_MAX_BINS; the caller gets [S][nbins] packed
    if (hist)
        DTX_CUDA(cudaMemcpy2DAsync(hist, sizeof(uint64_t) * nb, bs.d_hist.p, sizeof(uint64_t) * HIST_MAX_BINS,
                                   sizeof(uint64_t) * nb, bs.lay.S, cudaMemcpyDeviceToHost, ctx->stream));
    if (reset)
        DTX_CUDA(cudaMemsetAsync(bs.d_hist.p, 0, sizeof(uint64_t) * bs.lay.S * HIST_MAX_BINS, ctx->stream));
    DTX_CUDA(cudaStreamSynchronize(ctx->stream));
    return DTX_OK;
}

int dtx_get_fas(dtx_ctx* ctx, int set_id, double* fas, int64_t count, int reset) {
    if (!ctx) return DTX_ERR_ARG;
    auto it = ctx->sets.find(set_id);
    if (it == ctx->sets.end()) return fail(ctx, DTX_ERR_STATE, "dtx_get_fas: unknown basis set");
    BasisSet& bs = it->second;
    const int64_t nel = static_cast<int64_t>(bs.lay.S) * 5;
    if (fas && count < nel) return fail(ctx, DTX_ERR_CAPACITY, "dtx_get_fas: buffer too small");
    DTX_CUDA(cudaSetDevice(ctx->device));
    if (fas) DTX_CUDA(cudaMemcpyAsync(fas, bs.d_fas.p, sizeof(double) * nel, cudaMemcpyDeviceToHost, ctx->stream));
    if (reset) DTX_CUDA(cudaMemsetAsync(bs.d_fas.p, 0, sizeof(double) * nel, ctx->stream));
    DTX_CUDA(cudaStreamSynchronize(ctx->stream));
    return DTX_OK;
}

int dtx_get_candidates(dtx_ctx* ctx, dtx_cand* out, int64_t cap, int64_t* n) {
    if (!ctx || !n) return DTX_ERR_ARG;
    if (!ctx->ran) return fail(ctx, DTX_ERR_STATE, "dtx_get_candidates: no run");
    DTX_CUDA(cudaSetDevice(ctx->device));
    int nc = 0;
    DTX_CUDA(cudaMemcpyAsync(&nc, ctx->d_ncand.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    DTX_CUDA(cudaStreamSynchronize(ctx->stream));
    *n = nc;
    if (!out) {   // count query
        if (nc > ctx->cand_cap) return fail(ctx, DTX_ERR_CAPACITY, "candidate list truncated on the device");
        return DTX_OK;
    }
    int64_t ncopy = std::min<int64_t>(std::min<int64_t>(nc, ctx->cand_cap), cap);
    if (out && ncopy > 0) {
        static_assert(sizeof(dtx_cand) == sizeof(Candidate), "candidate layout");
        DTX_CUDA(cudaMemcpyAsync(out, ctx->d_cand.p, sizeof(Candidate) * ncopy, cudaMemcpyDeviceToHost, ctx->stream));
        DTX_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    if (nc > ctx->cand_cap || nc > cap) return fail(ctx, DTX_ERR_CAPACITY, "candidate list truncated");
    return DTX_OK;
}

int dtx_last_k1_ms(dtx_ctx* ctx, float* ms) {
    if (!ctx || !ms) return DTX_ERR_ARG;
    if (!ctx->k1_timed || ctx->k1_used == 0) return fail(ctx, DTX_ERR_STATE, "no tcgen05 run to time");
    DTX_CUDA(cudaSetDevice(ctx->device));
    const auto& ev = ctx->k1_events[ctx->k1_used - 1];
    DTX_CUDA(cudaEventSynchronize(ev.second));
    DTX_CUDA(cudaEventElapsedTime(ms, ev.first, ev.second));
    return DTX_OK;
}

int dtx_k1_ms_history(dtx_ctx* ctx, float* ms, int64_t cap, int64_t* n) {
    if (!ctx || !n) return DTX_ERR_ARG;
    DTX_CUDA(cudaSetDevice(ctx->device));
    *n = static_cast<int64_t>(ctx->k1_used);
    if (!ms) return DTX_OK;
    if (cap < *n) return fail(ctx, DTX_ERR_CAPACITY, "dtx_k1_ms_history: buffer too small");
    for (size_t i = 0; i < ctx->k1_used; ++i) {
        DTX_CUDA(cudaEventSynchronize(ctx->k1_events[i].second));
        DTX_CUDA(cudaEventElapsedTime(ms + i, ctx->k1_events[i].first, ctx->k1_events[i].second));
    }
    ctx->k1_used = 0;
    return DTX_OK;
}

// Shared body of dtx_preprocess_chunks / dtx_preprocess_chunks_dec.  factor > 1: ObsPy's
// Trace.decimate (forward-only low-pass SOS `dec_sos`, then data[::factor]) BEFORE detrend and
// band-pass, the order of construct._applyFilter (construct.py:1017-1029).
static int preprocess_impl(dtx_ctx* ctx, int nchunks, int Nc, const void* const* chan_ptrs, const int64_t* chan_len,
                           int dtype, const double* sos, int nsos, int zerophase, int detrend,
                           const double* dec_sos, int ndec, int factor) {
    if (!ctx || !chan_ptrs || !chan_len) return DTX_ERR_ARG;
    if (nchunks < 1 || Nc < 1 || nsos < 0 || (nsos > 0 && !sos)) return fail(ctx, DTX_ERR_ARG, "dtx_preprocess_chunks: bad arguments");
    if (factor < 1 || ndec < 0 || (ndec > 0 && !dec_sos)) return fail(ctx, DTX_ERR_ARG, "dtx_preprocess_chunks_dec: bad decimation arguments");
    if (dtype != DTX_F64 && dtype != DTX_F32) return fail(ctx, DTX_ERR_ARG, "dtx_preprocess_chunks: bad dtype");
    DTX_CUDA(cudaSetDevice(ctx->device));
    const int ntr = nchunks * Nc;
    std::vector<long long> off(ntr), out_off(nchunks), off2(ntr);
    std::vector<int> len(ntr), minlen(nchunks), len2(ntr);
    std::vector<int64_t> L(nchunks);
    long long tot = 0, tot2 = 0, otot = 0;
    int maxlen = 0, maxlen2 = 0;
    for (int ch = 0; ch < nchunks; ++ch) {
        int mn = INT32_MAX;
        for (int c = 0; c < Nc; ++c) {
            const int64_t l = chan_len[ch * Nc + c];
            if (l < 1 || l > (1LL << 30)) return fail(ctx, DTX_ERR_ARG, "dtx_preprocess_chunks: trace length out of range");
            off[ch * Nc + c] = tot;
            len[ch * Nc + c] = static_cast<int>(l);
            tot += (l + 1) & ~1LL;
            maxlen = std::max<int>(maxlen, static_cast<int>(l));
            const int64_t l2 = (l + factor - 1) / factor;      // len(data[::factor])
            off2[ch * Nc + c] = tot2;
            len2[ch * Nc + c] = static_cast<int>(l2);
            tot2 += (l2 + 1) & ~1LL;
            maxlen2 = std::max<int>(maxlen2, static_cast<int>(l2));
            mn = std::min<int>(mn, static_cast<int>(l2));
        }
        minlen[ch] = mn;            // multiplex trims to the shortest channel (construct.py:972-974)
        L[ch] = static_cast<int64_t>(mn) * Nc;
        out_off[ch] = otot;
        otot += (L[ch] + 1) & ~1LL;
    }
    cudaStream_t st = ctx->stream;
    DevBuf<double> dbuf, dbuf2, dstats, dseg;
    DevBuf<long long> doff, doff2, dooff;
    DevBuf<int> dlen, dlen2, dmin;
    const int maxseg = (maxlen + preproc_seg() - 1) / preproc_seg();
    DTX_CUDA(dbuf.reserve(tot)); DTX_CUDA(dstats.reserve(2 * static_cast<size_t>(ntr)));
    DTX_CUDA(dseg.reserve(static_cast<size_t>(ntr) * maxseg * 2));
    DTX_CUDA(doff.reserve(ntr)); DTX_CUDA(dlen.reserve(ntr)); DTX_CUDA(dmin.reserve(nchunks)); DTX_CUDA(dooff.reserve(nchunks));
    std::vector<double> conv;
    for (int t = 0; t < ntr; ++t) {
        const void* src = chan_ptrs[t];
        if (dtype == DTX_F32) {   // the reference filters float32 traces in float64 as well (scipy sosfilt)
            conv.resize(len[t]);
            const float* f = static_cast<const float*>(chan_ptrs[t]);
            for (int i = 0; i < len[t]; ++i) conv[i] = f[i];
            DTX_CUDA(cudaMemcpy(dbuf.p + off[t], conv.data(), sizeof(double) * len[t], cudaMemcpyHostToDevice));
        } else {
            DTX_CUDA(cudaMemcpyAsync(dbuf.p + off[t], src, sizeof(double) * len[t], cudaMemcpyHostToDevice, st));
        }
    }
    DTX_CUDA(cudaMemcpyAsync(doff.p, off.data(), sizeof(long long) * ntr, cudaMemcpyHostToDevice, st));
    DTX_CUDA(cudaMemcpyAsync(dlen.p, len.data(), sizeof(int) * ntr, cudaMemcpyHostToDevice, st));
    DTX_CUDA(cudaMemcpyAsync(dmin.p, minlen.data(), sizeof(int) * nchunks, cudaMemcpyHostToDevice, st));
    DTX_CUDA(cudaMemcpyAsync(dooff.p, out_off.data(), sizeof(long long) * nchunks, cudaMemcpyHostToDevice, st));
    // the channels of a chunk are trimmed to their common window BEFORE detrend and filter
    // (st.trim(startTrim, endTrim), construct.py:1019-1024; channels share a start time here, so
    // that is the first min-length samples of each, taken after the optional decimation)
    std::vector<int> lenw(ntr);
    for (int t = 0; t < ntr; ++t) lenw[t] = minlen[t / Nc];
    DevBuf<int> dlenw;
    DTX_CUDA(dlenw.reserve(ntr));
    DTX_CUDA(cudaMemcpyAsync(dlenw.p, lenw.data(), sizeof(int) * ntr, cudaMemcpyHostToDevice, st));
    double* work = dbuf.p;
    const long long* work_off = doff.p;
    const int* work_len = dlenw.p;
    int work_maxlen = maxlen;
    if (factor > 1) {
        // anti-alias low-pass (forward only, no detrend), then keep every factor-th sample
        launch_preproc(dbuf.p, doff.p, dlen.p, ntr, maxlen, dec_sos, ndec, 0, 0, dstats.p, dseg.p, st);
        DTX_CUDA(cudaGetLastError());
        DTX_CUDA(dbuf2.reserve(tot2)); DTX_CUDA(doff2.reserve(ntr)); DTX_CUDA(dlen2.reserve(ntr));
        DTX_CUDA(cudaMemcpyAsync(doff2.p, off2.data(), sizeof(long long) * ntr, cudaMemcpyHostToDevice, st));
        DTX_CUDA(cudaMemcpyAsync(dlen2.p, len2.data(), sizeof(int) * ntr, cudaMemcpyHostToDevice, st));
        launch_decimate(dbuf.p, doff.p, dbuf2.p, doff2.p, dlen2.p, ntr, factor, st);
        DTX_CUDA(cudaGetLastError());
        ctx->launches += 3 * ndec + 1;
        work = dbuf2.p; work_off = doff2.p; work_maxlen = maxlen2;
    }
    launch_preproc(work, work_off, work_len, ntr, work_maxlen, sos, nsos, zerophase, detrend, dstats.p, dseg.p, st);
    DTX_CUDA(cudaGetLastError());
    ctx->launches += (detrend ? 2 : 0) + 3 * nsos * (zerophase ? 2 : 1);
    // multiplexed result becomes the loaded batch
    int rc = set_chunk_table(ctx, nchunks, L.data(), nullptr, DTX_F64);
    if (rc) return rc;
    for (int ch = 0; ch < nchunks; ++ch)
        if (ctx->raw_off[ch] != out_off[ch]) return fail(ctx, DTX_ERR_STATE, "dtx_preprocess_chunks: layout mismatch");
    DTX_CUDA(ctx->raw_own.reserve(static_cast<size_t>(otot) * 8));
    launch_multiplex(work, work_off, dmin.p, dooff.p, nchunks, Nc, work_maxlen, reinterpret_cast<double*>(ctx->raw_own.p), st);
    DTX_CUDA(cudaGetLastError());
    mark_raw_access(ctx);
    ctx->launches += 1;
    ctx->d_raw = ctx->raw_own.p;
    DTX_CUDA(cudaStreamSynchronize(st));   // host trace buffers may be released by the caller
    dbuf.release(); dbuf2.release(); dstats.release(); dseg.release(); doff.release(); doff2.release();
    dooff.release(); dlen.release(); dlen2.release(); dmin.release();
    return DTX_OK;
}

int dtx_preprocess_chunks(dtx_ctx* ctx, int nchunks, int Nc, const void* const* chan_ptrs,
                          const int64_t* chan_len, int dtype, const double* sos, int nsos, int zerophase,
                          int detrend) {
    return preprocess_impl(ctx, nchunks, Nc, chan_ptrs, chan_len, dtype, sos, nsos, zerophase, detrend, nullptr, 0, 1);
}

int dtx_preprocess_chunks_dec(dtx_ctx* ctx, int nchunks, int Nc, const void* const* chan_ptrs,
                              const int64_t* chan_len, int dtype, const double* sos, int nsos, int zerophase,
                              int detrend, const double* dec_sos, int ndec, int factor) {
    return preprocess_impl(ctx, nchunks, Nc, chan_ptrs, chan_len, dtype, sos, nsos, zerophase, detrend, dec_sos, ndec,
                           factor);
}

int dtx_get_chunk(dtx_ctx* ctx, int chunk, double* out, int64_t count, int64_t* L) {
    if (!ctx || !out) return DTX_ERR_ARG;
    if (!ctx->d_raw || chunk < 0 || chunk >= ctx->nchunks) return fail(ctx, DTX_ERR_STATE, "dtx_get_chunk: no such chunk");
    if (ctx->dtype != DTX_F64) return fail(ctx, DTX_ERR_STATE, "dtx_get_chunk: float64 chunks only");
    if (L) *L = ctx->rawL[chunk];
    if (count < ctx->rawL[chunk]) return fail(ctx, DTX_ERR_CAPACITY, "dtx_get_chunk: buffer too small");
    DTX_CUDA(cudaSetDevice(ctx->device));
    DTX_CUDA(cudaMemcpyAsync(out, static_cast<const double*>(ctx->d_raw) + ctx->raw_off[chunk],
                             sizeof(double) * ctx->rawL[chunk], cudaMemcpyDeviceToHost, ctx->stream));
    DTX_CUDA(cudaStreamSynchronize(ctx->stream));
    return DTX_OK;
}

int dtx_sta_lta_max(dtx_ctx* ctx, int Nc, int chan, int nsta, int nlta, float* out, int64_t count) {
    if (!ctx || !out) return DTX_ERR_ARG;
    if (ctx->nchunks < 1 || !ctx->d_raw) return fail(ctx, DTX_ERR_STATE, "dtx_sta_lta_max: no chunks loaded");
    if (Nc < 1 || chan < 0 || chan >= Nc || nsta < 1 || nlta <= nsta)
        return fail(ctx, DTX_ERR_ARG, "dtx_sta_lta_max: need 0 <= chan < Nc and 1 <= nsta < nlta");
    if (count < ctx->nchunks) return fail(ctx, DTX_ERR_CAPACITY, "dtx_sta_lta_max: buffer too small");
    DTX_CUDA(cudaSetDevice(ctx->device));
    const int nch = ctx->nchunks;
    std::vector<long long> off(nch);
    std::vector<int> Ls(nch);
    int maxLs = 0;
    for (int i = 0; i < nch; ++i) {
        off[i] = ctx->raw_off[i];
        Ls[i] = static_cast<int>(ctx->rawL[i] / Nc);
        maxLs = std::max(maxLs, Ls[i]);
    }
    DevBuf<long long> doff;
    DevBuf<int> dLs;
    DevBuf<unsigned> dout;
    DTX_CUDA(doff.reserve(nch)); DTX_CUDA(dLs.reserve(nch)); DTX_CUDA(dout.reserve(nch));
    cudaStream_t st = ctx->stream;
    DTX_CUDA(cudaMemcpyAsync(doff.p, off.data(), sizeof(long long) * nch, cudaMemcpyHostToDevice, st));
    DTX_CUDA(cudaMemcpyAsync(dLs.p, Ls.data(), sizeof(int) * nch, cudaMemcpyHostToDevice, st));
    launch_stalta_max(ctx->d_raw, ctx->dtype == DTX_F32, doff.p, dLs.p, nch, maxLs, Nc, chan, nsta, nlta, dout.p, st);
    DTX_CUDA(cudaGetLastError());
    mark_raw_access(ctx);
    ctx->launches += 1;
    static_assert(sizeof(float) == sizeof(unsigned), "bit copy");
    DTX_CUDA(cudaMemcpyAsync(out, dout.p, sizeof(float) * nch, cudaMemcpyDeviceToHost, st));
    DTX_CUDA(cudaStreamSynchronize(st));
    doff.release(); dLs.release(); dout.release();
    return DTX_OK;
}

int dtx_set_events(dtx_ctx* ctx, int set_id, int subspace, int nev, const double* ewf, const double* mags,
                   const double* wfu_var, int is_single) {
    if (!ctx || !ewf || !mags) return DTX_ERR_ARG;
    auto it = ctx->sets.find(set_id);
    if (it == ctx->sets.end()) return fail(ctx, DTX_ERR_STATE, "dtx_set_events: unknown basis set");
    BasisSet& bs = it->second;
    if (subspace < 0 || subspace >= bs.lay.S || nev < 1) return fail(ctx, DTX_ERR_ARG, "dtx_set_events: bad index");
    if (!is_single && !wfu_var) return fail(ctx, DTX_ERR_ARG, "dtx_set_events: wfu_var required for subspaces");
    DTX_CUDA(cudaSetDevice(ctx->device));
    const int n = bs.lay.n;
    std::vector<double> blob(static_cast<size_t>(nev) * n + 4 * static_cast<size_t>(nev));
    double* pm = blob.data() + static_cast<size_t>(nev) * n;
    for (int i = 0; i < nev; ++i) {
        const double* e = ewf + static_cast<size_t>(i) * n;
        std::copy(e, e + n, blob.data() + static_cast<size_t>(i) * n);
        long double s = 0;
        for (int j = 0; j < n; ++j) s += e[j];
        const double mean = static_cast<double>(s / n);
        long double v = 0;
        for (int j = 0; j < n; ++j) v += (e[j] - mean) * (e[j] - mean);
        pm[i] = mags[i];
        pm[nev + i] = mean;
        pm[2 * nev + i] = std::sqrt(static_cast<double>(v / n));
        pm[3 * nev + i] = wfu_var ? wfu_var[i] : 0.0;
    }
    DevBuf<double>& d = bs.ev_blob[subspace];
    DTX_CUDA(d.reserve(blob.size()));
    DTX_CUDA(cudaStreamSynchronize(ctx->stream));
    DTX_CUDA(cudaMemcpy(d.p, blob.data(), sizeof(double) * blob.size(), cudaMemcpyHostToDevice));
    bs.ev_meta[subspace] = std::make_pair(nev, is_single);
    return DTX_OK;
}

int dtx_est_mags(dtx_ctx* ctx, int set_id, int ntrig, const int32_t* chunk, const int32_t* subspace,
                 const int32_t* t, double* out) {
    if (!ctx || !chunk || !subspace || !t || !out || ntrig < 0) return DTX_ERR_ARG;
    if (ntrig == 0) return DTX_OK;
    auto it = ctx->sets.find(set_id);
    if (it == ctx->sets.end()) return fail(ctx, DTX_ERR_STATE, "dtx_est_mags: unknown basis set");
    if (!ctx->ran || ctx->run_set != set_id) return fail(ctx, DTX_ERR_STATE, "dtx_est_mags: run dtx_detect_run on this set first");
    BasisSet& bs = it->second;
    DTX_CUDA(cudaSetDevice(ctx->device));
    const int n = bs.lay.n, S = bs.lay.S;
    std::vector<MagSubspace> subs(S);
    for (int s = 0; s < S; ++s) {
        MagSubspace& m = subs[s];
        std::memset(&m, 0, sizeof(m));
        m.row0 = bs.rank_off[s];
        m.rank = bs.rank_off[s + 1] - bs.rank_off[s];
        auto e = bs.ev_meta.find(s);
        if (e == bs.ev_meta.end()) continue;
        const int nev = e->second.first;
        const double* base = bs.ev_blob[s].p;
        m.nev = nev;
        m.is_single = e->second.second;
        m.ewf = base;
        m.mags = base + static_cast<size_t>(nev) * n;
        m.ev_mean = m.mags + nev;
        m.ev_std = m.mags + 2 * nev;
        m.wfu_var = m.mags + 3 * nev;
    }
    std::vector<MagTrigger> trig(ntrig);
    for (int i = 0; i < ntrig; ++i) {
        if (chunk[i] < 0 || chunk[i] >= ctx->nchunks || subspace[i] < 0 || subspace[i] >= S || t[i] < 0 ||
            t[i] >= ctx->h_chunks[chunk[i]].T)
            return fail(ctx, DTX_ERR_ARG, "dtx_est_mags: trigger out of range");
        if (subs[subspace[i]].nev == 0) return fail(ctx, DTX_ERR_STATE, "dtx_est_mags: dtx_set_events missing for a subspace");
        if (subs[subspace[i]].rank > MAG_MAX_RANK) return fail(ctx, DTX_ERR_ARG, "dtx_est_mags: rank > 64 not supported");
        trig[i].chunk = chunk[i]; trig[i].subspace = subspace[i]; trig[i].t = t[i]; trig[i].pad = 0;
    }
    const int stride = 6 * n + 8;
    DevBuf<MagSubspace> dsubs;
    DevBuf<MagTrigger> dtrig;
    DevBuf<double> dscr, dout;
    DTX_CUDA(dsubs.reserve(S)); DTX_CUDA(dtrig.reserve(ntrig));
    DTX_CUDA(dscr.reserve(static_cast<size_t>(ntrig) * stride)); DTX_CUDA(dout.reserve(static_cast<size_t>(ntrig) * 3));
    cudaStream_t st = ctx->stream;
    DTX_CUDA(cudaMemcpyAsync(dsubs.p, subs.data(), sizeof(MagSubspace) * S, cudaMemcpyHostToDevice, st));
    DTX_CUDA(cudaMemcpyAsync(dtrig.p, trig.data(), sizeof(MagTrigger) * ntrig, cudaMemcpyHostToDevice, st));
    launch_mag(ctx->d_raw, ctx->dtype == DTX_F32, ctx->d_chunks.p, ctx->d_sum.p, dtrig.p, ntrig, dsubs.p, bs.d_U.p, n,
               bs.lay.Nc, dscr.p, stride, dout.p, st);
    DTX_CUDA(cudaGetLastError());
    mark_raw_access(ctx);
    ctx->launches += 1;
    DTX_CUDA(cudaMemcpyAsync(out, dout.p, sizeof(double) * 3 * ntrig, cudaMemcpyDeviceToHost, st));
    DTX_CUDA(cudaStreamSynchronize(st));
    dsubs.release(); dtrig.release(); dscr.release(); dout.release();
    return DTX_OK;
}

int dtx_launch_count(dtx_ctx* ctx, int64_t* n) {
    if (!ctx || !n) return DTX_ERR_ARG;
    *n = ctx->launches;
    return DTX_OK;
}

int dtx_corr_zero_lag(dtx_ctx* ctx, const double* X, int N, int n, double* out) {
    if (!ctx || !X || !out) return DTX_ERR_ARG;
    if (N < 1 || n < 2) return fail(ctx, DTX_ERR_ARG, "dtx_corr_zero_lag: bad shape");
    DTX_CUDA(cudaSetDevice(ctx->device));
    DevBuf<double> dX, dout;
    DTX_CUDA(dX.reserve(static_cast<size_t>(N) * n));
    DTX_CUDA(dout.reserve(static_cast<size_t>(N) * N));
    cudaStream_t st = ctx->stream;
    DTX_CUDA(cudaMemcpyAsync(dX.p, X, sizeof(double) * N * n, cudaMemcpyHostToDevice, st));
    launch_corr0(dX.p, N, n, dout.p, st);
    DTX_CUDA(cudaGetLastError());
    ctx->launches += 1;
    DTX_CUDA(cudaMemcpyAsync(out, dout.p, sizeof(double) * N * N, cudaMemcpyDeviceToHost, st));
    DTX_CUDA(cudaStreamSynchronize(st));
    dX.release(); dout.release();
    return DTX_OK;
}

// ------------------------------------------------------------------------------------ CCX
// The loaded-chunk table and the basis sets of the detection path survive a CCX call; only the
// results of the last dtx_detect_run are invalidated (CCX shares the DS / split-signal workspace).
namespace {
struct BatchTable {
    int nchunks, dtype;
    std::vector<long long> raw_off, rawL;
    const void* d_raw;
};
BatchTable save_table(dtx_ctx* ctx) {
    return BatchTable{ctx->nchunks, ctx->dtype, ctx->raw_off, ctx->rawL, ctx->d_raw};
}
void restore_table(dtx_ctx* ctx, BatchTable& t) {
    ctx->nchunks = t.nchunks; ctx->dtype = t.dtype; ctx->raw_off.swap(t.raw_off); ctx->rawL.swap(t.rawL);
    ctx->d_raw = t.d_raw;
    ctx->ran = false;
    ctx->items_key.clear();
}
}  // namespace

// Tensor-core CCX: events rows[0..nrows) as rank-1 templates, padded events as chunks.
static int ccx_tcgen05(dtx_ctx* ctx, int dtype, const void* dX, int N, int n, int Nc, const int* h_rows, int nrows,
                       double* dcc, int* dlag, double* dsub) {
    const int ns = n / Nc, trunc = n / (2 * Nc) - 1, nl = 2 * ns - 1 - 2 * trunc;
    cudaStream_t st = ctx->stream;
    ctx->k1_used = 0;   // dtx_k1_ms_history after the call returns the K1 time of every signal batch
    // templates: x / ||x - mean||  (zero rows for zeroed-out waveforms), built on the device
    std::vector<int32_t> roff(nrows + 1);
    for (int r = 0; r <= nrows; ++r) roff[r] = r;
    auto fill = [&](double* d_U) -> int {
        launch_ccx_templates(dX, dtype == DTX_F32, n, ctx->cx_rows.p, nrows, d_U, st);
        ctx->launches += 1;
        DTX_CUDA(cudaGetLastError());
        return DTX_OK;
    };
    int rc = set_bases_device(ctx, ctx->ccx_set, INT32_MIN, fill, roff.data(), nrows, n, Nc, nullptr);
    if (rc != DTX_OK) return rc;
    BasisSet& bs = ctx->ccx_set;
    const int P = ns - trunc - 1;
    const int Lc = nl + ns - 1;
    const long long Lm = static_cast<long long>(Lc) * Nc;
    // batch of signals bounded by the DS buffer (rows x Tpad floats per signal)
    const long long per_sig = static_cast<long long>(nrows) * ((nl + TILE_T - 1) / TILE_T * TILE_T) * 4;
    int batch = static_cast<int>(std::max<long long>(
        1, std::min<long long>(ctx->ccx_max_batch, ctx->ccx_ds_bytes / std::max<long long>(1, per_sig))));
    {
        // the series buffer of a whole batch has to fit what is free on the device right now: halve the batch
        // until it does (the default budget is 16 GiB) instead of failing the call
        const long long tpad = (nl + TILE_T - 1) / TILE_T * TILE_T;
        const long long rows_ds = std::max<long long>(nrows, bs.ds_rows);
        const int nsig_max = std::max(1, N - (h_rows[0] + 1));
        for (;;) {
            const size_t elems = static_cast<size_t>(std::min(batch, nsig_max)) * rows_ds * tpad;
            if (ctx->d_DS.reserve(elems) == cudaSuccess) break;
            cudaGetLastError();
            if (batch <= 8) return fail(ctx, DTX_ERR_CUDA, "dtx_ccx: out of device memory for the correlation series");
            batch /= 2;
        }
    }
    const int flag_cap = 1 << 20;
    DTX_CUDA(ctx->cx_pad.reserve(static_cast<size_t>(batch) * Lm));
    DTX_CUDA(ctx->cx_karg.reserve(static_cast<size_t>(batch) * nrows));
    DTX_CUDA(ctx->cx_nflag.reserve(1));
    DTX_CUDA(ctx->cx_flag.reserve(flag_cap));
    DTX_CUDA(cudaMemsetAsync(ctx->cx_nflag.p, 0, sizeof(int), st));
    // float64 de-multiplexed copy of the events: the ring re-scoring stages a waveform with one bulk copy
    DTX_CUDA(ctx->cx_xd.reserve(static_cast<size_t>(N) * n));
    launch_ccx_demux(dX, dtype == DTX_F32, N, n, Nc, ctx->cx_xd.p, st);
    DTX_CUDA(cudaGetLastError());
    ctx->launches += 1;
    std::vector<int64_t> offs, lens;
    std::vector<int> blk_hi;
    const int c_first = h_rows[0] + 1;   // signals c <= rows[0] have no template b < c
    for (int c0 = c_first; c0 < N; c0 += batch) {
        const int nsig = std::min(batch, N - c0);
        launch_ccx_pad(dX, dtype == DTX_F32, n, Nc, c0, nsig, P, Lc, ctx->cx_pad.p, st);
        DTX_CUDA(cudaGetLastError());
        offs.resize(nsig); lens.resize(nsig); blk_hi.resize(nsig);
        for (int i = 0; i < nsig; ++i) {
            offs[i] = static_cast<int64_t>(i) * Lm;
            lens[i] = Lm;
            // templates b < c are a prefix of the (ascending) row list
            const int nrow = static_cast<int>(std::lower_bound(h_rows, h_rows + nrows, c0 + i) - h_rows);
            blk_hi[i] = (nrow + VEC_PER_BLOCK - 1) / VEC_PER_BLOCK;
        }
        rc = dtx_attach_device_chunks(ctx, nsig, ctx->cx_pad.p, offs.data(), lens.data(), DTX_F64);
        if (rc != DTX_OK) return rc;
        // The float32 series only LOCATES the maximum (every lag within a band of it is re-scored in
        // float64), so by default it is computed with ONE fp16 MMA per K step: operands rounded to 11
        // bits move a normalised value by at most 2 * 2^-11 (Cauchy-Schwarz) times the amplification
        // factors the scan applies, and accumulating all taps in TMEM adds ~2e-5 of truncation bias.
        // band0 = 2 * (2^-10 + 5e-5) covers both operands of the comparison.  passes = 3: the
        // fp16x3 series of the detection path (error ~2e-6, band 3e-5).
        const int hi_only = ctx->ccx_passes == 1;
        const float band0 = hi_only ? 2.2e-3f : 3e-5f;
        rc = project_run(ctx, bs, DTX_ENGINE_TCGEN05, hi_only ? bs.lay.nchunks : 2, 1, 0, blk_hi.data(), hi_only);
        if (rc != DTX_OK) return rc;
        launch_ccx_post(ctx->d_DS.p, ctx->d_chunks.p, c0, nsig, dX, dtype == DTX_F32, ctx->cx_xd.p, N, n, Nc, ctx->cx_rows.p, nrows,
                        ctx->cx_wa.p, ctx->cx_wb.p, ctx->cx_es.p, ctx->cx_ed.p, dcc, dlag, dsub, ctx->cx_nflag.p,
                        ctx->cx_flag.p, flag_cap, ctx->cx_karg.p, ctx->d_k4bits.p, band0, st);
        DTX_CUDA(cudaGetLastError());
        ctx->launches += 4;   // pad, scan, tiled re-scoring, leftover pairs
    }
    // degenerate pairs (|res| > 1 from zero-variance windows, a crowded maximum, NaN): the float64
    // kernel re-does exactly those pairs, from the device-side list -- no host round trip
    launch_ccx_fp64_pairs(dX, dtype == DTX_F32, N, n, Nc, ctx->cx_rows.p, ctx->cx_wa.p, ctx->cx_wb.p, ctx->cx_es.p,
                          ctx->cx_ed.p, dcc, dlag, dsub, ctx->cx_flag.p, ctx->cx_nflag.p, flag_cap, ctx->num_sms, st);
    DTX_CUDA(cudaGetLastError());
    ctx->launches += 1;
    DTX_CUDA(ctx->cx_nflag_h.reserve(1));
    DTX_CUDA(cudaMemcpyAsync(ctx->cx_nflag_h.p, ctx->cx_nflag.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    ctx->cx_flag_check = true;
    return DTX_OK;
}

// After a stream synchronisation: did the degenerate-pair list of the last CCX call overflow?
static int ccx_check_flags(dtx_ctx* ctx) {
    if (!ctx->cx_flag_check) return DTX_OK;
    ctx->cx_flag_check = false;
    if (ctx->cx_nflag_h.p && *ctx->cx_nflag_h.p > (1 << 20))
        return fail(ctx, DTX_ERR_CAPACITY, "dtx_ccx: more than 2^20 degenerate pairs (use DTX_ENGINE_FP64)");
    return DTX_OK;
}

// Shared body: X on the host (x_on_device = 0, copied H2D) or already on the device; results into
// the device arrays d_cc / d_lag / d_sub, dense [nrows][N] (slot r = event rows[r]).
static int ccx_run(dtx_ctx* ctx, const void* X, int x_on_device, int dtype, int N, int n, int Nc, const int32_t* rows,
                   int nrows, int engine, double* d_cc, int* d_lag, double* d_sub) {
    if (N < 2 || Nc < 1 || n % Nc != 0 || n / Nc < 4)
        return fail(ctx, DTX_ERR_ARG, "dtx_ccx: lengths not equal / not a multiple of Nc (construct.py:430-436)");
    if (!rows || nrows < 1) return fail(ctx, DTX_ERR_ARG, "dtx_ccx: empty row list");
    for (int r = 0; r < nrows; ++r)
        if (rows[r] < 0 || rows[r] >= N || (r > 0 && rows[r] <= rows[r - 1]))
            return fail(ctx, DTX_ERR_ARG, "dtx_ccx: rows must be ascending, unique and inside [0, N)");
    if (dtype != DTX_F64 && dtype != DTX_F32) return fail(ctx, DTX_ERR_ARG, "dtx_ccx: bad dtype");
    if (engine != DTX_ENGINE_TCGEN05 && engine != DTX_ENGINE_FP64) return fail(ctx, DTX_ERR_ARG, "dtx_ccx: bad engine");
    DTX_CUDA(cudaSetDevice(ctx->device));
    const int ns = n / Nc, trunc = n / (2 * Nc) - 1, nl = 2 * ns - 1 - 2 * trunc;
    const size_t esz = dtype == DTX_F32 ? 4 : 8;
    cudaStream_t st = ctx->stream;
    const void* dX = X;
    if (!x_on_device) {
        DTX_CUDA(ctx->cx_X.reserve(static_cast<size_t>(N) * n * esz));
        DTX_CUDA(cudaMemcpyAsync(ctx->cx_X.p, X, static_cast<size_t>(N) * n * esz, cudaMemcpyHostToDevice, st));
        dX = ctx->cx_X.p;
    }
    DTX_CUDA(ctx->cx_wa.reserve(static_cast<size_t>(N) * nl));
    DTX_CUDA(ctx->cx_wb.reserve(static_cast<size_t>(N) * nl));
    DTX_CUDA(ctx->cx_es.reserve(N));
    DTX_CUDA(ctx->cx_ed.reserve(N));
    DTX_CUDA(ctx->cx_rows.reserve(nrows));
    DTX_CUDA(cudaMemcpyAsync(ctx->cx_rows.p, rows, sizeof(int) * nrows, cudaMemcpyHostToDevice, st));
    const size_t cells = static_cast<size_t>(nrows) * N;
    DTX_CUDA(cudaMemsetAsync(d_cc, 0, cells * sizeof(double), st));
    DTX_CUDA(cudaMemsetAsync(d_sub, 0, cells * sizeof(double), st));
    DTX_CUDA(cudaMemsetAsync(d_lag, 0, cells * sizeof(int), st));
    launch_ccx_stats(dX, dtype == DTX_F32, N, n, Nc, ctx->cx_wa.p, ctx->cx_wb.p, ctx->cx_es.p, ctx->cx_ed.p, st);
    DTX_CUDA(cudaGetLastError());
    ctx->launches += 1;
    if (engine == DTX_ENGINE_FP64 || nl < 16) {  // tiny templates: not worth a GEMM
        launch_ccx_fp64(dX, dtype == DTX_F32, N, n, Nc, 0, nrows, ctx->cx_rows.p, ctx->cx_wa.p, ctx->cx_wb.p,
                        ctx->cx_es.p, ctx->cx_ed.p, d_cc, d_lag, d_sub, ctx->num_sms, st);
        DTX_CUDA(cudaGetLastError());
        ctx->launches += 1;
        return DTX_OK;
    }
    BatchTable saved = save_table(ctx);
    const int rc = ccx_tcgen05(ctx, dtype, dX, N, n, Nc, rows, nrows, d_cc, d_lag, d_sub);
    restore_table(ctx, saved);
    return rc;
}

int dtx_ccx_device(dtx_ctx* ctx, const void* X, int x_on_device, int dtype, int N, int n, int Nc,
                   const int32_t* rows, int nrows, int engine, double* d_cc, int32_t* d_lag, double* d_sub) {
    if (!ctx) return DTX_ERR_ARG;
    if (!X || !d_cc || !d_lag || !d_sub) return fail(ctx, DTX_ERR_ARG, "dtx_ccx_device: null pointer");
    return ccx_run(ctx, X, x_on_device, dtype, N, n, Nc, rows, nrows, engine, d_cc, d_lag, d_sub);
}

int dtx_ccx(dtx_ctx* ctx, const void* X, int dtype, int N, int n, int Nc, int row_begin, int row_end,
            int engine, double* cc, int32_t* lag, double* subsamp) {
    if (!ctx) return DTX_ERR_ARG;
    if (!X || !cc || !lag || !subsamp) return fail(ctx, DTX_ERR_ARG, "dtx_ccx: null pointer");
    if (row_begin < 0 || row_end > N || row_begin >= row_end) return fail(ctx, DTX_ERR_ARG, "dtx_ccx: bad row range");
    DTX_CUDA(cudaSetDevice(ctx->device));
    const size_t nrows = static_cast<size_t>(row_end - row_begin);
    std::vector<int32_t> rows(nrows);
    for (size_t r = 0; r < nrows; ++r) rows[r] = row_begin + static_cast<int>(r);
    DTX_CUDA(ctx->cx_cc.reserve(nrows * N));
    DTX_CUDA(ctx->cx_sub.reserve(nrows * N));
    DTX_CUDA(ctx->cx_lag.reserve(nrows * N));
    const int rc = ccx_run(ctx, X, 0, dtype, N, n, Nc, rows.data(), static_cast<int>(nrows), engine, ctx->cx_cc.p,
                           ctx->cx_lag.p, ctx->cx_sub.p);
    if (rc != DTX_OK) return rc;
    cudaStream_t st = ctx->stream;
    DTX_CUDA(cudaMemcpyAsync(cc, ctx->cx_cc.p, nrows * N * sizeof(double), cudaMemcpyDeviceToHost, st));
    DTX_CUDA(cudaMemcpyAsync(lag, ctx->cx_lag.p, nrows * N * sizeof(int), cudaMemcpyDeviceToHost, st));
    DTX_CUDA(cudaMemcpyAsync(subsamp, ctx->cx_sub.p, nrows * N * sizeof(double), cudaMemcpyDeviceToHost, st));
    DTX_CUDA(cudaStreamSynchronize(st));
    return ccx_check_flags(ctx);
}

int dtx_ccx_pack(dtx_ctx* ctx, const double* d_cc, const int32_t* d_lag, const double* d_sub,
                 const int32_t* slot_rows, int nslots, int N, double* cc, int32_t* lag, double* subsamp) {
    if (!ctx) return DTX_ERR_ARG;
    if (!d_cc || !d_lag || !d_sub || !slot_rows || !cc || !lag || !subsamp || N < 2 || nslots < N - 1)
        return fail(ctx, DTX_ERR_ARG, "dtx_ccx_pack: bad arguments");
    DTX_CUDA(cudaSetDevice(ctx->device));
    std::vector<int> slot_of_row(N, -1);
    for (int s = 0; s < nslots; ++s)
        if (slot_rows[s] >= 0) {
            if (slot_rows[s] >= N) return fail(ctx, DTX_ERR_ARG, "dtx_ccx_pack: slot row out of range");
            slot_of_row[slot_rows[s]] = s;
        }
    for (int b = 0; b < N - 1; ++b)
        if (slot_of_row[b] < 0) return fail(ctx, DTX_ERR_ARG, "dtx_ccx_pack: a row is missing from the slots");
    const size_t np = static_cast<size_t>(N) * (N - 1) / 2;
    DTX_CUDA(ctx->cx_slot.reserve(N));
    DTX_CUDA(ctx->cx_pcc.reserve(np)); DTX_CUDA(ctx->cx_psub.reserve(np)); DTX_CUDA(ctx->cx_plag.reserve(np));
    cudaStream_t st = ctx->stream;
    DTX_CUDA(cudaMemcpyAsync(ctx->cx_slot.p, slot_of_row.data(), sizeof(int) * N, cudaMemcpyHostToDevice, st));
    launch_ccx_pack(d_cc, d_lag, d_sub, ctx->cx_slot.p, N, ctx->cx_pcc.p, ctx->cx_plag.p, ctx->cx_psub.p, st);
    DTX_CUDA(cudaGetLastError());
    ctx->launches += 1;
    DTX_CUDA(cudaMemcpyAsync(cc, ctx->cx_pcc.p, np * sizeof(double), cudaMemcpyDeviceToHost, st));
    DTX_CUDA(cudaMemcpyAsync(lag, ctx->cx_plag.p, np * sizeof(int), cudaMemcpyDeviceToHost, st));
    DTX_CUDA(cudaMemcpyAsync(subsamp, ctx->cx_psub.p, np * sizeof(double), cudaMemcpyDeviceToHost, st));
    DTX_CUDA(cudaStreamSynchronize(st));
    return ccx_check_flags(ctx);
}

int dtx_ccx_pack_rows(dtx_ctx* ctx, const double* d_cc, const int32_t* d_lag, const double* d_sub,
                      const int32_t* rows, int nrows, int N, double* cc, int32_t* lag, double* subsamp) {
    if (!ctx) return DTX_ERR_ARG;
    if (!d_cc || !d_lag || !d_sub || !rows || !cc || !lag || !subsamp || N < 2 || nrows < 0)
        return fail(ctx, DTX_ERR_ARG, "dtx_ccx_pack_rows: bad arguments");
    DTX_CUDA(cudaSetDevice(ctx->device));
    for (int r = 0; r < nrows; ++r)
        if (rows[r] < 0 || rows[r] >= N - 1) return fail(ctx, DTX_ERR_ARG, "dtx_ccx_pack_rows: row out of range");
    // the outputs are written by the kernel itself: they must be visible to the device (page-locked + mapped)
    void *p_cc = nullptr, *p_lag = nullptr, *p_sub = nullptr;
    if (cudaHostGetDevicePointer(&p_cc, cc, 0) != cudaSuccess || cudaHostGetDevicePointer(&p_lag, lag, 0) != cudaSuccess ||
        cudaHostGetDevicePointer(&p_sub, subsamp, 0) != cudaSuccess) {
        cudaGetLastError();
        return fail(ctx, DTX_ERR_ARG, "dtx_ccx_pack_rows: outputs must be page-locked (dtx_host_alloc / dtx_host_register)");
    }
    cudaStream_t st = ctx->stream;
    if (nrows > 0) {
        DTX_CUDA(ctx->cx_slot.reserve(nrows));
        DTX_CUDA(cudaMemcpyAsync(ctx->cx_slot.p, rows, sizeof(int) * nrows, cudaMemcpyHostToDevice, st));
        launch_ccx_pack_rows(d_cc, d_lag, d_sub, ctx->cx_slot.p, nrows, N, static_cast<double*>(p_cc),
                             static_cast<int*>(p_lag), static_cast<double*>(p_sub), st);
        DTX_CUDA(cudaGetLastError());
        ctx->launches += 1;
    }
    DTX_CUDA(cudaStreamSynchronize(st));
    return ccx_check_flags(ctx);
}

int dtx_ccx_condensed(dtx_ctx* ctx, const void* X, int dtype, int N, int n, int Nc, int engine, double* cc,
                      int32_t* lag, double* subsamp) {
    if (!ctx) return DTX_ERR_ARG;
    if (!X || !cc || !lag || !subsamp || N < 2) return fail(ctx, DTX_ERR_ARG, "dtx_ccx_condensed: bad arguments");
    DTX_CUDA(cudaSetDevice(ctx->device));
    const size_t nrows = static_cast<size_t>(N - 1);
    std::vector<int32_t> rows(nrows);
    for (size_t r = 0; r < nrows; ++r) rows[r] = static_cast<int>(r);
    DTX_CUDA(ctx->cx_cc.reserve(nrows * N));
    DTX_CUDA(ctx->cx_sub.reserve(nrows * N));
    DTX_CUDA(ctx->cx_lag.reserve(nrows * N));
    const int rc = ccx_run(ctx, X, 0, dtype, N, n, Nc, rows.data(), static_cast<int>(nrows), engine, ctx->cx_cc.p,
                           ctx->cx_lag.p, ctx->cx_sub.p);
    if (rc != DTX_OK) return rc;
    return dtx_ccx_pack(ctx, ctx->cx_cc.p, ctx->cx_lag.p, ctx->cx_sub.p, rows.data(), static_cast<int>(nrows), N, cc,
                        lag, subsamp);
}

int dtx_set_ccx_passes(dtx_ctx* ctx, int passes) {
    if (!ctx) return DTX_ERR_ARG;
    if (passes != 1 && passes != 3) return fail(ctx, DTX_ERR_ARG, "dtx_set_ccx_passes: 1 (hi*hi screening) or 3 (fp16x3)");
    ctx->ccx_passes = passes;
    return DTX_OK;
}

int dtx_set_ccx_batch(dtx_ctx* ctx, int max_signals, int64_t ds_bytes) {
    if (!ctx) return DTX_ERR_ARG;
    if (max_signals < 1 || ds_bytes < (1 << 20)) return fail(ctx, DTX_ERR_ARG, "dtx_set_ccx_batch: bad limits");
    ctx->ccx_max_batch = max_signals;
    ctx->ccx_ds_bytes = ds_bytes;
    return DTX_OK;
}

int dtx_host_alloc(void** out, int64_t bytes) {
    if (!out || bytes < 1) return DTX_ERR_ARG;
    return cudaMallocHost(out, static_cast<size_t>(bytes)) == cudaSuccess ? DTX_OK : DTX_ERR_CUDA;
}

int dtx_host_register(void* p, int64_t bytes) {
    if (!p || bytes < 1) return DTX_ERR_ARG;
    const cudaError_t e = cudaHostRegister(p, static_cast<size_t>(bytes), cudaHostRegisterPortable | cudaHostRegisterMapped);
    if (e != cudaSuccess) { cudaGetLastError(); return DTX_ERR_CUDA; }
    return DTX_OK;
}

int dtx_host_unregister(void* p) {
    if (!p) return DTX_OK;
    if (cudaHostUnregister(p) != cudaSuccess) { cudaGetLastError(); return DTX_ERR_CUDA; }
    return DTX_OK;
}

int dtx_host_free(void* p) {
    if (!p) return DTX_OK;
    return cudaFreeHost(p) == cudaSuccess ? DTX_OK : DTX_ERR_CUDA;
}

}  // extern "C"
