// k1_project.cu -- K1+K2: the subspace projection as a Hankel-tiled tcgen05 GEMM with the
// detection-statistic normalisation fused into its epilogue.
//
// Replaces, for ALL subspaces of a station at once, the per-subspace FFT correlation
//   m1 = ssFD * MPconFD ; if1 = real(ifft(m1))[:, n-1:L] - av_norm ; sum(if1^2)/b ; [::Nc]
// of `_SSDetex._MPXDS` (reference detex/detect.py:559-578) == `fas._MPXSSCorr`
// (detex/fas.py:120-134).  In MODE 1 the same contraction yields the signed normalised
// cross-correlation series of `construct._CCX2` (detex/construct.py:438-453) with events as
// rank-1 templates and zero-padded events as the data.
//
// Math (per channel c, de-multiplexed; exact because only channel-aligned lags are kept):
//   P[k,t]  = sum_c sum_j U_c[k,j] x_c[t+j]
//   DS[s,t] = ((n-1)/n) sum_{k in s} (P[k,t] - mu[t] sumU[k])^2 / (S2[t] - S1[t]^2/n)
//
// GEMM mapping.  fp16 elements are 2 B, a UMMA core-matrix row is 16 B = 8 elements, so the
// output lag is split t = t0 + 8q + p.  For phase p the basis is shifted instead of the data:
//   P[k, t0+8q+p] = sum_j' u_k[j'-p] x[t0 + 8q + j']
//   D[(k,p), q]   = sum_j' A[(k,p), j'] B[q, j'],  A[(k,p),j'] = u_k[j'-p],  B[q,j'] = x[t0+8q+j']
// * A (M = 128 rows = 16 basis vectors x 8 phases) is a plain matrix: pre-built once per
//   basis set in HBM in the exact no-swizzle K-major smem image (Aimg) and streamed through a
//   4-stage ring with 32 KB bulk copies (UBLKCP); it is L2-resident across the 148 CTAs.
// * B (N = NQ rows) is a HANKEL matrix and is never materialised: row q starts 8 elements
//   = 16 B after row q-1, which is exactly the row pitch inside a core matrix.  A K-major
//   no-swizzle descriptor with LBO = 16 B (next K core matrix) and SBO = 128 B (next 8 rows)
//   makes the tensor core read overlapping core matrices straight out of the 1-D signal
//   span held in smem (verified on hardware: profiles/r01_hankel_probe.log).
// * fp32-equivalent precision from fp16 tensor cores: both operands are split hi+lo (22 bits)
//   and three MMAs are issued per K step (hi*hi, hi*lo, lo*hi).
// * The tensor core TRUNCATES when it accumulates into fp32 TMEM (measured bias ~ -1e-7 per
//   MMA relative to the running sum), so K is accumulated in TMEM only over `kblk` 64-tap
//   chunks; 8 drain warps pull each partial sum out of TMEM (double-buffered accumulators)
//   and add it to register accumulators with round-to-nearest.
//
// One persistent CTA per SM; work item = (chunk, tile of 8*NQ lags, basis block).  The host
// orders the items (chunk group, block, tile) so that at any time all 148 CTAs stream the SAME
// 4.6 MB A block from L2 and a group's split signal stays L2-resident: DRAM traffic is then
// about the algorithmic minimum (inputs once per group + the dense DS write).
#include <algorithm>

#include "dtx_kernels.cuh"
#include "tc_common.cuh"

namespace dtx {
namespace {

constexpr int STAGES = 4;
constexpr int STAGE_BYTES = 32768;                      // A_hi tile | A_lo tile (16 KB each)
constexpr int TILE_BYTES = 16384;
constexpr int SIG_HALFS = TILE_T + MAX_SEG_TAPS;        // per hi or lo span (sized for NQ = 256)
constexpr int SIG_BUF_BYTES = 2 * SIG_HALFS * 2;        // hi + lo
constexpr int NORM_BUF_BYTES = 2 * TILE_T * 4;          // mu + invE
constexpr int NTHREADS = 384;                           // warps 0-3: producer, MMA, 2 idle; warps 4-11: drain
constexpr int FIRST_DRAIN_WARP = 4;
constexpr int NDRAIN_WARPS = 8;
constexpr int REGS_CTRL = 56, REGS_DRAIN = 224;         // setmaxnreg budgets (128*56 + 256*224 = 64512)
// Epilogue transposition buffer (one per half of the accumulator columns): the 8 phases of a lag
// octet live in 4 different warps, so DS is staged [slot][128 lags] in smem and written out as
// 512-byte rows.  Row stride +4 floats keeps rows 16 B aligned; slots >= 8 store with column ^ 2 so
// the 32 lanes of a store hit 32 banks.
constexpr int EPI_QC = 16;                               // accumulator columns per chunk
constexpr int EPI_LAGS = EPI_QC * 8;                     // = 128 lags per chunk
constexpr int EPI_STRIDE = EPI_LAGS + 4;                 // floats
constexpr int EPI_BUF_BYTES = VEC_PER_BLOCK * EPI_STRIDE * 4;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 2 * SIG_BUF_BYTES + 2 * NORM_BUF_BYTES + 2 * EPI_BUF_BYTES;

// Issue style of the producer / MMA warps.  Default: the whole warp runs the role loop converged
// and an elected lane issues (descriptors stay in uniform registers, the MMAs of a stage go out
// back to back).  -DDTX_SINGLE_LANE_ISSUE: ONE lane runs the loop (ptxas then wraps every UTCHMMA /
// UBLKCP in an ELECT loop, ~128 cycles per MMA, i.e. the issuer paces the tensor pipe).  Same-box
// A/B (experiments/gpu_round1k.sh, profiles/r01_issue_ab.md): converged is +0.8 % with 3 MMAs per
// K step and +2.2 % with the 8-bit cross terms, where the pipe is not saturated.
// Drain warps add each TMEM partial sum to their register accumulators with packed fp32x2 additions
// (FADD2: 64 instead of 128 instructions per drain; same round-to-nearest results).  1 = on.
#ifndef DTX_FADD2
#define DTX_FADD2 1
#endif
// Software-pipelined drain (needs DTX_FADD2): 1 = on.
#ifndef DTX_DRAIN_PIPE
#define DTX_DRAIN_PIPE 0
#endif
// A-operand collector reuse between the hi*hi and hi*lo MMAs of a K step (1 = on).
#ifndef DTX_COLLECTOR_A
#define DTX_COLLECTOR_A 1
#endif
#ifndef DTX_SINGLE_LANE_ISSUE
#define ISSUE_LANE elect_one()
#define ISSUE_SYNC() __syncwarp()
#define ROLE_LANES(lane) true
#else
#define ISSUE_LANE true
#define ISSUE_SYNC() ((void)0)
#define ROLE_LANES(lane) ((lane) == 0)
#endif

struct K1Params {
    K1Args a;
    int nblocks, nchunks, nseg;
    float u_inv_scale;
    float inv_cn;  // n / (n-1)
    Seg seg[MAX_SEGS];
};

// A tile image, no-swizzle K-major: 16 row groups x 8 k-cores x (8 rows x 16 B)
//   byte(r, jj) = (r/8)*1024 + (jj/8)*128 + (r%8)*16 + (jj%8)*2
constexpr uint32_t A_LBO = 128, A_SBO = 1024;
// Hankel view of the 1-D signal
constexpr uint32_t B_LBO = 16, B_SBO = 128;

struct Ring {
    int idx = 0;
    uint32_t phase = 0;
    int n;
    __device__ explicit Ring(int n_) : n(n_) {}
    __device__ void advance() {
        if (++idx == n) {
            idx = 0;
            phase ^= 1;
        }
    }
};

struct Smem {
    uint8_t* stage;
    uint8_t* sig;
    uint8_t* norm;
    uint8_t* epi;
    uint64_t *full, *empty, *sigfull, *sigempty, *accfull, *accempty, *normfull, *normempty;
    uint64_t *pfull, *psigfull;   // CTA pairs: the peer's stage / signal span has landed (leader's copies are used)
};

// 64-tap stages accumulated in TMEM before accumulation `nacc` of a work item is drained.  The
// first FIRST_LONG_ACCS accumulations of an item are twice as long: while the drain warps are
// still in the previous item's epilogue the MMA warp then has four instead of two kblk's of work
// before it runs out of accumulators (the truncation bias of those two partial sums doubles, over
// ~3 % of the taps).
#ifndef DTX_FIRST_LONG_ACCS
#define DTX_FIRST_LONG_ACCS 2
#endif
// Only for long templates (>= 16 kblk stages): there the two long accumulations are a few percent of the taps;
// for a short template they would be most of them and their doubled truncation bias the whole error
// (ns = 300, kblk 3: max |DS - float64| 3.6e-6 with, 2e-6 without).
#ifndef DTX_FIRST_LONG_FACTOR
#define DTX_FIRST_LONG_FACTOR 2
#endif
__device__ __forceinline__ int acc_stages(int nacc, int kblk, int nstages) {
    return (nacc < DTX_FIRST_LONG_ACCS && nstages >= 16 * kblk) ? DTX_FIRST_LONG_FACTOR * kblk : kblk;
}

// precision mode of a chunk: 1 = both cross terms in one 8-bit MMA (decided on the device by k0_split)
__device__ __forceinline__ int item_x8(const K1Params& P, int chunk) {
    return P.a.chunk_mode ? P.a.chunk_mode[chunk] : 0;
}

// L2 eviction hints of the producer's bulk copies (profiles/r01_k1_traffic_ab.md).
// DTX_SIG_HINT = 1 (default): the split signal and the norm tiles are copied evict_last -- they are
// re-read once per basis-block pass (~1 ms apart) and the 18 GB DS write stream of a launch would
// otherwise push them out (DRAM reads of a 48-chunk launch 16.0 -> 11.2 GB, the level of a build
// that does not write DS at all).  DTX_A_HINT = 1 marks the basis image evict_first; measured
// dead end (318 GB of reads: the CTAs stream a block at slightly different times), kept for the record.
#ifndef DTX_SIG_HINT
#define DTX_SIG_HINT 1
#endif
#ifndef DTX_A_HINT
#define DTX_A_HINT 0
#endif
#if DTX_SIG_HINT
#define SIG_G2S(dst, src, bytes, bar) bulk_g2s_hint(dst, src, bytes, bar, pol_sig)
#else
#define SIG_G2S(dst, src, bytes, bar) bulk_g2s(dst, src, bytes, bar)
#endif
#if DTX_A_HINT
#define A_G2S(dst, src, bytes, bar) bulk_g2s_hint(dst, src, bytes, bar, pol_a)
#else
#define A_G2S(dst, src, bytes, bar) bulk_g2s(dst, src, bytes, bar)
#endif

// ------------------------------------------------ producer (warp 0, an elected lane issues)
// Work items of a CTA: every CTA its own (blockIdx.x, stride gridDim.x), or -- CTA pairs -- every cluster
// its own, the two CTAs of the pair taking basis blocks it.z and it.z + 1 of the same (chunk, tile).
template <bool CG2> __device__ __forceinline__ int item_first() { return CG2 ? cluster_id_x() : blockIdx.x; }
template <bool CG2> __device__ __forceinline__ int item_step() { return CG2 ? cluster_nctaid_x() : gridDim.x; }

// DUAL (CCX screening series, hi-only): an item is a PAIR of chunks (it.x, it.w) of at most 1024 lags each against
// one basis block.  Each chunk supplies 128 B rows from its own span (the second one sits in the unused lo-plane
// slot), two N = 128 MMAs per K step share the A tile, and the halves of the 256-column accumulator, of the norm
// tile and of the read-out belong to the two chunks -- the basis image, whose stream out of L2 bounds the
// one-chunk form of this mode (786 KB per item at n = 3000, 7 TB/s over the 148 SMs), is read once per pair.
template <int NQ, bool CG2, bool DUAL>
__device__ __forceinline__ void producer_loop(const K1Params& P, const Smem& S) {
    constexpr int TT = 8 * NQ;
    constexpr int TSPAN = (CG2 || DUAL) ? TT / 2 : TT;   // lags whose B rows this CTA supplies (per chunk of a pair)
    const int brank = CG2 ? cluster_ctarank() : 0;
    Ring st(STAGES), sg(2), nm(2);
    const bool hi_only = P.a.hi_only != 0;
    const uint32_t a_bytes = hi_only ? TILE_BYTES : STAGE_BYTES;
#if DTX_SIG_HINT
    const uint64_t pol_sig = l2_policy_evict_last();
#endif
#if DTX_A_HINT
    const uint64_t pol_a = l2_policy_evict_first();
#endif
    for (int item = item_first<CG2>(); item < P.a.nitems; item += item_step<CG2>()) {
        const int4 it = P.a.items[item];
        const ChunkDesc cd = P.a.chunks[it.x];
        const ChunkDesc cd1 = P.a.chunks[DUAL ? it.w : it.x];
        // window mean / inverse energy of this tile
        mbar_wait(&S.normempty[nm.idx], nm.phase ^ 1);
        if (ISSUE_LANE) {
            mbar_arrive_expect_tx(&S.normfull[nm.idx], 2 * TT * 4);
            if (DUAL) {
                uint8_t* nb = S.norm + nm.idx * NORM_BUF_BYTES;
                const long long o0 = cd.norm_off + static_cast<long long>(it.y) * TSPAN;
                const long long o1 = cd1.norm_off + static_cast<long long>(it.y) * TSPAN;
                SIG_G2S(nb, P.a.mu + o0, TSPAN * 4, &S.normfull[nm.idx]);
                SIG_G2S(nb + TSPAN * 4, P.a.mu + o1, TSPAN * 4, &S.normfull[nm.idx]);
                SIG_G2S(nb + TILE_T * 4, P.a.invE + o0, TSPAN * 4, &S.normfull[nm.idx]);
                SIG_G2S(nb + TILE_T * 4 + TSPAN * 4, P.a.invE + o1, TSPAN * 4, &S.normfull[nm.idx]);
            } else {
                SIG_G2S(S.norm + nm.idx * NORM_BUF_BYTES, P.a.mu + cd.norm_off + static_cast<long long>(it.y) * TT,
                         TT * 4, &S.normfull[nm.idx]);
                SIG_G2S(S.norm + nm.idx * NORM_BUF_BYTES + TILE_T * 4,
                         P.a.invE + cd.norm_off + static_cast<long long>(it.y) * TT, TT * 4, &S.normfull[nm.idx]);
            }
        }
        ISSUE_SYNC();
        nm.advance();
        const __half* sig0 = P.a.xsplit + cd.sig_off + static_cast<long long>(it.y) * (DUAL ? TSPAN : TT) + brank * TSPAN;
        const __half* sig1 = P.a.xsplit + cd1.sig_off + static_cast<long long>(it.y) * TSPAN;   // DUAL only
        {
            const int b = it.z + brank;
            const uint8_t* ablk = (item_x8(P, it.x) ? P.a.Aimg8 : P.a.Aimg) +
                                  static_cast<size_t>(b) * P.nchunks * STAGE_BYTES;
            for (int g = 0; g < P.nseg; ++g) {
                const Seg sgm = P.seg[g];
                const uint32_t bytes = (TSPAN + sgm.ntaps) * 2;
                mbar_wait(&S.sigempty[sg.idx], sg.phase ^ 1);
                if (ISSUE_LANE) {
                    mbar_arrive_expect_tx(&S.sigfull[sg.idx], (hi_only && !DUAL) ? bytes : 2 * bytes);
                    const __half* src = sig0 + static_cast<long long>(sgm.chan * 2) * cd.Lpad + sgm.tap0;
                    SIG_G2S(S.sig + sg.idx * SIG_BUF_BYTES, src, bytes, &S.sigfull[sg.idx]);
                    if (DUAL)        // the second chunk's hi plane, in the lo-plane slot
                        SIG_G2S(S.sig + sg.idx * SIG_BUF_BYTES + SIG_HALFS * 2,
                                 sig1 + static_cast<long long>(sgm.chan * 2) * cd1.Lpad + sgm.tap0, bytes,
                                 &S.sigfull[sg.idx]);
                    else if (!hi_only)
                        SIG_G2S(S.sig + sg.idx * SIG_BUF_BYTES + SIG_HALFS * 2, src + cd.Lpad, bytes,
                                 &S.sigfull[sg.idx]);
                }
                ISSUE_SYNC();
                sg.advance();
                const int nck = sgm.ntaps / CHUNK_TAPS;
                for (int kc = 0; kc < nck; ++kc) {
                    mbar_wait(&S.empty[st.idx], st.phase ^ 1);
                    if (ISSUE_LANE) {
                        // hi-only screening: the A_hi tile is the first half of the stage image
                        mbar_arrive_expect_tx(&S.full[st.idx], a_bytes);
                        A_G2S(S.stage + st.idx * STAGE_BYTES,
                                 ablk + static_cast<size_t>(sgm.chunk0 + kc) * STAGE_BYTES, a_bytes,
                                 &S.full[st.idx]);
                    }
                    ISSUE_SYNC();
                    st.advance();
                }
            }
        }
    }
}

// ---------------------------------------------- MMA issuer (warp 1, an elected lane issues)
template <int NQ, bool CG2, bool DUAL>
__device__ __forceinline__ void mma_loop(const K1Params& P, const Smem& S, uint32_t tmem) {
    const uint32_t idesc = idesc_f16_f32(CG2 ? 256 : 128, DUAL ? NQ / 2 : NQ);
    const uint32_t idesc8 = idesc_e4m3_e5m2_f32(CG2 ? 256 : 128, NQ);
    const uint64_t a_base = smem_desc_kmajor_noswz(0, A_LBO, A_SBO);
    const uint64_t b_base = smem_desc_kmajor_noswz(0, B_LBO, B_SBO);
    const uint32_t stage0 = smem_u32(S.stage), sig0 = smem_u32(S.sig);
    Ring st(STAGES), sg(2), ac(2);
    const bool hi_only = P.a.hi_only != 0;
    int x8_next = item_first<CG2>() < P.a.nitems ? item_x8(P, P.a.items[item_first<CG2>()].x) : 0;
    for (int item = item_first<CG2>(); item < P.a.nitems; item += item_step<CG2>()) {
        const bool x8 = x8_next != 0;
        if (item + item_step<CG2>() < P.a.nitems)   // fetched while this item's MMAs issue
            x8_next = item_x8(P, P.a.items[item + item_step<CG2>()].x);
        const int kblk = x8 ? P.a.kblk8 : P.a.kblk;
        {
            int cib = 0, done = 0, nacc = 0;
            for (int g = 0; g < P.nseg; ++g) {
                const int nck = P.seg[g].ntaps / CHUNK_TAPS;
                mbar_wait(&S.sigfull[sg.idx], sg.phase);
                if (CG2) mbar_wait_cluster(&S.psigfull[sg.idx], sg.phase);   // the peer's half of the B rows
                tc_fence_after();
                const uint32_t sh = sig0 + sg.idx * SIG_BUF_BYTES;
                const uint32_t sl = sh + SIG_HALFS * 2;
                for (int kc = 0; kc < nck; ++kc) {
                    mbar_wait(&S.full[st.idx], st.phase);
                    if (CG2) mbar_wait_cluster(&S.pfull[st.idx], st.phase);      // the peer's A stage
                    if (cib == 0) {
                        if (CG2) mbar_wait_cluster(&S.accempty[ac.idx], ac.phase ^ 1);   // both CTAs' drain warps
                        else mbar_wait(&S.accempty[ac.idx], ac.phase ^ 1);
                    }
                    tc_fence_after();
                    const uint32_t d = tmem + ac.idx * 256;
                    const uint32_t ah = stage0 + st.idx * STAGE_BYTES;
                    const uint32_t al = ah + TILE_BYTES;
                    const uint32_t bo = kc * (CHUNK_TAPS * 2);
                    const bool last = (cib + 1 == acc_stages(nacc, kblk, P.nchunks)) || (done + 1 == P.nchunks);
                    if (ISSUE_LANE) {
                        if (CG2) {
                            // one MMA of M = 256 for the pair: this CTA's and the peer's basis block against the
                            // 2048-lag tile whose B rows the two CTAs hold half each
                            if (x8) {
#pragma unroll
                                for (int kk = 0; kk < CHUNK_TAPS / 16; ++kk) {
                                    const uint64_t dah = a_base | ((ah + kk * 256) >> 4);
                                    const uint64_t dal = a_base | ((al + kk * 256) >> 4);
                                    const uint64_t dbh = b_base | ((sh + bo + kk * 32) >> 4);
                                    const uint64_t dbl = b_base | ((sl + bo + kk * 32) >> 4);
                                    umma2_f16(d, dah, dbh, idesc, (cib | kk) ? 1u : 0u);
                                    umma2_f8(d, dal, dbl, idesc8, 1u);
                                }
                            } else {
#pragma unroll
                                for (int kk = 0; kk < CHUNK_TAPS / 16; ++kk) {
                                    const uint64_t dah = a_base | ((ah + kk * 256) >> 4);
                                    const uint64_t dal = a_base | ((al + kk * 256) >> 4);
                                    const uint64_t dbh = b_base | ((sh + bo + kk * 32) >> 4);
                                    const uint64_t dbl = b_base | ((sl + bo + kk * 32) >> 4);
                                    umma2_f16_a_fill(d, dah, dbh, idesc, (cib | kk) ? 1u : 0u);
                                    umma2_f16_a_lastuse(d, dah, dbl, idesc, 1u);
                                    umma2_f16(d, dal, dbh, idesc, 1u);
                                }
                            }
                            umma2_commit_mc(&S.empty[st.idx], 0x3);
                            if (last) umma2_commit_mc(&S.accfull[ac.idx], 0x3);
                            if (kc == nck - 1) umma2_commit_mc(&S.sigempty[sg.idx], 0x3);
                        } else if (DUAL) {
                            // two chunks, one A tile: the second MMA takes it from the operand collector
#pragma unroll
                            for (int kk = 0; kk < CHUNK_TAPS / 16; ++kk) {
                                const uint64_t dah = a_base | ((ah + kk * 256) >> 4);
                                const uint64_t db0 = b_base | ((sh + bo + kk * 32) >> 4);
                                const uint64_t db1 = b_base | ((sl + bo + kk * 32) >> 4);
                                umma_f16_a_fill(d, dah, db0, idesc, (cib | kk) ? 1u : 0u);
                                umma_f16_a_lastuse(d + NQ / 2, dah, db1, idesc, (cib | kk) ? 1u : 0u);
                            }
                        } else if (hi_only) {
#pragma unroll
                            for (int kk = 0; kk < CHUNK_TAPS / 16; ++kk) {
                                const uint64_t dah = a_base | ((ah + kk * 256) >> 4);
                                const uint64_t dbh = b_base | ((sh + bo + kk * 32) >> 4);
                                umma_f16(d, dah, dbh, idesc, (cib | kk) ? 1u : 0u);
                            }
                        } else if (x8) {
                            // both cross terms in ONE 8-bit MMA: the "lo" tiles hold byte pairs
                            // per tap, (u_lo, u_hi) in e4m3 against (x_hi, x_lo) in e5m2
#pragma unroll
                            for (int kk = 0; kk < CHUNK_TAPS / 16; ++kk) {
                                const uint64_t dah = a_base | ((ah + kk * 256) >> 4);
                                const uint64_t dal = a_base | ((al + kk * 256) >> 4);
                                const uint64_t dbh = b_base | ((sh + bo + kk * 32) >> 4);
                                const uint64_t dbl = b_base | ((sl + bo + kk * 32) >> 4);
                                umma_f16(d, dah, dbh, idesc, (cib | kk) ? 1u : 0u);
                                umma_f8(d, dal, dbl, idesc8, 1u);
                            }
                        } else {
#pragma unroll
                            for (int kk = 0; kk < CHUNK_TAPS / 16; ++kk) {
                                const uint64_t dah = a_base | ((ah + kk * 256) >> 4);
                                const uint64_t dal = a_base | ((al + kk * 256) >> 4);
                                const uint64_t dbh = b_base | ((sh + bo + kk * 32) >> 4);
                                const uint64_t dbl = b_base | ((sl + bo + kk * 32) >> 4);
#if DTX_COLLECTOR_A
                                // hi*hi and hi*lo share the A tile: the second takes it from the operand
                                // collector instead of reading its 4 KB from shared memory again
                                umma_f16_a_fill(d, dah, dbh, idesc, (cib | kk) ? 1u : 0u);
                                umma_f16_a_lastuse(d, dah, dbl, idesc, 1u);
#else
                                umma_f16(d, dah, dbh, idesc, (cib | kk) ? 1u : 0u);
                                umma_f16(d, dah, dbl, idesc, 1u);
#endif
                                umma_f16(d, dal, dbh, idesc, 1u);
                            }
                        }
                        if (!CG2) {
                            umma_commit(&S.empty[st.idx]);
                            if (last) umma_commit(&S.accfull[ac.idx]);
                            if (kc == nck - 1) umma_commit(&S.sigempty[sg.idx]);
                        }
                    }
                    ISSUE_SYNC();
                    st.advance();
                    ++cib;
                    ++done;
                    if (last) {
                        ac.advance();
                        cib = 0;
                        ++nacc;
                    }
                }
                sg.advance();
            }
        }
    }
}

// --------------------------------------- CTA pairs: the peer CTA's forwarder (warp 2 of cluster rank 1)
// Only the leader issues MMAs, so it has to learn when the PEER's signal spans and A stages have landed:
// this warp follows the peer's own full barriers and arrives on the leader's copies.
template <int NQ>
__device__ __forceinline__ void forward_loop(const K1Params& P, const Smem& S) {
    Ring st(STAGES), sg(2);
    for (int item = item_first<true>(); item < P.a.nitems; item += item_step<true>()) {
        for (int g = 0; g < P.nseg; ++g) {
            mbar_wait(&S.sigfull[sg.idx], sg.phase);
            mbar_arrive_leader(&S.psigfull[sg.idx]);
            sg.advance();
            const int nck = P.seg[g].ntaps / CHUNK_TAPS;
            for (int kc = 0; kc < nck; ++kc) {
                mbar_wait(&S.full[st.idx], st.phase);
                mbar_arrive_leader(&S.pfull[st.idx]);
                st.advance();
            }
        }
    }
}

// ------------------------------------------------- drain + epilogue (8 warps, 256 threads)
// DS rows are written with the streaming (evict-first) hint so that the 18 GB DS stream of a launch
// does not evict the L2-resident inputs; -DDTX_DS_STREAM=0 uses plain stores.
#ifndef DTX_DS_STREAM
#define DTX_DS_STREAM 1
#endif
__device__ __forceinline__ void store_ds_row(float* dst, float4 v) {
#ifdef DTX_NO_DS_STORE   // traffic experiment only: results are NOT written (except NaN, never true here)
    if (v.x == 1.2345e30f) *reinterpret_cast<float4*>(dst) = v;
    return;
#endif
#if DTX_DS_STREAM
    __stcs(reinterpret_cast<float4*>(dst), v);
#else
    *reinterpret_cast<float4*>(dst) = v;
#endif
}

// Fused read-out, per-value path (MODE 2).  Kept OUT of line on purpose: the read-out loop is unrolled 8 x 4
// times (register-resident accumulators), and with this body inlined at every site the kernel grew to 31 k
// SASS instructions -- the drain warps then streamed ~400 KB of code per work item through the instruction
// cache and the fused mode ran 9-14 % slower than writing DS (profiles/r02_fused_epilogue.md).
struct FusedState {
    float mx;
    int cur, cnt, zero;
    float s1, s2, s3, s4;
    int n;
};
__device__ __noinline__ FusedState fused_values(FusedState st, float4 acc, float4 ie, int t0, int t_lo, int t_hi,
                                                float thr, int out_row, int row, const K1Args* a, float flo,
                                                float finv, double step) {
    {
        // four values at once, branch free (the quad path of k3_fast_kernel): all four core lags, finite
        // energy, well inside the bin of the current run, below the threshold
        const float toff = -flo * finv;
        const float curf = static_cast<float>(st.cur);
        const float q0 = fmaf(acc.x, finv, toff), q1 = fmaf(acc.y, finv, toff), q2 = fmaf(acc.z, finv, toff),
                    q3 = fmaf(acc.w, finv, toff);
        const float g0 = floorf(q0), g1 = floorf(q1), g2 = floorf(q2), g3 = floorf(q3);
        const bool inside = fabsf(q0 - g0 - 0.5f) < 0.499f && fabsf(q1 - g1 - 0.5f) < 0.499f &&
                            fabsf(q2 - g2 - 0.5f) < 0.499f && fabsf(q3 - g3 - 0.5f) < 0.499f;
        const bool same = st.cur >= 0 && g0 == curf && g1 == curf && g2 == curf && g3 == curf;
        const float m4 = fmaxf(fmaxf(acc.x, acc.y), fmaxf(acc.z, acc.w));
        const float e4 = fmaxf(fmaxf(ie.x, ie.y), fmaxf(ie.z, ie.w));
        if (inside && same && m4 < thr && e4 < INFINITY && t0 >= t_lo && t0 + 3 < t_hi && !a->fas) {
            st.cnt += 4;
            st.mx = fmaxf(st.mx, m4);
            return st;
        }
    }
    const float vv[4] = {acc.x, acc.y, acc.z, acc.w};
    const float iv[4] = {ie.x, ie.y, ie.z, ie.w};
#pragma unroll 1
    for (int e = 0; e < 4; ++e) {
        const int t = t0 + e;
        if (t < t_lo || t >= t_hi) continue;
        float v = vv[e];
        if (isinf(iv[e])) {      // zero-energy window: the reference's inf, zeroed (detect.py:275-281)
            v = 0.f;
            st.zero = 1;
        }
        st.mx = fmaxf(st.mx, v);
        const int bin = hist_bin_fast(v, flo, finv, a->hist_lo, a->hist_hi, step, a->nbins);
        if (bin == st.cur) ++st.cnt;
        else {
            if (st.cnt > 0 && st.cur >= 0)
                atomicAdd(&a->hist[static_cast<long long>(out_row) * HIST_MAX_BINS + st.cur],
                          static_cast<unsigned long long>(st.cnt));
            st.cur = bin;
            st.cnt = 1;
        }
        if (v >= thr) {
            const int q = atomicAdd(a->ncand, 1);
            if (q < a->cand_cap) {
                Candidate cnd;
                cnd.row = row;
                cnd.t = t; cnd.ds = v; cnd.lta = 0.f;
                a->cand[q] = cnd;
            }
        }
        if (a->fas) {
            st.s1 += v;
            st.s2 = fmaf(v, v, st.s2);
            st.s3 += logf(fmaxf(v, 1e-30f));
            st.s4 += log1pf(-fminf(v, 0.99999994f));
            ++st.n;
        }
    }
    return st;
}

template <int NQ, int MODE, bool CG2, bool DUAL>
__device__ __forceinline__ void drain_loop(const K1Params& P, const Smem& S, uint32_t tmem, int warp,
                                           int lane) {
    const int brank = CG2 ? cluster_ctarank() : 0;
    constexpr int TT = 8 * NQ;
    constexpr int NCOL = NQ / 2;                 // accumulator columns per thread
    const int dw = warp - FIRST_DRAIN_WARP;      // 0..7
    const int lq = warp & 3;                     // TMEM lane quarter this warp may read
    const int colhalf = dw >> 2;                 // which half of the NQ accumulator columns
    const int p = 2 * lq + (lane & 1);           // phase of this thread's row
    const int kl = lane >> 1;                    // basis-vector slot of this thread's row
    const uint32_t taddr0 = tmem + (static_cast<uint32_t>(lq * 32) << 16) + colhalf * NCOL;
    Ring ac(2), nm(2);
#if DTX_FADD2
    unsigned long long sums2[NCOL / 2];
#define SUM_AT(idx) f32x2_get(sums2[(idx) >> 1], (idx) & 1)
#else
    float sums[NCOL];
#define SUM_AT(idx) sums[idx]
#endif
    const double f_step = MODE == 2 ? (P.a.hist_hi - P.a.hist_lo) / P.a.nbins : 0.0;
    const float f_flo = static_cast<float>(P.a.hist_lo);
    const float f_finv = MODE == 2 ? static_cast<float>(P.a.nbins / (P.a.hist_hi - P.a.hist_lo)) : 0.f;
    const float f_toff = -f_flo * f_finv;
    for (int item = item_first<CG2>(); item < P.a.nitems; item += item_step<CG2>()) {
        const int4 it = P.a.items[item];
        const int mychunk = (DUAL && colhalf) ? it.w : it.x;   // DUAL: the column halves belong to two chunks
        const ChunkDesc cd = P.a.chunks[mychunk];
        const int kblk = item_x8(P, it.x) ? P.a.kblk8 : P.a.kblk;
        int ndrains = 0;
        for (int rem = P.nchunks; rem > 0; ++ndrains) rem -= acc_stages(ndrains, kblk, P.nchunks);
        const float sc = P.a.chunk_scale[mychunk] * P.u_inv_scale;
        const float* smu = reinterpret_cast<const float*>(S.norm + nm.idx * NORM_BUF_BYTES);
        const float* sie = smu + TILE_T;
        {
            const int b = it.z + brank;
#if DTX_FADD2
#pragma unroll
            for (int i = 0; i < NCOL / 2; ++i) sums2[i] = 0ull;
#else
#pragma unroll
            for (int i = 0; i < NCOL; ++i) sums[i] = 0.f;
#endif
            for (int dr = 0; dr < ndrains; ++dr) {
                mbar_wait(&S.accfull[ac.idx], ac.phase);
                tc_fence_after();
                const uint32_t ta = taddr0 + ac.idx * 256;
#if DTX_DRAIN_PIPE
                // two 32-column loads in flight: the additions of one overlap the TMEM read of the next
                uint32_t va[32], vb[32];
                tmem_ld_x32(ta, va);
                tmem_wait_ld();
#pragma unroll
                for (int i = 0; i < NCOL / 32; ++i) {
                    uint32_t(&v)[32] = (i & 1) ? vb : va;
                    uint32_t(&vn)[32] = (i & 1) ? va : vb;
                    if (i + 1 < NCOL / 32) tmem_ld_x32(ta + (i + 1) * 32, vn);
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        sums2[i * 16 + j] = f32x2_add(sums2[i * 16 + j], f32x2_pack(v[2 * j], v[2 * j + 1]));
                    if (i + 1 < NCOL / 32) tmem_wait_ld();
                    if (i == NCOL / 32 - 2) {     // the last load has landed: the accumulator is free
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&S.accempty[ac.idx]);
                    }
                }
#else
#pragma unroll
                for (int i = 0; i < NCOL / 32; ++i) {
                    uint32_t v[32];
                    tmem_ld_x32(ta + i * 32, v);
                    tmem_wait_ld();
                    if (i == NCOL / 32 - 1) {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) {
                            if (CG2) mbar_arrive_leader(&S.accempty[ac.idx]);   // the leader waits for both CTAs
                            else mbar_arrive(&S.accempty[ac.idx]);
                        }
                    }
#if DTX_FADD2
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        sums2[i * 16 + j] = f32x2_add(sums2[i * 16 + j], f32x2_pack(v[2 * j], v[2 * j + 1]));
#else
#pragma unroll
                    for (int j = 0; j < 32; ++j) sums[i * 32 + j] += __uint_as_float(v[j]);
#endif
                }
#endif
                ac.advance();
            }
            // ---------------------------------------------------- K2 epilogue
            // Every thread turns its accumulators into normalised projections c (squared for the
            // detection statistic) and parks them in smem as [slot][lag]; the four warps of this
            // column half then sum the slots of each subspace and write 512-byte DS rows.
            mbar_wait(&S.normfull[nm.idx], nm.phase);
            const float nsumU = -P.a.binfo[b * VEC_PER_BLOCK + kl].sumU;
            float* ebuf = reinterpret_cast<float*>(S.epi + colhalf * EPI_BUF_BYTES);
            float* wr = ebuf + kl * EPI_STRIDE + (p ^ (kl >= 8 ? 2 : 0));
            // mu tile is phase-major per 1024-lag group (k0_norm): this thread's 128 columns are contiguous
            const float* pmu = smu + (8 * NCOL * colhalf / 1024) * 1024 + p * 128 + (NCOL * colhalf) % 128;
            const float* pie4 = sie + 8 * NCOL * colhalf + lane * 4;
            float* dsbase = DUAL ? P.a.DS + cd.ds_off + static_cast<long long>(it.y) * (TT / 2) + lane * 4
                                 : P.a.DS + cd.ds_off + static_cast<long long>(it.y) * TT + 8 * NCOL * colhalf + lane * 4;
            // read-out role of this warp: the subspaces whose first slot is lq, lq+4, lq+8, lq+12
            BlockInfo hb[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) hb[k] = P.a.binfo[b * VEC_PER_BLOCK + lq + 4 * k];
            // MODE 2 (fused K3): per-role running maximum, histogram run, FAS partial sums of this item
            float f_thr[4], f_max[4], f_s1[4], f_s2[4], f_s3[4], f_s4[4];
            int f_cur[4], f_cnt[4], f_n[4];
            unsigned f_zero = 0;
            const bool f_skip = MODE == 2 && P.a.chunk_bad[it.x] != 0;   // non-finite samples: rows are NaN, nothing counted
            if (MODE == 2) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const bool on = hb[k].nrows != 0 && hb[k].out_row >= 0;
                    f_thr[k] = on ? P.a.thr[hb[k].out_row] : INFINITY;
                    f_max[k] = 0.f; f_cur[k] = -1; f_cnt[k] = 0; f_n[k] = 0;
                    f_s1[k] = f_s2[k] = f_s3[k] = f_s4[k] = 0.f;
                }
            }
#pragma unroll
            for (int c = 0; c < NCOL / EPI_QC; ++c) {
#pragma unroll
                for (int j4 = 0; j4 < EPI_QC / 4; ++j4) {
                    const float4 m4 = *reinterpret_cast<const float4*>(pmu + c * EPI_QC + 4 * j4);
                    const float mm[4] = {m4.x, m4.y, m4.z, m4.w};
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) {
                        const int j = 4 * j4 + jj;
                        const float cc = fmaf(SUM_AT(c * EPI_QC + j), sc, mm[jj] * nsumU);
                        wr[8 * j] = MODE != 1 ? cc * cc : cc;
                    }
                }
                named_bar_sync(1 + colhalf, 128);
                float4 ie = *reinterpret_cast<const float4*>(pie4 + c * EPI_LAGS);
                if (MODE == 1) {
                    // signed Pearson coefficient: templates are pre-scaled by 1/||x1 - mean||, so
                    // res = c / sqrt(E) = c * sqrt(invE * n/(n-1))
                    ie.x = sqrtf(ie.x * P.inv_cn); ie.y = sqrtf(ie.y * P.inv_cn);
                    ie.z = sqrtf(ie.z * P.inv_cn); ie.w = sqrtf(ie.w * P.inv_cn);
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int slot = lq + 4 * k;
                    if (hb[k].nrows == 0 || hb[k].out_row < 0) continue;   // not the first slot of a subspace
                    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
                    for (int sl = slot; sl < hb[k].seg_end; ++sl) {
                        const float4 v = *reinterpret_cast<const float4*>(ebuf + sl * EPI_STRIDE + lane * 4);
                        if (sl >= 8) { acc.x += v.z; acc.y += v.w; acc.z += v.x; acc.w += v.y; }
                        else { acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w; }
                    }
                    acc.x *= ie.x; acc.y *= ie.y; acc.z *= ie.z; acc.w *= ie.w;
                    if (MODE == 2) {
                        // the statistic of 4 consecutive lags of one subspace, never written: maximum,
                        // histogram (run-length: noise sits in one bin), candidates, FAS sums -- exactly
                        // what k3_fast_kernel does with the stored row
                        if (!f_skip) {
                            const int t0 = it.y * TT + 8 * NCOL * colhalf + c * EPI_LAGS + lane * 4;
                            {
                                FusedState stt{f_max[k], f_cur[k], f_cnt[k], 0, f_s1[k], f_s2[k], f_s3[k], f_s4[k], f_n[k]};
                                stt = fused_values(stt, acc, ie, t0, cd.t_lo, cd.t_hi, f_thr[k], hb[k].out_row,
                                                   P.a.row_base + it.x * P.a.S + hb[k].out_row, &P.a, f_flo, f_finv, f_step);
                                f_max[k] = stt.mx; f_cur[k] = stt.cur; f_cnt[k] = stt.cnt;
                                f_zero |= static_cast<unsigned>(stt.zero) << k;
                                f_s1[k] = stt.s1; f_s2[k] = stt.s2; f_s3[k] = stt.s3; f_s4[k] = stt.s4; f_n[k] = stt.n;
                            }
                        }
                    } else {
                        float* dst = dsbase + static_cast<long long>(hb[k].out_row) * cd.Tpad + c * EPI_LAGS;
                        // (pieces of a rank > 16 subspace have rows of their own: launch_sum_pieces adds them
                        // up in a fixed order afterwards, so DS does not depend on CTA timing)
                        store_ds_row(dst, acc);
                    }
                }
                named_bar_sync(1 + colhalf, 128);
            }
            if (MODE == 2 && !f_skip) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (hb[k].nrows == 0 || hb[k].out_row < 0) continue;   // warp-uniform
                    const int orow = hb[k].out_row;
                    const int row = P.a.row_base + it.x * P.a.S + orow;
                    float m = f_max[k];
                    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
                    if (lane == 0 && m > 0.f) atomicMax(&P.a.rowmax_bits[row], __float_as_uint(m));
                    if (__any_sync(0xffffffffu, (f_zero >> k) & 1u) && lane == 0) atomicOr(&P.a.rowflags[row], 2);
                    // histogram runs: one atomic per distinct bin of the warp
                    const int cur = f_cnt[k] > 0 ? f_cur[k] : -1;
                    const unsigned grp = __match_any_sync(0xffffffffu, cur);
                    const int tot = __reduce_add_sync(grp, f_cnt[k]);
                    if (cur >= 0 && lane == __ffs(grp) - 1)
                        atomicAdd(&P.a.hist[static_cast<long long>(orow) * HIST_MAX_BINS + cur],
                                  static_cast<unsigned long long>(tot));
                    if (P.a.fas) {
                        double d1 = f_s1[k], d2 = f_s2[k], d3 = f_s3[k], d4 = f_s4[k], dn = f_n[k];
                        for (int o = 16; o > 0; o >>= 1) {
                            d1 += __shfl_xor_sync(0xffffffffu, d1, o);
                            d2 += __shfl_xor_sync(0xffffffffu, d2, o);
                            d3 += __shfl_xor_sync(0xffffffffu, d3, o);
                            d4 += __shfl_xor_sync(0xffffffffu, d4, o);
                            dn += __shfl_xor_sync(0xffffffffu, dn, o);
                        }
                        if (lane == 0) {
                            double* f = P.a.fas + orow * 5;
                            atomicAdd(f, dn); atomicAdd(f + 1, d1); atomicAdd(f + 2, d2);
                            atomicAdd(f + 3, d3); atomicAdd(f + 4, d4);
                        }
                    }
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&S.normempty[nm.idx]);
        nm.advance();
    }
}

template <int NQ, int MODE, bool CG2, bool DUAL = false>
__global__ void __launch_bounds__(NTHREADS, 1) k1_kernel(const __grid_constant__ K1Params P) {
    static_assert(!DUAL || (MODE == 1 && NQ == 256 && !CG2), "DUAL: CCX screening series only");
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t full[STAGES], empty[STAGES], pfull[STAGES];
    __shared__ uint64_t sigfull[2], sigempty[2], accfull[2], accempty[2], normfull[2], normempty[2], psigfull[2];
    __shared__ uint32_t tmem_base_s;
    Smem S;
    S.stage = smem;
    S.sig = smem + STAGES * STAGE_BYTES;
    S.norm = S.sig + 2 * SIG_BUF_BYTES;
    S.epi = S.norm + 2 * NORM_BUF_BYTES;
    S.full = full; S.empty = empty; S.sigfull = sigfull; S.sigempty = sigempty;
    S.accfull = accfull; S.accempty = accempty; S.normfull = normfull; S.normempty = normempty;
    S.pfull = pfull; S.psigfull = psigfull;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
            mbar_init(&pfull[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&sigfull[s], 1);
            mbar_init(&sigempty[s], 1);
            mbar_init(&accfull[s], 1);
            mbar_init(&accempty[s], CG2 ? 2 * NDRAIN_WARPS : NDRAIN_WARPS);
            mbar_init(&psigfull[s], 1);
            mbar_init(&normfull[s], 1);
            mbar_init(&normempty[s], NDRAIN_WARPS);
        }
        fence_barrier_init();
    }
    if (warp == 0) {
        if (CG2) {
            tmem_alloc_cg2(&tmem_base_s, 512);
            tmem_relinquish_cg2();
        } else {
            tmem_alloc(&tmem_base_s, 512);
            tmem_relinquish();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (CG2) cluster_sync_all();   // both CTAs' barriers exist before anything arrives on them
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;

    // setmaxnreg must sit at the top of each role branch for ptxas to budget the branch
    if (warp < FIRST_DRAIN_WARP) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS_CTRL));
        if (warp == 0 && ROLE_LANES(lane)) {
            producer_loop<NQ, CG2, DUAL>(P, S);
        } else if (warp == 1 && ROLE_LANES(lane)) {
            if (!CG2 || cluster_ctarank() == 0) mma_loop<NQ, CG2, DUAL>(P, S, tmem);   // pairs: only the leader issues
        } else if (CG2 && warp == 2 && lane == 0) {
            if (cluster_ctarank() == 1) forward_loop<NQ>(P, S);
        }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS_DRAIN));
        drain_loop<NQ, MODE, CG2, DUAL>(P, S, tmem, warp, lane);
    }
    tc_fence_before();
    __syncthreads();
    if (CG2) cluster_sync_all();   // no CTA leaves while its peer may still signal it or read its shared memory
    if (warp == 0) {
        if (CG2) tmem_dealloc_cg2(tmem, 512);
        else tmem_dealloc(tmem, 512);
    }
}

// ---------------------------------------------------------------- basis image
// One thread per 16-byte unit (8 taps of one row of one hi/lo tile).
__global__ void __launch_bounds__(256)
basis_image_kernel(const double* __restrict__ U, const int* __restrict__ slot_row,
                   const __grid_constant__ BasisLayout lay, uint8_t* __restrict__ Aimg, int x8) {
    const long long units_per_tile = 128 * 8;
    const long long total = static_cast<long long>(lay.nblocks) * lay.nchunks * 2 * units_per_tile;
    const long long gid = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (gid >= total) return;
    long long r_ = gid;
    const int unit = static_cast<int>(r_ % units_per_tile);
    r_ /= units_per_tile;
    const int hl = static_cast<int>(r_ % 2);
    r_ /= 2;
    const int chunk = static_cast<int>(r_ % lay.nchunks);
    const int b = static_cast<int>(r_ / lay.nchunks);
    // unit -> (row group, k-core, row in group) in image order
    const int rin = unit % 8, kc = (unit / 8) % 8, rg = unit / 64;
    const int r = rg * 8 + rin;
    const int w = r / 32, kslot = (r % 32) / 2, e = r & 1, ph = 2 * w + e;
    // find the segment of this chunk
    int g = 0;
    while (g + 1 < lay.nseg && lay.seg[g + 1].chunk0 <= chunk) ++g;
    const Seg sg = lay.seg[g];
    const int tap_base = sg.tap0 + (chunk - sg.chunk0) * CHUNK_TAPS + kc * 8;
    const int row = slot_row[b * VEC_PER_BLOCK + kslot];
    __align__(16) __half out[8];
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) {
        const int j = tap_base + jj - ph;
        float v = 0.f;
        if (row >= 0 && j >= 0 && j < lay.ns)
            v = static_cast<float>(
                ldexp(U[static_cast<long long>(row) * lay.n + static_cast<long long>(j) * lay.Nc + sg.chan],
                      lay.u_exp));
        const __half hi = __float2half_rn(v);
        if (hl == 0) out[jj] = hi;
        else if (!x8) out[jj] = __float2half_rn(v - __half2float(hi));
        else {
            // 8-bit cross-term image: byte 0 = e4m3(u_lo * 2^X8_S), byte 1 = e4m3(u_hi * 2^-X8_S)
            const float fh = __half2float(hi);
            const unsigned b0 = __nv_cvt_float_to_fp8(ldexpf(v - fh, X8_SHIFT), __NV_SATFINITE, __NV_E4M3);
            const unsigned b1 = __nv_cvt_float_to_fp8(ldexpf(fh, -X8_SHIFT), __NV_SATFINITE, __NV_E4M3);
            out[jj] = __ushort_as_half(static_cast<unsigned short>(b0 | (b1 << 8)));
        }
    }
    uint8_t* dst = Aimg + (static_cast<size_t>(b) * lay.nchunks + chunk) * STAGE_BYTES +
                   static_cast<size_t>(hl) * TILE_BYTES + rg * 1024 + kc * 128 + rin * 16;
    *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(out);
}

template <int NQ, int MODE>
void launch_k1_t(const K1Params& P, int grid, cudaStream_t st) {
    if (P.a.cg2) {
        // CTA pairs: clusters of 2 (same TPC), one work item per cluster
        cudaFuncSetAttribute(k1_kernel<NQ, MODE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
        cudaLaunchConfig_t cfg = {};
        const int nclusters = std::max(1, std::min(P.a.nitems, P.a.num_sms / 2));
        cfg.gridDim = dim3(2 * nclusters);
        cfg.blockDim = dim3(NTHREADS);
        cfg.dynamicSmemBytes = SMEM_BYTES;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        cudaLaunchKernelEx(&cfg, k1_kernel<NQ, MODE, true>, P);
        return;
    }
    cudaFuncSetAttribute(k1_kernel<NQ, MODE, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    k1_kernel<NQ, MODE, false><<<grid, NTHREADS, SMEM_BYTES, st>>>(P);
}

}  // namespace

// Per basis vector (row of U): sum, max |u|, sum u^2, sum u^4 in float64 -- what dtx_set_bases needs
// for the mu * sumU correction, the fp16 scale exponent and the precision policy.  One block per row.
__global__ void __launch_bounds__(256)
basis_row_stats_kernel(const double* __restrict__ U, int n, double* __restrict__ out) {
    const double* u = U + static_cast<long long>(blockIdx.x) * n;
    double s1 = 0, s2 = 0, s4 = 0, mx = 0;
    for (int j = threadIdx.x; j < n; j += 256) {
        const double v = u[j];
        s1 += v;
        s2 += v * v;
        s4 += (v * v) * (v * v);
        mx = fmax(mx, fabs(v));   // NaN-propagating max is not needed: a NaN shows up in s1
    }
    __shared__ double sh[4][8];
    for (int o = 16; o > 0; o >>= 1) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        s4 += __shfl_xor_sync(0xffffffffu, s4, o);
        mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) { sh[0][w] = s1; sh[1][w] = s2; sh[2][w] = s4; sh[3][w] = mx; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < 8; ++i) { s1 += sh[0][i]; s2 += sh[1][i]; s4 += sh[2][i]; mx = fmax(mx, sh[3][i]); }
        double* o = out + static_cast<long long>(blockIdx.x) * 4;
        o[0] = s1; o[1] = mx; o[2] = s2; o[3] = s4;
    }
}

int k1_smem_bytes() { return SMEM_BYTES; }

void launch_basis_image(const double* d_U, const int* d_slot_row, const BasisLayout& lay,
                        uint8_t* d_Aimg, int x8, cudaStream_t st) {
    const long long total = static_cast<long long>(lay.nblocks) * lay.nchunks * 2 * 128 * 8;
    const int grid = static_cast<int>((total + 255) / 256);
    basis_image_kernel<<<grid, 256, 0, st>>>(d_U, d_slot_row, lay, d_Aimg, x8);
}

void launch_basis_row_stats(const double* d_U, int R, int n, double* d_out, cudaStream_t st) {
    basis_row_stats_kernel<<<R, 256, 0, st>>>(d_U, n, d_out);
}

void launch_k1(const K1Args& a, const BasisLayout& lay, cudaStream_t st) {
    K1Params P;
    P.a = a;
    P.nblocks = lay.nblocks;
    P.nchunks = lay.nchunks;
    P.nseg = lay.nseg;
    P.u_inv_scale = lay.u_inv_scale;
    P.inv_cn = static_cast<float>(static_cast<double>(lay.n) / (lay.n - 1.0));
    for (int i = 0; i < lay.nseg; ++i) P.seg[i] = lay.seg[i];
    const int grid = a.nitems < a.num_sms ? a.nitems : a.num_sms;
    if (grid < 1) return;
    if (a.dual) {
        cudaFuncSetAttribute(k1_kernel<256, 1, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
        k1_kernel<256, 1, false, true><<<grid, NTHREADS, SMEM_BYTES, st>>>(P);
    } else if (a.nq == 128) {
        if (a.mode == 1) launch_k1_t<128, 1>(P, grid, st);
        else if (a.fused) launch_k1_t<128, 2>(P, grid, st);
        else launch_k1_t<128, 0>(P, grid, st);
    } else {
        if (a.mode == 1) launch_k1_t<256, 1>(P, grid, st);
        else if (a.fused) launch_k1_t<256, 2>(P, grid, st);
        else launch_k1_t<256, 0>(P, grid, st);
    }
}

}  // namespace dtx
