// dtx_kernels.cuh -- device-side data structures and kernel launchers of detex_b200.
//
// Data layout in HBM (one "batch" = the chunks of one dtx_run_* call):
//   raw      : the caller's multiplexed chunks, f64 or f32, back to back
//   xsplit   : per chunk [Nc][2][Lpad] fp16 -- de-multiplexed, centred, power-of-two
//              scaled signal split into hi + lo halves (x*2^ex ~= hi + lo), zero padded
//   mu/invE  : per chunk [Tpad] fp32 -- window mean and ((n-1)/n)/||w - mean||^2
//   DS       : per chunk [S][Tpad] fp32 -- detection statistic, row per subspace
//   Aimg     : per basis set [nblocks][nchunks][2][16 KB] fp16 -- the 8 phase-shifted
//              copies of 16 basis vectors per block, pre-tiled in the exact
//              shared-memory image the MMA descriptor reads (see k1_project.cu)
#pragma once
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_fp8.h>
#include <cuda_runtime.h>

namespace dtx {

constexpr int TILE_T = 2048;        // output lags per K1 tile (256 q-rows x 8 phases)
constexpr int CHUNK_TAPS = 64;      // taps per A stage
constexpr int VEC_PER_BLOCK = 16;   // basis vectors per K1 basis block (x 8 phases = 128 rows)
constexpr int MAX_SEG_TAPS = 3072;  // taps per K segment (bounded by the smem signal span)
constexpr int MAX_SEGS = 48;
constexpr int HIST_BINS = 400;        // default: np.linspace(lo, hi, 401) (detect.py:80, fas.py:31)
constexpr int HIST_MAX_BINS = 1024;   // row pitch of the device histograms; numBins - 1 <= this
// 8-bit cross-term engine: u_lo * 2^6 and u_hi * 2^-6 fit e4m3 (max|u * 2^eu| < 2^14), x_hi * 2^-6 and
// x_lo * 2^6 fit e5m2 (max|x * 2^ex| < 2^15); the shifts cancel in each product.
constexpr int X8_SHIFT = 6;
enum { X8_OFF = 0, X8_AUTO = 1, X8_FORCE = 2,     // per-run policy handed to k0_split
       X8_RATIO = 3 };   // no 8-bit mode; k0_norm records max_t sum(x - chunk mean)^2 / window energy per chunk

struct ChunkDesc {
    long long raw_off;   // element offset of the chunk in the raw buffer
    long long sig_off;   // fp16-element offset in xsplit
    long long norm_off;  // element offset in mu / invE
    long long ds_off;    // element offset in DS
    int L;               // multiplexed samples used (multiple of Nc)
    int Ls;              // samples per channel
    int T;               // number of output lags  = Ls - ns + 1
    int ntiles;          // ceil(T / TILE_T)
    int Lpad;            // padded per-channel length of the split arrays
    int Tpad;            // ntiles * TILE_T
    int blk_lo, blk_hi;  // basis blocks K1 runs for this chunk: [blk_lo, blk_hi)
    int t_lo, t_hi;      // core lags [t_lo, t_hi) that K3 counts (max, histogram, candidates, FAS sums); the
                         // lags outside are a halo that only the LTA windows read (time-segment sharding)
};

struct Seg {
    int chan;    // channel index
    int tap0;    // first (phase-extended) tap of the segment
    int ntaps;   // multiple of CHUNK_TAPS, <= MAX_SEG_TAPS
    int chunk0;  // index of the segment's first 64-tap chunk inside a block image
};

struct BlockInfo {   // one per (basis block, vector slot)
    float sumU;      // sum of the basis vector's entries (true units)
    int out_row;     // DS row this vector's piece is written to, -1 = padding: the subspace's own row,
                     // or for the 2nd, 3rd, .. 16-vector piece of a rank > 16 subspace a scratch row >= S
                     // (launch_sum_pieces adds those to the subspace row afterwards)
    int nrows;       // rank of the piece if this slot is its first vector, else 0
    int seg_end;     // slot index one past the last vector of this slot's subspace
};

struct BasisLayout {
    int Nc, ns, n;            // channels, taps per channel, multiplexed length
    int nseg;                 // K segments
    int nchunks;              // 64-tap chunks per block image
    int nblocks;              // basis blocks
    int S;                    // subspaces (DS rows)
    float u_inv_scale;        // 2^-eu
    int u_exp;                // eu
    Seg seg[MAX_SEGS];
};

// Histogram bin of np.histogram(x, bins=np.linspace(lo, hi, nb + 1)) from a float32 pre-binning: the
// exact float64 edge comparison is only needed when the value sits within 1e-3 of a bin edge (float32
// rounding of (x - lo) * nb / (hi - lo) is ~1e-5 of a bin).  Near edge e = round(t) the value lies
// strictly between edges e-1 and e+1, so the bin is e or e-1 by ONE float64 comparison against
// lo + e*step -- no float64 division.  -1 = outside [lo, hi] (or NaN).  Used by K3 and by K1's fused
// epilogue, so both produce the same counts.
__device__ __forceinline__ int hist_bin_fast(float x, float flo, float finv, double lo, double hi, double step,
                                             int nb) {
    const float t = (x - flo) * finv;
    const float fl = floorf(t);
    const float fr = t - fl;
    if (fr > 1e-3f && fr < 0.999f && t > 0.f && t < static_cast<float>(nb)) return static_cast<int>(fl);
    const float r = rintf(t);
    if (!(r >= 0.f) || r > static_cast<float>(nb)) return -1;   // outside by more than half a bin
    const int e = static_cast<int>(r);
    const double xd = static_cast<double>(x);
    if (e == 0) return xd >= lo ? 0 : -1;
    if (e == nb) return xd <= hi ? nb - 1 : -1;
    return xd >= lo + e * step ? e : e - 1;
}

// ------------------------------------------------------------------ launchers
// k0_prep.cu
// d_zeroE (may be null): per chunk, set to 1 when a window's energy is zero within round-off; such
// windows get invE = +inf (the reference's x/0, detect.py:577) and launch_zero_energy_fix then makes
// their DS +inf for every subspace instead of the 0*inf = NaN of an exactly-zero projection.
void launch_k0(const void* raw, int dtype_f32, const ChunkDesc* d_chunks, int nchunks, int Nc, int n,
               int max_Lpad, int max_ntiles, double* d_sum, unsigned* d_maxbits, float* d_scale,
               __half* d_xsplit, float* d_mu, float* d_invE, int x8_policy, float k4_limit,
               unsigned* d_k4bits, int* d_chunk_mode, int* d_zeroE, int* d_chunk_bad, cudaStream_t st);
// DS[s][t] = +inf for every subspace row at the lags whose window energy is zero (invE = +inf), for
// the chunks k0_norm flagged.  nrows = DS rows per chunk.  DS64 may be null.
void launch_zero_energy_fix(const ChunkDesc* d_chunks, int nchunks, int max_ntiles, int nrows, const float* d_invE,
                            const int* d_zeroE, float* d_DS, double* d_DS64, cudaStream_t st);
// Rank > 16 subspaces: their 16-vector pieces land in separate DS rows (the first in the subspace's
// own row, the others in scratch rows behind the S subspace rows); this adds the scratch rows to the
// subspace row in piece order, so the result does not depend on the order the CTAs finished in.
struct PieceSum { int dst_row, src_row; };
void launch_sum_pieces(const ChunkDesc* d_chunks, int nchunks, int max_Tpad, const PieceSum* d_pieces, int npieces,
                       float* d_DS, cudaStream_t st);

// basis image (k1_project.cu)
// per row of U: {sum, max |u|, sum u^2, sum u^4} (float64)
void launch_basis_row_stats(const double* d_U, int R, int n, double* d_out, cudaStream_t st);
void launch_basis_image(const double* d_U, const int* d_slot_row, const BasisLayout& lay,
                        uint8_t* d_Aimg, int x8, cudaStream_t st);

// one candidate trigger (K3 / K1's fused epilogue -> launch_lta -> host)
struct Candidate {
    int row;     // chunk * S + subspace
    int t;       // lag index
    float ds;    // detection statistic
    float lta;   // denominator of DS_STALTA: |ds| / lta = STA / LTA at t (filled by launch_lta)
};

// k1_project.cu : tcgen05 Hankel projection + normalisation -> DS
struct K1Args {
    const uint8_t* Aimg;      // fp16 hi | fp16 lo tiles (3 MMAs per K step)
    const uint8_t* Aimg8;     // fp16 hi | (e4m3 u_lo, e4m3 u_hi) byte-pair tiles (2 MMAs per K step), may be null
    const int* chunk_mode;    // per chunk: 1 = 8-bit cross terms (written by k0_split)
    const __half* xsplit;
    const float* mu;
    const float* invE;
    const float* chunk_scale;
    const ChunkDesc* chunks;
    const int4* items;     // (chunk, tile, basis block, 0)
    const BlockInfo* binfo;
    float* DS;
    int nitems;
    int kblk;      // 64-tap chunks accumulated in TMEM between drains (fp16 cross terms)
    int kblk8;     // same for chunks in 8-bit cross-term mode
    int num_sms;
    int nq;        // MMA N: 256 (tiles of 2048 lags) or 128 (tiles of 1024 lags)
    int mode;      // 0 = detection statistic, 1 = signed correlation coefficient (CCX)
    // fused epilogue (MODE 2): the per-row reductions of K3 run on the statistic while it is still in
    // registers and DS is never written.  Same outputs as launch_k3.
    int fused;                       // 1 = MODE 2
    const float* thr;                // [S] thresholds (INFINITY = none)
    unsigned* rowmax_bits;           // [rows] float bits, zeroed before the launch (DS >= 0: uint order == float order)
    int* rowflags;                   // [rows], zeroed before the launch; bit 1 = zero-energy windows counted as 0
    unsigned long long* hist;        // [S][HIST_MAX_BINS]
    double hist_lo, hist_hi;
    int nbins;
    Candidate* cand;
    int cand_cap;
    int* ncand;
    double* fas;                     // [S][5] or null
    int row_base;                    // rows of this batch start here (accumulation over batches)
    const int* chunk_bad;            // [chunks] 1 = non-finite samples: every row of the chunk is NaN
    int S;                           // subspaces (rows per chunk)
    int cg2;       // 1 = CTA pairs (cta_group::2): items are (chunk, tile, EVEN basis block), the two CTAs of a
                   // cluster take blocks it.z and it.z + 1 and share the tile's B rows (N = 256 only)
    int dual;      // 1 = CCX screening series on PAIRS of chunks (items (chunk, tile, block, partner chunk)): two
                   //     N = 128 MMAs per K step share the A tile, the basis image is streamed once per pair
    int hi_only;   // 1 = ONE MMA per K step (fp16 hi * hi only, 11-bit operands): a screening series whose
                   // error is bounded by ~2^-10 of the normalised value; CCX uses it to LOCATE the maximum,
                   // which is then re-scored in float64 (k4_ccx.cu)
};
void launch_k1(const K1Args& a, const BasisLayout& lay, cudaStream_t st);
int k1_smem_bytes();

// k_direct.cu : float64 CUDA-core evaluation of the closed form (validation / small jobs)
void launch_direct(const void* raw, int dtype_f32, const ChunkDesc* d_chunks, int nchunks,
                   const double* d_U, const int* d_rank_off, int S, int n, int Nc, int maxT,
                   const double* d_sum, float* d_DS, double* d_DS64, cudaStream_t st);

// fused mode (K1 MODE 2): LTA windows of the candidates from the float64 closed form, NaN rows of bad chunks
void launch_lta_direct(const void* raw, int dtype_f32, const ChunkDesc* d_chunks, const double* d_sum, const double* d_U,
                       const int* d_rank_off, int n, int Nc, int S, Candidate* d_cand, const int* d_ncand,
                       const int* d_ncand_before, int cand_cap, int row_base, int W, int Wsta, double* d_acc2,
                       cudaStream_t st);
void launch_fused_bad_rows(const int* d_chunk_bad, int nchunks, int S, int row_base, unsigned* d_rowmax_bits,
                           int* d_rowflags, cudaStream_t st);

// k3_post.cu : per (chunk, subspace) row -> max, histogram, candidates
// row_base: candidate rows and the rowmax / rowflags entries of this batch start at row_base
// (= chunks of earlier batches * S when results accumulate over the batches of a station)
void launch_k3(const float* DS, const ChunkDesc* d_chunks, int nchunks, int S, const float* d_thr,
               float* d_rowmax, int* d_rowflags, unsigned long long* d_hist, double hist_lo,
               double hist_hi, int nbins, Candidate* d_cand, int cand_cap, int* d_ncand, double* d_fas,
               int row_base, cudaStream_t st);
// candidates [*d_ncand_before, *d_ncand) belong to the current batch (d_ncand_before may be null = 0)
void launch_lta(const float* DS, const ChunkDesc* d_chunks, int S, const int* d_rowflags,
                Candidate* d_cand, const int* d_ncand, const int* d_ncand_before, int cand_cap, int row_base, int W,
                int Wsta, cudaStream_t st);

// k4_ccx.cu : pairwise CCX
void launch_ccx_stats(const void* d_X, int dtype_f32, int N, int n, int Nc, double* wa, double* wb, double* es,
                      double* ed, cudaStream_t st);
// template rows: d_rows[0..nrows) if given, else row_begin .. row_begin + nrows - 1
void launch_ccx_fp64(const void* d_X, int dtype_f32, int N, int n, int Nc, int row_begin, int nrows, const int* d_rows,
                     const double* wa, const double* wb, const double* es, const double* ed, double* d_cc,
                     int* d_lag, double* d_sub, int num_sms, cudaStream_t st);
void launch_ccx_fp64_pairs(const void* d_X, int dtype_f32, int N, int n, int Nc, const int* d_rows, const double* wa,
                           const double* wb, const double* es, const double* ed, double* d_cc, int* d_lag,
                           double* d_sub, const int2* d_pairs, const int* d_npairs, int pair_cap, int num_sms,
                           cudaStream_t st);
void launch_corr0(const double* d_X, int N, int n, double* d_out, cudaStream_t st);
// d_rows[r] = event index of template row r (sorted ascending)
void launch_ccx_templates(const void* d_X, int dtype_f32, int n, const int* d_rows, int rows, double* d_U,
                          cudaStream_t st);
void launch_ccx_pad(const void* d_X, int dtype_f32, int n, int Nc, int c0, int nsig, int P, int Lc, double* out,
                    cudaStream_t st);
void launch_ccx_demux(const void* d_X, int dtype_f32, int N, int n, int Nc, double* d_Xd, cudaStream_t st);
void launch_ccx_post(const float* DS, const ChunkDesc* d_chunks, int c0, int nsig, const void* d_X, int dtype_f32,
                     const double* d_Xd, int N, int n, int Nc, const int* d_rows, int nrows, const double* wa,
                     const double* wb, const double* es, const double* ed, double* d_cc, int* d_lag, double* d_sub,
                     int* d_nflag, int2* d_flagged, int flag_cap, int4* d_karg, const unsigned* d_ratio_bits,
                     float band0, cudaStream_t st);   // d_karg: [nsig][nrows] scratch; d_ratio_bits: per signal, from k0_norm
// dense per-slot rows [nslots][N] -> SciPy condensed order (pair (b, c), b < c, at b*N - b(b+1)/2 + c-b-1);
// d_slot_of_row[b] = slot holding event b's row
void launch_ccx_pack_rows(const double* d_cc, const int* d_lag, const double* d_sub, const int* d_rows, int nrows, int N,
                          double* o_cc, int* o_lag, double* o_sub, cudaStream_t st);
void launch_ccx_pack(const double* d_cc, const int* d_lag, const double* d_sub, const int* d_slot_of_row, int N,
                     double* o_cc, int* o_lag, double* o_sub, cudaStream_t st);

// k7_mag.cu : per-detection magnitude / SNR estimates (_estMag)
constexpr int MAG_MAX_RANK = 64;
struct MagTrigger {
    int chunk, subspace, t, pad;
};
struct MagSubspace {
    int is_single, row0, rank, nev;
    const double* ewf;      // [nev][n] event waveforms (single: WFU[0])
    const double* mags;     // [nev]
    const double* ev_mean;  // [nev]
    const double* ev_std;   // [nev] population std
    const double* wfu_var;  // [nev] var(WFU_i), population
};
void launch_mag(const void* raw, int dtype_f32, const ChunkDesc* d_chunks, const double* d_sum,
                const MagTrigger* d_trig, int ntrig, const MagSubspace* d_subs, const double* d_U, int n, int Nc,
                double* d_scratch, int scratch_stride, double* d_out, cudaStream_t st);

// k8_preproc.cu : detrend + SOS filter + multiplex (construct._applyFilter array part)
void launch_preproc(double* d_buf, const long long* d_off, const int* d_len, int ntr, int maxlen, const double* sos,
                    int nsos, int zerophase, int detrend, double* d_stats, double* d_segstate, cudaStream_t st);
void launch_multiplex(const double* d_buf, const long long* d_off, const int* d_minlen, const long long* d_out_off,
                      int nchunks, int Nc, int maxlen, double* d_out, cudaStream_t st);
void launch_decimate(const double* d_src, const long long* d_off, double* d_dst, const long long* d_off2,
                     const int* d_len2, int ntr, int factor, cudaStream_t st);
int preproc_seg();

// k6_stalta.cu : classic STA/LTA screen of raw chunks (fas._checkSTALTA)
void launch_stalta_max(const void* raw, int dtype_f32, const long long* d_raw_off, const int* d_Ls, int nchunks,
                       int maxLs, int Nc, int chan, int nsta, int nlta, unsigned* d_out_bits, cudaStream_t st);

void launch_stalta_dense(const float* row, int T, int W, int Wsta, int zero_inf, float* out, float* tmp,
                         cudaStream_t st);
int stalta_dense_max_window();

}  // namespace dtx
