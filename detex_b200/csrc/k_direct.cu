// k_direct.cu -- float64 CUDA-core evaluation of the detection-statistic closed form.
//
//   DS[t] = ((n-1)/n) * sum_k (u_k . (w_t - mean w_t))^2 / ||w_t - mean w_t||^2,
//   w_t = x[t*Nc : t*Nc + n]        (reference detex/detect.py:559-578)
//
// This is NOT the production path (it is O(n) FMAs per basis vector per lag on the FP64
// pipe); it exists so that the tcgen05 path can be checked on the device at full BASELINE
// sizes against an independent float64 evaluation, and for tiny jobs.
#include "dtx_kernels.cuh"

namespace dtx {
namespace {

constexpr int RMAX = 16;
constexpr int JT = 128;

template <typename T>
__global__ void __launch_bounds__(128)
direct_kernel(const T* __restrict__ raw, const ChunkDesc* __restrict__ chunks,
              const double* __restrict__ U, const int* __restrict__ rank_off, int n, int Nc,
              const double* __restrict__ sum, float* __restrict__ DS, double* __restrict__ DS64) {
    const ChunkDesc cd = chunks[blockIdx.z];
    const int s = blockIdx.y;
    const int t = blockIdx.x * 128 + threadIdx.x;
    if (blockIdx.x * 128 >= cd.T) return;
    const int k0 = rank_off[s], r = rank_off[s + 1] - k0;
    const double mean = sum[blockIdx.z] / static_cast<double>(cd.L);
    const T* x = raw + cd.raw_off;
    __shared__ double Us[RMAX][JT];
    const bool live = t < cd.T;
    const long long o = static_cast<long long>(live ? t : 0) * Nc;
    double s1 = 0.0, s2 = 0.0, num = 0.0;
    const double nn = static_cast<double>(n);
    for (int kc = 0; kc < r; kc += RMAX) {   // ranks above RMAX: several passes over the window
        const int rc = min(RMAX, r - kc);
        double acc[RMAX], su[RMAX];
#pragma unroll
        for (int k = 0; k < RMAX; ++k) { acc[k] = 0.0; su[k] = 0.0; }
        double p1 = 0.0, p2 = 0.0;
        for (int j0 = 0; j0 < n; j0 += JT) {
            const int jn = min(JT, n - j0);
            __syncthreads();
            for (int i = threadIdx.x; i < RMAX * JT; i += 128) {
                const int k = i / JT, j = i % JT;
                Us[k][j] = (k < rc && j < jn) ? U[static_cast<long long>(k0 + kc + k) * n + j0 + j] : 0.0;
            }
            __syncthreads();
            for (int j = 0; j < jn; ++j) {
                const double xv = static_cast<double>(x[o + j0 + j]) - mean;
                p1 += xv;
                p2 += xv * xv;
#pragma unroll
                for (int k = 0; k < RMAX; ++k) {
                    acc[k] = fma(Us[k][j], xv, acc[k]);
                    su[k] += Us[k][j];
                }
            }
        }
        s1 = p1;
        s2 = p2;
        const double mu = s1 / nn;
#pragma unroll
        for (int k = 0; k < RMAX; ++k) {
            const double c = acc[k] - mu * su[k];
            num += (k < rc) ? c * c : 0.0;
        }
    }
    if (!live) return;
    double E = s2 - s1 * s1 / nn;
    if (E < 0.0) E = 0.0;
    const double ds = ((nn - 1.0) / nn) * num / E;
    const long long idx = cd.ds_off + static_cast<long long>(s) * cd.Tpad + t;
    if (DS) DS[idx] = static_cast<float>(ds);
    if (DS64) DS64[idx] = ds;
}

// ------------------------------------------------------------------------------------------
// Fused mode (K1 MODE 2 never writes DS): the denominator of DS_STALTA at the few candidate lags
// needs the statistic in a window of W lags around each (detect.py:501-524).  One CTA per 128 lags
// of a candidate's window evaluates the float64 closed form there (the same arithmetic as
// direct_kernel; ranks <= 16), sums |DS| over the LTA (and STA) window and adds the partial sums to
// the candidate's accumulators.  Sparse by construction: a station-month has a few hundred candidates.
__device__ __forceinline__ void centred_window(int T, int W, int i, int& a0) {
    // pandas' centred rolling window + _replaceNanWithMean's edge rule (detect.py:517-524), as
    // centred_abs_mean in k3_post.cu
    const int off = (W - 1) / 2;
    const int first = W - 1 - off, last = T - 1 - off;
    if (i < first) i = (first + 1 <= last) ? first + 1 : first;
    if (i > last) i = last;
    a0 = i - (W - 1) + off;
}

template <typename T>
__global__ void __launch_bounds__(128)
lta_direct_kernel(const T* __restrict__ raw, const ChunkDesc* __restrict__ chunks, const double* __restrict__ sum,
                  const double* __restrict__ U, const int* __restrict__ rank_off, int n, int Nc, int S,
                  const Candidate* __restrict__ cand, const int* __restrict__ ncand,
                  const int* __restrict__ ncand_before, int cand_cap, int row_base, int W, int Wsta,
                  double* __restrict__ acc2) {
    __shared__ double Us[RMAX][JT];
    __shared__ double red[4][2];
    const int nc = min(*ncand, cand_cap);
    const double nn = static_cast<double>(n);
    for (int ci = (ncand_before ? *ncand_before : 0) + blockIdx.y; ci < nc; ci += gridDim.y) {
        const Candidate c = cand[ci];
        const int lrow = c.row - row_base;
        const ChunkDesc cd = chunks[lrow / S];
        const int s = lrow % S;
        if (cd.T < W || cd.T < Wsta) {                 // no STA/LTA array in the reference: NaN (host: 0.0)
            if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&acc2[2 * ci], nan(""));
            continue;
        }
        int aL, aS = 0;
        centred_window(cd.T, W, c.t, aL);
        int lo = aL, hi = aL + W;
        if (Wsta > 0) {
            centred_window(cd.T, Wsta, c.t, aS);
            lo = min(lo, aS);
            hi = max(hi, aS + Wsta);
        }
        if (blockIdx.x * 128 >= hi - lo) continue;     // CTA-uniform
        const int t = lo + blockIdx.x * 128 + threadIdx.x;
        const bool live = t < hi;
        const int k0 = rank_off[s], r = min(RMAX, rank_off[s + 1] - k0);
        const double mean = sum[lrow / S] / static_cast<double>(cd.L);
        const T* x = raw + cd.raw_off;
        const long long o = static_cast<long long>(live ? t : lo) * Nc;
        double accv[RMAX], su[RMAX];
#pragma unroll
        for (int k = 0; k < RMAX; ++k) { accv[k] = 0.0; su[k] = 0.0; }
        double p1 = 0.0, p2 = 0.0;
        for (int j0 = 0; j0 < n; j0 += JT) {
            const int jn = min(JT, n - j0);
            __syncthreads();
            for (int i = threadIdx.x; i < RMAX * JT; i += 128) {
                const int k = i / JT, j = i % JT;
                Us[k][j] = (k < r && j < jn) ? U[static_cast<long long>(k0 + k) * n + j0 + j] : 0.0;
            }
            __syncthreads();
            for (int j = 0; j < jn; ++j) {
                const double xv = static_cast<double>(x[o + j0 + j]) - mean;
                p1 += xv;
                p2 += xv * xv;
#pragma unroll
                for (int k = 0; k < RMAX; ++k) {
                    accv[k] = fma(Us[k][j], xv, accv[k]);
                    su[k] += Us[k][j];
                }
            }
        }
        const double mu = p1 / nn;
        double num = 0.0;
#pragma unroll
        for (int k = 0; k < RMAX; ++k) {
            const double cc = accv[k] - mu * su[k];
            num += (k < r) ? cc * cc : 0.0;
        }
        double E = p2 - p1 * p1 / nn;
        if (E < 0.0) E = 0.0;
        double v = fabs(((nn - 1.0) / nn) * num / E);
        // the float32 statistic of the dense path is what the reference-shaped LTA averages; infs
        // (zero-energy windows) count as 0 (detect.py:275-281)
        v = static_cast<double>(static_cast<float>(v));
        if (isinf(v) || E == 0.0) v = 0.0;
        double vL = (live && t >= aL && t < aL + W) ? v : 0.0;
        double vS = (live && Wsta > 0 && t >= aS && t < aS + Wsta) ? v : 0.0;
        for (int of = 16; of > 0; of >>= 1) {
            vL += __shfl_xor_sync(0xffffffffu, vL, of);
            vS += __shfl_xor_sync(0xffffffffu, vS, of);
        }
        __syncthreads();
        if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5][0] = vL; red[threadIdx.x >> 5][1] = vS; }
        __syncthreads();
        if (threadIdx.x == 0) {
            atomicAdd(&acc2[2 * ci], red[0][0] + red[1][0] + red[2][0] + red[3][0]);
            if (Wsta > 0) atomicAdd(&acc2[2 * ci + 1], red[0][1] + red[1][1] + red[2][1] + red[3][1]);
        }
    }
}

// cand.lta from the window sums (and reset them): LTA mean, or LTA mean * |DS| / STA mean (see lta_kernel)
__global__ void __launch_bounds__(256)
lta_finish_kernel(Candidate* __restrict__ cand, const int* __restrict__ ncand, const int* __restrict__ ncand_before,
                  int cand_cap, int W, int Wsta, double* __restrict__ acc2) {
    const int nc = min(*ncand, cand_cap);
    const int ci = (ncand_before ? *ncand_before : 0) + blockIdx.x * 256 + threadIdx.x;
    if (ci >= nc) return;
    double out = acc2[2 * ci] / W;
    if (Wsta > 0) out = out * fabs(static_cast<double>(cand[ci].ds)) / (acc2[2 * ci + 1] / Wsta);
    cand[ci].lta = static_cast<float>(out);
    acc2[2 * ci] = 0.0;
    acc2[2 * ci + 1] = 0.0;
}

// rows of chunks with non-finite samples: MaxDS = NaN, flag bit 0 (what K3 reports for a row with NaN)
__global__ void __launch_bounds__(256)
fused_bad_rows_kernel(const int* __restrict__ chunk_bad, int nchunks, int S, int row_base, unsigned* __restrict__ rowmax_bits,
                      int* __restrict__ rowflags) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= nchunks * S || !chunk_bad[i / S]) return;
    rowmax_bits[row_base + i] = 0x7fc00000u;
    rowflags[row_base + i] = 1;
}

}  // namespace

void launch_lta_direct(const void* raw, int dtype_f32, const ChunkDesc* d_chunks, const double* d_sum, const double* d_U,
                       const int* d_rank_off, int n, int Nc, int S, Candidate* d_cand, const int* d_ncand,
                       const int* d_ncand_before, int cand_cap, int row_base, int W, int Wsta, double* d_acc2,
                       cudaStream_t st) {
    const dim3 grid((W + Wsta + 127) / 128 + 1, 1024);
    if (dtype_f32)
        lta_direct_kernel<float><<<grid, 128, 0, st>>>(static_cast<const float*>(raw), d_chunks, d_sum, d_U, d_rank_off, n,
                                                       Nc, S, d_cand, d_ncand, d_ncand_before, cand_cap, row_base, W,
                                                       Wsta, d_acc2);
    else
        lta_direct_kernel<double><<<grid, 128, 0, st>>>(static_cast<const double*>(raw), d_chunks, d_sum, d_U, d_rank_off,
                                                        n, Nc, S, d_cand, d_ncand, d_ncand_before, cand_cap, row_base, W,
                                                        Wsta, d_acc2);
    // candidates of this batch: at most cand_cap; the finish grid covers the capacity, threads beyond exit
    lta_finish_kernel<<<(cand_cap + 255) / 256, 256, 0, st>>>(d_cand, d_ncand, d_ncand_before, cand_cap, W, Wsta, d_acc2);
}

void launch_fused_bad_rows(const int* d_chunk_bad, int nchunks, int S, int row_base, unsigned* d_rowmax_bits,
                           int* d_rowflags, cudaStream_t st) {
    fused_bad_rows_kernel<<<(nchunks * S + 255) / 256, 256, 0, st>>>(d_chunk_bad, nchunks, S, row_base, d_rowmax_bits,
                                                                     d_rowflags);
}

void launch_direct(const void* raw, int dtype_f32, const ChunkDesc* d_chunks, int nchunks,
                   const double* d_U, const int* d_rank_off, int S, int n, int Nc, int maxT,
                   const double* d_sum, float* d_DS, double* d_DS64, cudaStream_t st) {
    const dim3 grid((maxT + 127) / 128, S, nchunks);
    if (dtype_f32)
        direct_kernel<float><<<grid, 128, 0, st>>>(static_cast<const float*>(raw), d_chunks, d_U,
                                                   d_rank_off, n, Nc, d_sum, d_DS, d_DS64);
    else
        direct_kernel<double><<<grid, 128, 0, st>>>(static_cast<const double*>(raw), d_chunks, d_U,
                                                    d_rank_off, n, Nc, d_sum, d_DS, d_DS64);
}

}  // namespace dtx
