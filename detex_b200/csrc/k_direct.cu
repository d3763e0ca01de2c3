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

}  // namespace

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
