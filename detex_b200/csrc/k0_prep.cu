// k0_prep.cu -- K0: conditioning of the continuous data for the tensor-core projection.
//
// Replaces `pd.rolling_mean` / `pd.rolling_var` of _MPXDS (reference detex/detect.py:567-569,
// detex/fas.py:126-127) and the de-multiplexing implied by `result[::Nc]` (detect.py:578).
//
//   k0_stats : per chunk sum(x) and max|x|                     (HBM-bound, 1 read)
//   k0_split : x -> (x - mean) * 2^ex -> fp16 hi + lo per channel, zero padded
//   k0_norm  : window mean mu[t] and invE[t] = ((n-1)/n) / (S2 - S1^2/n) for every
//              channel-aligned lag t, float64 running sums (block scan of the in/out
//              differences), so the denominator never suffers the cancellation a
//              float32 prefix sum would.
//
// DS is invariant to adding a constant to x and to scaling x, so centring on the chunk
// mean and scaling by a power of two change nothing mathematically; they put the data in
// the range where the fp16 hi/lo split is worth 22 bits.
#include "dtx_kernels.cuh"

namespace dtx {

template <typename T>
__global__ void __launch_bounds__(256)
k0_stats(const T* __restrict__ raw, const ChunkDesc* __restrict__ chunks, double* __restrict__ sum,
         unsigned* __restrict__ maxbits) {
    const ChunkDesc cd = chunks[blockIdx.y];
    const T* x = raw + cd.raw_off;
    double s = 0.0;
    float m = 0.f;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < cd.L; i += gridDim.x * blockDim.x) {
        const double v = static_cast<double>(x[i]);
        s += v;
        m = fmaxf(m, fabsf(static_cast<float>(v)));
    }
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    }
    __shared__ double ss[8];
    __shared__ float sm[8];
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) {
        ss[w] = s;
        sm[w] = m;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < 8; ++i) {
            s += ss[i];
            m = fmaxf(m, sm[i]);
        }
        atomicAdd(&sum[blockIdx.y], s);
        atomicMax(&maxbits[blockIdx.y], __float_as_uint(m));  // m >= 0: uint order == float order
    }
}

// scale exponent: (max|x| + |mean|) * 2^ex in [2^14, 2^15)  (fp16 max is 65504)
__device__ __forceinline__ int scale_exp(float maxabs, double mean) {
    const float bound = maxabs + fabsf(static_cast<float>(mean));
    if (!(bound > 0.f) || !isfinite(bound)) return 0;
    return 14 - ilogbf(bound);
}

template <typename T>
__global__ void __launch_bounds__(256)
k0_split(const T* __restrict__ raw, const ChunkDesc* __restrict__ chunks,
         const double* __restrict__ sum, const unsigned* __restrict__ maxbits,
         float* __restrict__ scale_out, __half* __restrict__ xsplit, int Nc) {
    const ChunkDesc cd = chunks[blockIdx.y];
    const double mean = sum[blockIdx.y] / static_cast<double>(cd.L);
    const int ex = scale_exp(__uint_as_float(maxbits[blockIdx.y]), mean);
    if (blockIdx.x == 0 && threadIdx.x == 0) scale_out[blockIdx.y] = exp2f(static_cast<float>(-ex));
    const T* x = raw + cd.raw_off;
    __half* out = xsplit + cd.sig_off;
    const long long total = static_cast<long long>(Nc) * cd.Lpad;
    for (long long j = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; j < total;
         j += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int c = static_cast<int>(j % Nc);
        const int i = static_cast<int>(j / Nc);
        float v = 0.f;
        if (j < cd.L) v = static_cast<float>(ldexp(static_cast<double>(x[j]) - mean, ex));
        const __half hi = __float2half_rn(v);
        const __half lo = __float2half_rn(v - __half2float(hi));
        out[static_cast<long long>(c * 2 + 0) * cd.Lpad + i] = hi;
        out[static_cast<long long>(c * 2 + 1) * cd.Lpad + i] = lo;
    }
}

// One block per tile of TILE_T output lags.
template <typename T>
__global__ void __launch_bounds__(256)
k0_norm(const T* __restrict__ raw, const ChunkDesc* __restrict__ chunks,
        const double* __restrict__ sum, float* __restrict__ mu, float* __restrict__ invE, int Nc,
        int n) {
    const ChunkDesc cd = chunks[blockIdx.y];
    if (blockIdx.x >= cd.ntiles) return;
    const double mean = sum[blockIdx.y] / static_cast<double>(cd.L);
    const T* x = raw + cd.raw_off;
    const int t0 = blockIdx.x * TILE_T;
    const int tid = threadIdx.x, w = tid >> 5, l = tid & 31;
    __shared__ double sh1[8], sh2[8];
    __shared__ double base1, base2;

    // window sums at t0 (t0 < T always holds for an existing tile)
    double a1 = 0.0, a2 = 0.0;
    {
        const long long o = static_cast<long long>(t0) * Nc;
        for (int j = tid; j < n; j += 256) {
            const double v = static_cast<double>(x[o + j]) - mean;
            a1 += v;
            a2 += v * v;
        }
        for (int s = 16; s > 0; s >>= 1) {
            a1 += __shfl_xor_sync(0xffffffffu, a1, s);
            a2 += __shfl_xor_sync(0xffffffffu, a2, s);
        }
        if (l == 0) {
            sh1[w] = a1;
            sh2[w] = a2;
        }
        __syncthreads();
        if (tid == 0) {
            double b1 = 0, b2 = 0;
            for (int i = 0; i < 8; ++i) {
                b1 += sh1[i];
                b2 += sh2[i];
            }
            base1 = b1;
            base2 = b2;
        }
        __syncthreads();
    }
    // differences: d[i] moves the window from t0+i-1 to t0+i (i >= 1); thread owns 8 lags
    constexpr int PER = TILE_T / 256;
    double d1[PER], d2[PER];
    double r1 = 0.0, r2 = 0.0;
#pragma unroll
    for (int k = 0; k < PER; ++k) {
        const int i = tid * PER + k;
        const int t = t0 + i;
        double e1 = 0.0, e2 = 0.0;
        if (i >= 1 && t < cd.T) {
            const long long o = static_cast<long long>(t - 1) * Nc;
            for (int c = 0; c < Nc; ++c) {
                const double vin = static_cast<double>(x[o + n + c]) - mean;
                const double vout = static_cast<double>(x[o + c]) - mean;
                e1 += vin - vout;
                e2 += vin * vin - vout * vout;
            }
        }
        r1 += e1;
        r2 += e2;
        d1[k] = r1;
        d2[k] = r2;
    }
    // exclusive scan of per-thread totals across the block
    double p1 = r1, p2 = r2;
    for (int s = 1; s < 32; s <<= 1) {
        const double q1 = __shfl_up_sync(0xffffffffu, p1, s);
        const double q2 = __shfl_up_sync(0xffffffffu, p2, s);
        if (l >= s) {
            p1 += q1;
            p2 += q2;
        }
    }
    __syncthreads();
    if (l == 31) {
        sh1[w] = p1;
        sh2[w] = p2;
    }
    __syncthreads();
    double o1 = p1 - r1, o2 = p2 - r2;  // exclusive within warp
    for (int i = 0; i < w; ++i) {
        o1 += sh1[i];
        o2 += sh2[i];
    }
    const double nn = static_cast<double>(n);
    const double cn = (nn - 1.0) / nn;
    float* pm = mu + cd.norm_off + t0;
    float* pe = invE + cd.norm_off + t0;
#pragma unroll
    for (int k = 0; k < PER; ++k) {
        const int i = tid * PER + k;
        const int t = t0 + i;
        float fm = 0.f, fe = 0.f;
        if (t < cd.T) {
            const double s1 = base1 + o1 + d1[k];
            const double s2 = base2 + o2 + d2[k];
            double E = s2 - s1 * s1 / nn;
            if (E < 0.0) E = 0.0;
            fm = static_cast<float>(s1 / nn);
            fe = static_cast<float>(cn / E);  // E == 0 -> +inf, as the reference's x/0
        }
        pm[i] = fm;
        pe[i] = fe;
    }
}

void launch_k0(const void* raw, int dtype_f32, const ChunkDesc* d_chunks, int nchunks, int Nc, int n,
               int max_Lpad, int max_ntiles, double* d_sum, unsigned* d_maxbits, float* d_scale,
               __half* d_xsplit, float* d_mu, float* d_invE, cudaStream_t st) {
    cudaMemsetAsync(d_sum, 0, sizeof(double) * nchunks, st);
    cudaMemsetAsync(d_maxbits, 0, sizeof(unsigned) * nchunks, st);
    const dim3 g1(64, nchunks);
    long long tot = static_cast<long long>(Nc) * max_Lpad;
    int gx = static_cast<int>((tot + 256 * 8 - 1) / (256 * 8));
    if (gx < 1) gx = 1;
    const dim3 g2(gx, nchunks);
    const dim3 g3(max_ntiles, nchunks);
    if (dtype_f32) {
        const float* r = static_cast<const float*>(raw);
        k0_stats<float><<<g1, 256, 0, st>>>(r, d_chunks, d_sum, d_maxbits);
        k0_split<float><<<g2, 256, 0, st>>>(r, d_chunks, d_sum, d_maxbits, d_scale, d_xsplit, Nc);
        k0_norm<float><<<g3, 256, 0, st>>>(r, d_chunks, d_sum, d_mu, d_invE, Nc, n);
    } else {
        const double* r = static_cast<const double*>(raw);
        k0_stats<double><<<g1, 256, 0, st>>>(r, d_chunks, d_sum, d_maxbits);
        k0_split<double><<<g2, 256, 0, st>>>(r, d_chunks, d_sum, d_maxbits, d_scale, d_xsplit, Nc);
        k0_norm<double><<<g3, 256, 0, st>>>(r, d_chunks, d_sum, d_mu, d_invE, Nc, n);
    }
}

}  // namespace dtx
