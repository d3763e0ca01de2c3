// k0_prep.cu -- K0: conditioning of the continuous data for the tensor-core projection.
//
// Replaces `pd.rolling_mean` / `pd.rolling_var` of _MPXDS (reference detex/detect.py:567-569,
// detex/fas.py:126-127) and the de-multiplexing implied by `result[::Nc]` (detect.py:578).
//
//   k0_stats : per chunk sum(x) and max|x|                     (HBM-bound, 1 read)
//   k0_norm  : window mean mu[t] and invE[t] = ((n-1)/n) / (S2 - S1^2/n) for every
//              channel-aligned lag t, float64 running sums (block scan of the in/out
//              differences), so the denominator never suffers the cancellation a
//              float32 prefix sum would.  For the adaptive engine also the worst window
//              kurtosis of the chunk (precision policy, see k0_split).
//   k0_split : x -> (x - mean) * 2^ex -> fp16 hi + lo per channel, zero padded; the lo plane
//              holds fp16 residuals or, for chunks in 8-bit cross-term mode, e5m2 byte pairs
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
         float* __restrict__ scale_out, __half* __restrict__ xsplit, int Nc, int x8_policy,
         float k4_limit, const unsigned* __restrict__ k4bits, int* __restrict__ chunk_mode,
         int* __restrict__ chunk_bad) {
    const ChunkDesc cd = chunks[blockIdx.y];
    // precision mode of this chunk ("8-bit cross terms"): off, forced, or chosen from the
    // worst window kurtosis K4 = sum (x - chunk mean)^4 / ||w - window mean||^4 that k0_norm found
    // (NaN / inf / zero-energy windows give K4 = 3e38, i.e. the fp16 cross terms)
    const bool x8 = x8_policy == X8_FORCE ||
                    (x8_policy == X8_AUTO && __uint_as_float(k4bits[blockIdx.y]) <= k4_limit);
    if (blockIdx.x == 0 && threadIdx.x == 0) chunk_mode[blockIdx.y] = x8 ? 1 : 0;
    const double mean = sum[blockIdx.y] / static_cast<double>(cd.L);
    const int ex = scale_exp(__uint_as_float(maxbits[blockIdx.y]), mean);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        scale_out[blockIdx.y] = exp2f(static_cast<float>(-ex));
        // a NaN / inf sample makes the chunk's sum non-finite: every statistic of the chunk is NaN in the
        // reference (FFT of the whole chunk) and every row is dropped here (K3: has_nan; fused: this flag)
        if (chunk_bad) chunk_bad[blockIdx.y] = isfinite(mean) ? 0 : 1;
    }
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
        __half lo;
        if (!x8) lo = __float2half_rn(v - __half2float(hi));
        else {
            // 8-bit cross-term plane: byte 0 = e5m2(x_hi * 2^-X8_S), byte 1 = e5m2(x_lo * 2^X8_S)
            const float fh = __half2float(hi);
            const unsigned b0 = __nv_cvt_float_to_fp8(ldexpf(fh, -X8_SHIFT), __NV_SATFINITE, __NV_E5M2);
            const unsigned b1 = __nv_cvt_float_to_fp8(ldexpf(v - fh, X8_SHIFT), __NV_SATFINITE, __NV_E5M2);
            lo = __ushort_as_half(static_cast<unsigned short>(b0 | (b1 << 8)));
        }
        out[static_cast<long long>(c * 2 + 0) * cd.Lpad + i] = hi;
        out[static_cast<long long>(c * 2 + 1) * cd.Lpad + i] = lo;
    }
}

// One block per tile of TILE_T output lags.
template <typename T>
__global__ void __launch_bounds__(256)
k0_norm(const T* __restrict__ raw, const ChunkDesc* __restrict__ chunks,
        const double* __restrict__ sum, float* __restrict__ mu, float* __restrict__ invE, int Nc,
        int n, unsigned* __restrict__ k4bits, const unsigned* __restrict__ maxbits, int* __restrict__ zeroE,
        int ratio_mode) {
    const ChunkDesc cd = chunks[blockIdx.y];
    if (blockIdx.x >= cd.ntiles) return;
    const double mean = sum[blockIdx.y] / static_cast<double>(cd.L);
    // A window of constant data (a zero-filled gap, fillZeros=True) has zero energy, but S2 - S1^2/n of
    // the centred samples only cancels to ~1e-16 n max^2.  Below 1e-13 n max^2 (dead data: a window
    // standard deviation under 3e-7 of the chunk's largest sample) the window counts as constant and
    // gets invE = +inf, the reference's sum(if1^2)/0 (detect.py:577; the infs are zeroed at :278-281).
    double etol = 0.0;
    if (zeroE) {
        const double mc = static_cast<double>(__uint_as_float(maxbits[blockIdx.y])) + fabs(mean);
        etol = 1e-13 * static_cast<double>(n) * mc * mc;
    }
    bool any_zero = false;
    const T* x = raw + cd.raw_off;
    const int t0 = blockIdx.x * TILE_T;
    const int tid = threadIdx.x, w = tid >> 5, l = tid & 31;
    __shared__ double sh1[8], sh2[8], sh4[8];
    __shared__ double base1, base2, base4;

    // window sums at t0 (t0 < T always holds for an existing tile)
    // (the fourth-power sums only feed the precision policy; k4bits == nullptr skips them)
    const bool want4 = k4bits != nullptr;
    double a1 = 0.0, a2 = 0.0, a4 = 0.0;
    {
        const long long o = static_cast<long long>(t0) * Nc;
        for (int j = tid; j < n; j += 256) {
            const double v = static_cast<double>(x[o + j]) - mean;
            a1 += v;
            a2 += v * v;
            a4 += (v * v) * (v * v);
        }
        for (int s = 16; s > 0; s >>= 1) {
            a1 += __shfl_xor_sync(0xffffffffu, a1, s);
            a2 += __shfl_xor_sync(0xffffffffu, a2, s);
            a4 += __shfl_xor_sync(0xffffffffu, a4, s);
        }
        if (l == 0) {
            sh1[w] = a1;
            sh2[w] = a2;
            sh4[w] = a4;
        }
        __syncthreads();
        if (tid == 0) {
            double b1 = 0, b2 = 0, b4 = 0;
            for (int i = 0; i < 8; ++i) {
                b1 += sh1[i];
                b2 += sh2[i];
                b4 += sh4[i];
            }
            base1 = b1;
            base2 = b2;
            base4 = b4;
        }
        __syncthreads();
    }
    // differences: d[i] moves the window from t0+i-1 to t0+i (i >= 1); thread owns 8 lags
    constexpr int PER = TILE_T / 256;
    double d1[PER], d2[PER], d4[PER];
    double r1 = 0.0, r2 = 0.0, r4 = 0.0;
#pragma unroll
    for (int k = 0; k < PER; ++k) {
        const int i = tid * PER + k;
        const int t = t0 + i;
        double e1 = 0.0, e2 = 0.0, e4 = 0.0;
        if (i >= 1 && t < cd.T) {
            const long long o = static_cast<long long>(t - 1) * Nc;
            for (int c = 0; c < Nc; ++c) {
                const double vin = static_cast<double>(x[o + n + c]) - mean;
                const double vout = static_cast<double>(x[o + c]) - mean;
                e1 += vin - vout;
                e2 += vin * vin - vout * vout;
                if (want4) e4 += (vin * vin) * (vin * vin) - (vout * vout) * (vout * vout);
            }
        }
        r1 += e1;
        r2 += e2;
        r4 += e4;
        d1[k] = r1;
        d2[k] = r2;
        d4[k] = r4;
    }
    // exclusive scan of per-thread totals across the block
    double p1 = r1, p2 = r2, p4 = r4;
    for (int s = 1; s < 32; s <<= 1) {
        const double q1 = __shfl_up_sync(0xffffffffu, p1, s);
        const double q2 = __shfl_up_sync(0xffffffffu, p2, s);
        const double q4 = __shfl_up_sync(0xffffffffu, p4, s);
        if (l >= s) {
            p1 += q1;
            p2 += q2;
            p4 += q4;
        }
    }
    __syncthreads();
    if (l == 31) {
        sh1[w] = p1;
        sh2[w] = p2;
        sh4[w] = p4;
    }
    __syncthreads();
    double o1 = p1 - r1, o2 = p2 - r2, o4 = p4 - r4;  // exclusive within warp
    for (int i = 0; i < w; ++i) {
        o1 += sh1[i];
        o2 += sh2[i];
        o4 += sh4[i];
    }
    const double nn = static_cast<double>(n);
    const double cn = (nn - 1.0) / nn;
    float* pm = mu + cd.norm_off + t0;
    float* pe = invE + cd.norm_off + t0;
    float k4 = 0.f;   // worst window of the tile: sum (x - chunk mean)^4 / (window energy)^2
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
            if (zeroE && E <= etol) { E = 0.0; any_zero = true; }
            fm = static_cast<float>(s1 / nn);
            fe = static_cast<float>(cn / E);  // E == 0 -> +inf, as the reference's x/0
            if (want4) {
                // ratio_mode: sum (x - chunk mean)^2 / window energy, the factor by which rounding the
                // operands of the projection is amplified in the normalised value (CCX screening band)
                const float r = ratio_mode ? static_cast<float>(s2 / E)
                                           : static_cast<float>((base4 + o4 + d4[k]) / (E * E));
                k4 = fmaxf(k4, r >= 0.f && r < 3e38f ? r : 3e38f);   // NaN / inf -> never 8-bit
            }
        }
        // mu is stored phase-major inside every 1024-lag group, [group][t % 8][t / 8 % 128], the
        // order K1's epilogue threads (one phase, consecutive accumulator columns) read it in
        pm[(i & ~1023) + (i & 7) * 128 + ((i & 1023) >> 3)] = fm;
        pe[i] = fe;
    }
    if (zeroE && __any_sync(0xffffffffu, any_zero) && l == 0) atomicOr(&zeroE[blockIdx.y], 1);
    if (want4) {
        for (int s = 16; s > 0; s >>= 1) k4 = fmaxf(k4, __shfl_xor_sync(0xffffffffu, k4, s));
        if (l == 0) atomicMax(&k4bits[blockIdx.y], __float_as_uint(k4));  // k4 >= 0: uint order == float order
    }
}

void launch_k0(const void* raw, int dtype_f32, const ChunkDesc* d_chunks, int nchunks, int Nc, int n,
               int max_Lpad, int max_ntiles, double* d_sum, unsigned* d_maxbits, float* d_scale,
               __half* d_xsplit, float* d_mu, float* d_invE, int x8_policy, float k4_limit,
               unsigned* d_k4bits, int* d_chunk_mode, int* d_zeroE, int* d_chunk_bad, cudaStream_t st) {
    cudaMemsetAsync(d_sum, 0, sizeof(double) * nchunks, st);
    if (d_zeroE) cudaMemsetAsync(d_zeroE, 0, sizeof(int) * nchunks, st);
    cudaMemsetAsync(d_maxbits, 0, sizeof(unsigned) * nchunks, st);
    cudaMemsetAsync(d_k4bits, 0, sizeof(unsigned) * nchunks, st);
    unsigned* k4 = (x8_policy == X8_AUTO || x8_policy == X8_RATIO) ? d_k4bits : nullptr;   // k0_norm runs before k0_split reads it
    const int ratio_mode = x8_policy == X8_RATIO;
    const dim3 g1(64, nchunks);
    long long tot = static_cast<long long>(Nc) * max_Lpad;
    int gx = static_cast<int>((tot + 256 * 8 - 1) / (256 * 8));
    if (gx < 1) gx = 1;
    const dim3 g2(gx, nchunks);
    const dim3 g3(max_ntiles, nchunks);
    if (dtype_f32) {
        const float* r = static_cast<const float*>(raw);
        k0_stats<float><<<g1, 256, 0, st>>>(r, d_chunks, d_sum, d_maxbits);
        k0_norm<float><<<g3, 256, 0, st>>>(r, d_chunks, d_sum, d_mu, d_invE, Nc, n, k4, d_maxbits, d_zeroE, ratio_mode);
        k0_split<float><<<g2, 256, 0, st>>>(r, d_chunks, d_sum, d_maxbits, d_scale, d_xsplit, Nc, x8_policy,
                                            k4_limit, d_k4bits, d_chunk_mode, d_chunk_bad);
    } else {
        const double* r = static_cast<const double*>(raw);
        k0_stats<double><<<g1, 256, 0, st>>>(r, d_chunks, d_sum, d_maxbits);
        k0_norm<double><<<g3, 256, 0, st>>>(r, d_chunks, d_sum, d_mu, d_invE, Nc, n, k4, d_maxbits, d_zeroE, ratio_mode);
        k0_split<double><<<g2, 256, 0, st>>>(r, d_chunks, d_sum, d_maxbits, d_scale, d_xsplit, Nc, x8_policy,
                                             k4_limit, d_k4bits, d_chunk_mode, d_chunk_bad);
    }
}

// DS = +inf at the zero-energy lags of the flagged chunks (one block per K1 tile of TILE_T lags).
__global__ void __launch_bounds__(256)
zero_energy_fix_kernel(const ChunkDesc* __restrict__ chunks, int nrows, const float* __restrict__ invE,
                       const int* __restrict__ zeroE, float* __restrict__ DS, double* __restrict__ DS64) {
    if (!zeroE[blockIdx.y]) return;
    const ChunkDesc cd = chunks[blockIdx.y];
    if (blockIdx.x >= cd.ntiles) return;
    const int t0 = blockIdx.x * TILE_T;
    for (int i = threadIdx.x; i < TILE_T; i += 256) {
        const int t = t0 + i;
        if (t >= cd.T || !isinf(invE[cd.norm_off + t])) continue;
        for (int s = 0; s < nrows; ++s) {
            const long long o = cd.ds_off + static_cast<long long>(s) * cd.Tpad + t;
            if (DS) DS[o] = INFINITY;
            if (DS64) DS64[o] = static_cast<double>(INFINITY);
        }
    }
}

void launch_zero_energy_fix(const ChunkDesc* d_chunks, int nchunks, int max_ntiles, int nrows, const float* d_invE,
                            const int* d_zeroE, float* d_DS, double* d_DS64, cudaStream_t st) {
    const dim3 g(max_ntiles, nchunks);
    zero_energy_fix_kernel<<<g, 256, 0, st>>>(d_chunks, nrows, d_invE, d_zeroE, d_DS, d_DS64);
}

// dst row += src row, in the order of the list (pieces of a rank > 16 subspace, see launch_sum_pieces)
__global__ void __launch_bounds__(256)
sum_pieces_kernel(const ChunkDesc* __restrict__ chunks, const PieceSum* __restrict__ pieces, int npieces,
                  float* __restrict__ DS) {
    const ChunkDesc cd = chunks[blockIdx.y];
    const int i = (blockIdx.x * 256 + threadIdx.x) * 4;
    if (i >= cd.Tpad) return;
    // pieces are grouped by dst_row; a thread owns 4 lags of every row, so the order is fixed
    for (int k = 0; k < npieces; ++k) {
        float4* d = reinterpret_cast<float4*>(DS + cd.ds_off + static_cast<long long>(pieces[k].dst_row) * cd.Tpad + i);
        const float4 a = *d;
        const float4 b = *reinterpret_cast<const float4*>(DS + cd.ds_off +
                                                           static_cast<long long>(pieces[k].src_row) * cd.Tpad + i);
        *d = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
    }
}

void launch_sum_pieces(const ChunkDesc* d_chunks, int nchunks, int max_Tpad, const PieceSum* d_pieces, int npieces,
                       float* d_DS, cudaStream_t st) {
    const dim3 g((max_Tpad / 4 + 255) / 256, nchunks);
    sum_pieces_kernel<<<g, 256, 0, st>>>(d_chunks, d_pieces, npieces, d_DS);
}

}  // namespace dtx
