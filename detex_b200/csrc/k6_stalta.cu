// k6_stalta.cu -- classic STA/LTA screen of raw continuous data, `fas._checkSTALTA`
// (reference detex/fas.py:175-205): cft = classic_sta_lta(z, nsta, nlta); pass if
// max(cft) <= limit.  The arithmetic lives in ObsPy 1.0.2 (not vendored in the reference:
// obspy/signal/src/stalta.c == trigger.classic_sta_lta_py):
//   cft[i] = (sum_{i-nsta<j<=i} x_j^2 / nsta) / (sum_{i-nlta<j<=i} x_j^2 / nlta), i >= nlta-1
//   cft[i] = 0 for i < nlta-1;  lta clamped to the smallest positive double.
// One CTA per 4096 samples of one chunk's screening channel: window sums at the tile start
// by reduction, then a float64 block scan of the in/out differences.  HBM-bound: reads the
// channel once (strided by Nc in the multiplexed chunk).
#include <cfloat>
#include "dtx_kernels.cuh"

namespace dtx {
namespace {

constexpr int SLT = 4096;

template <typename T>
__global__ void __launch_bounds__(256)
stalta_max_kernel(const T* __restrict__ raw, const long long* __restrict__ raw_off, const int* __restrict__ Ls_arr,
                  int Nc, int chan, int nsta, int nlta, unsigned* __restrict__ out_bits) {
    const int Ls = Ls_arr[blockIdx.y];
    const int i0 = nlta - 1 + blockIdx.x * SLT;  // first output index of this tile
    if (i0 >= Ls) return;
    const T* x = raw + raw_off[blockIdx.y] + chan;
    const int tid = threadIdx.x, w = tid >> 5, l = tid & 31;
    __shared__ double sh[8][2];
    __shared__ double base[2];
    auto sq = [&](int j) { const double v = static_cast<double>(x[static_cast<long long>(j) * Nc]); return v * v; };
    // sums of the windows ending at i0
    double a = 0.0, b = 0.0;
    for (int j = tid; j < nlta; j += 256) {
        const double v = sq(i0 - j);
        b += v;
        if (j < nsta) a += v;
    }
    for (int s = 16; s > 0; s >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, s);
        b += __shfl_xor_sync(0xffffffffu, b, s);
    }
    if (l == 0) { sh[w][0] = a; sh[w][1] = b; }
    __syncthreads();
    if (tid == 0) {
        double s0 = 0, s1 = 0;
        for (int i = 0; i < 8; ++i) { s0 += sh[i][0]; s1 += sh[i][1]; }
        base[0] = s0; base[1] = s1;
    }
    __syncthreads();
    constexpr int PER = SLT / 256;
    double d0[PER], d1[PER], r0 = 0.0, r1 = 0.0;
#pragma unroll
    for (int k = 0; k < PER; ++k) {
        const int i = i0 + tid * PER + k;
        double e0 = 0.0, e1 = 0.0;
        if (tid * PER + k >= 1 && i < Ls) {
            const double v = sq(i);
            e0 = v - sq(i - nsta);
            e1 = v - sq(i - nlta);
        }
        r0 += e0; r1 += e1;
        d0[k] = r0; d1[k] = r1;
    }
    double p0 = r0, p1 = r1;
    for (int s = 1; s < 32; s <<= 1) {
        const double q0 = __shfl_up_sync(0xffffffffu, p0, s), q1 = __shfl_up_sync(0xffffffffu, p1, s);
        if (l >= s) { p0 += q0; p1 += q1; }
    }
    __syncthreads();
    if (l == 31) { sh[w][0] = p0; sh[w][1] = p1; }
    __syncthreads();
    double o0 = p0 - r0, o1 = p1 - r1;
    for (int i = 0; i < w; ++i) { o0 += sh[i][0]; o1 += sh[i][1]; }
    const double frac = static_cast<double>(nsta) / static_cast<double>(nlta);
    float mx = 0.f;
#pragma unroll
    for (int k = 0; k < PER; ++k) {
        const int i = i0 + tid * PER + k;
        if (i < Ls) {
            const double sta = base[0] + o0 + d0[k];
            double lta = base[1] + o1 + d1[k];
            if (lta < DBL_MIN) lta = DBL_MIN;
            mx = fmaxf(mx, static_cast<float>(sta / frac / lta));
        }
    }
    for (int s = 16; s > 0; s >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, s));
    if (l == 0) atomicMax(&out_bits[blockIdx.y], __float_as_uint(mx));
}

}  // namespace

void launch_stalta_max(const void* raw, int dtype_f32, const long long* d_raw_off, const int* d_Ls, int nchunks,
                       int maxLs, int Nc, int chan, int nsta, int nlta, unsigned* d_out_bits, cudaStream_t st) {
    cudaMemsetAsync(d_out_bits, 0, sizeof(unsigned) * nchunks, st);
    const int ntile = (maxLs - (nlta - 1) + SLT - 1) / SLT;
    if (ntile < 1) return;
    const dim3 grid(ntile, nchunks);
    if (dtype_f32)
        stalta_max_kernel<float><<<grid, 256, 0, st>>>(static_cast<const float*>(raw), d_raw_off, d_Ls, Nc, chan, nsta,
                                                       nlta, d_out_bits);
    else
        stalta_max_kernel<double><<<grid, 256, 0, st>>>(static_cast<const double*>(raw), d_raw_off, d_Ls, Nc, chan,
                                                        nsta, nlta, d_out_bits);
}

}  // namespace dtx
