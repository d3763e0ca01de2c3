// k4_ccx.cu -- K4: pairwise maximum normalised cross-correlation of multiplexed event
// waveforms; replaces construct._makeDFcclags / _CCX2 / _subSamp
// (reference detex/construct.py:369-466).
//
// Closed form (checked against the reference to 1e-16): with ns = n/Nc,
// trunc = n//(2Nc) - 1, for m = 0 .. nl-1 (nl = 2ns-1-2trunc) and per-channel shift
// kappa = m + trunc + 1 - ns,
//   res[m] = ( sum_c sum_j x1_c[j] x2_c[j+kappa]  -  sum(x1) a_c[m] ) / ( n b_c[m] std(x1) )
// where x2 is zero outside [0, ns) and a, b are mean and population std of the zero-padded
// length-n window of x2.  Then nanmax / nanargmax, the |res|>1 zeroing rule, the cosine-fit
// sub-sample shift, lag = (argmax + 1 + trunc) Nc - n.
//
// This first engine evaluates the correlations in float64 on the FP64 pipe (register
// window over 4 lags per thread, de-interleaved smem so the sliding read is conflict
// free).  The per-event window statistics are computed once per event by ccx_stats.
#include <algorithm>
#include <cstdlib>

#include "dtx_kernels.cuh"
#include "tc_common.cuh"

// unroll factors of the re-scoring loops (experiment knobs, profiles/r02_ccx_rescoring.md)
#ifndef DTX_CCX_UNROLL3
#define DTX_CCX_UNROLL3 3
#endif
#ifndef DTX_CCX_UNROLL4
#define DTX_CCX_UNROLL4 4
#endif
#define DTX_PRAGMA_(x) _Pragma(#x)
#define DTX_UNROLL(n) DTX_PRAGMA_(unroll n)

namespace dtx {
namespace {

constexpr int CT = 256;          // threads
constexpr int LT = 4 * CT;       // lags per tile (4 per thread)
constexpr int JT = 256;          // taps per tile
constexpr int X2ROW = (JT + LT) / 4 + 2;
constexpr double CC_GUARD = 1e-9;

template <typename T>
__global__ void __launch_bounds__(CT)
ccx_stats(const T* __restrict__ X, int N, int n, int Nc, int trunc, int nl,
          double* __restrict__ wa, double* __restrict__ wb, double* __restrict__ evsum,
          double* __restrict__ evstd) {
    extern __shared__ double sm[];  // P1[Nc][ns+1], P2[Nc][ns+1]
    const int ev = blockIdx.x;
    const int ns = n / Nc;
    const T* x = X + static_cast<long long>(ev) * n;
    double* P1 = sm;
    double* P2 = sm + static_cast<size_t>(Nc) * (ns + 1);
    if (threadIdx.x < Nc) {
        const int c = threadIdx.x;
        double s1 = 0, s2 = 0;
        P1[c * (ns + 1)] = 0;
        P2[c * (ns + 1)] = 0;
        for (int i = 0; i < ns; ++i) {
            const double v = static_cast<double>(x[static_cast<long long>(i) * Nc + c]);
            s1 += v;
            s2 += v * v;
            P1[c * (ns + 1) + i + 1] = s1;
            P2[c * (ns + 1) + i + 1] = s2;
        }
    }
    __syncthreads();
    const double nn = static_cast<double>(n);
    if (threadIdx.x == 0) {
        double s1 = 0, s2 = 0;
        for (int c = 0; c < Nc; ++c) {
            s1 += P1[c * (ns + 1) + ns];
            s2 += P2[c * (ns + 1) + ns];
        }
        const double mean = s1 / nn;
        double var = 0;  // np.std: two-pass population variance
        for (int i = 0; i < n; ++i) {
            const double d = static_cast<double>(x[i]) - mean;
            var += d * d;
        }
        evsum[ev] = s1;
        evstd[ev] = sqrt(var / nn);
        (void)s2;
    }
    for (int m = threadIdx.x; m < nl; m += CT) {
        const int kappa = m + trunc + 1 - ns;
        const int lo = max(0, kappa), hi = min(ns, ns + kappa);
        double s1 = 0, s2 = 0;
        if (hi > lo)
            for (int c = 0; c < Nc; ++c) {
                s1 += P1[c * (ns + 1) + hi] - P1[c * (ns + 1) + lo];
                s2 += P2[c * (ns + 1) + hi] - P2[c * (ns + 1) + lo];
            }
        const double a = s1 / nn;
        double var = s2 / nn - a * a;
        if (var < 0) var = 0;
        wa[static_cast<long long>(ev) * nl + m] = a;
        wb[static_cast<long long>(ev) * nl + m] = sqrt(var);
    }
}

struct MaxLoc {
    double mx, mn;
    int imx;
    int cnt;
};

__device__ __forceinline__ void ml_merge(MaxLoc& a, const MaxLoc& b) {
    if (b.cnt) {
        if (!a.cnt || b.mx > a.mx || (b.mx == a.mx && b.imx < a.imx)) {
            a.mx = b.mx;
            a.imx = b.imx;
        }
        if (!a.cnt || b.mn < a.mn) a.mn = b.mn;
        a.cnt += b.cnt;
    }
}

__device__ MaxLoc block_maxloc(const double* res, int nl, MaxLoc* sh) {
    MaxLoc m{0, 0, 0, 0};
    for (int i = threadIdx.x; i < nl; i += CT) {
        const double v = res[i];
        if (!isnan(v)) {
            MaxLoc o{v, v, i, 1};
            ml_merge(m, o);
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        MaxLoc t;
        t.mx = __shfl_xor_sync(0xffffffffu, m.mx, o);
        t.mn = __shfl_xor_sync(0xffffffffu, m.mn, o);
        t.imx = __shfl_xor_sync(0xffffffffu, m.imx, o);
        t.cnt = __shfl_xor_sync(0xffffffffu, m.cnt, o);
        ml_merge(m, t);
    }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = m;
    __syncthreads();
    MaxLoc r = sh[0];
    for (int i = 1; i < CT / 32; ++i) ml_merge(r, sh[i]);
    __syncthreads();
    return r;
}

template <typename T>
__global__ void __launch_bounds__(CT)
ccx_kernel(const T* __restrict__ X, int N, int n, int Nc, int trunc, int nl, int row_begin,
           const int* __restrict__ rows, const double* __restrict__ wa, const double* __restrict__ wb,
           const double* __restrict__ evsum, const double* __restrict__ evstd,
           double* __restrict__ cc, int* __restrict__ lag, double* __restrict__ sub,
           const int2* __restrict__ pairs, const int* __restrict__ npairs, int pair_cap) {
    extern __shared__ double res[];  // [nl]
    __shared__ double x1t[JT];
    __shared__ double x2t[4][X2ROW];
    __shared__ MaxLoc shml[CT / 32];
    const int ns = n / Nc;
    const int tid = threadIdx.x;
    const double nn = static_cast<double>(n);
    // two ways to enumerate the pairs: every c > b of the template row blockIdx.y, or (pairs != null) a
    // device-side list of (slot r, event c) entries -- the degenerate pairs the tensor-core engine
    // hands over -- walked with a grid stride
    const int npr = pairs ? min(*npairs, pair_cap) : 0;
    int b = 0, orow = 0, c = 0, it = blockIdx.x;
    if (!pairs) {
        b = rows ? rows[blockIdx.y] : row_begin + blockIdx.y;
        orow = blockIdx.y;
        c = b + 1 + blockIdx.x;
    }
    for (;; ) {
        if (pairs) {
            if (it >= npr) break;
            orow = pairs[it].x;
            c = pairs[it].y;
            b = rows[orow];
            it += gridDim.x;
        } else if (c >= N) break;
        const T* x1 = X + static_cast<long long>(b) * n;
        const double sum1 = evsum[b], std1 = evstd[b];
        const T* x2 = X + static_cast<long long>(c) * n;
        for (int m0 = 0; m0 < nl; m0 += LT) {
            double acc[4] = {0, 0, 0, 0};
            const int kappa0 = m0 + trunc + 1 - ns;  // shift of lag m0
            for (int ch = 0; ch < Nc; ++ch)
                for (int j0 = 0; j0 < ns; j0 += JT) {
                    __syncthreads();
                    for (int i = tid; i < JT; i += CT) {
                        const int j = j0 + i;
                        x1t[i] = j < ns ? static_cast<double>(x1[static_cast<long long>(j) * Nc + ch]) : 0.0;
                    }
                    for (int i = tid; i < JT + LT; i += CT) {
                        const int j = j0 + kappa0 + i;
                        const double v =
                            (j >= 0 && j < ns) ? static_cast<double>(x2[static_cast<long long>(j) * Nc + ch]) : 0.0;
                        x2t[i & 3][i >> 2] = v;
                    }
                    __syncthreads();
                    // thread owns lags m0 + 4*tid + {0,1,2,3}: window w[q] = x2t[jj + 4*tid + q]
                    double w0 = x2t[0][tid], w1 = x2t[1][tid], w2 = x2t[2][tid];
#pragma unroll 4
                    for (int jj = 0; jj < JT; ++jj) {
                        const int e = jj + 3;
                        const double w3 = x2t[e & 3][(e >> 2) + tid];
                        const double a = x1t[jj];
                        acc[0] = fma(a, w0, acc[0]);
                        acc[1] = fma(a, w1, acc[1]);
                        acc[2] = fma(a, w2, acc[2]);
                        acc[3] = fma(a, w3, acc[3]);
                        w0 = w1; w1 = w2; w2 = w3;
                    }
                }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int m = m0 + 4 * tid + q;
                if (m < nl) {
                    const double a = wa[static_cast<long long>(c) * nl + m];
                    const double sb = wb[static_cast<long long>(c) * nl + m];
                    res[m] = (acc[q] - sum1 * a) / (nn * sb * std1);
                }
            }
        }
        __syncthreads();
        // ---- _CCX2 tail (construct.py:453-466)
        MaxLoc ml = block_maxloc(res, nl, shml);
        double maxcc = 0.0, ss = 0.0;
        int lg = 0;
        if (ml.cnt > 0) {
            // `if maxcc > 1. or mincc < -1.` (construct.py:457): meant for the infs of zeroed-out
            // waveforms.  A Pearson coefficient of exactly 1 can come out as 1 + 2e-16 depending on
            // summation order, so the comparison carries a round-off guard band.
            if (ml.mx > 1.0 + CC_GUARD || ml.mn < -1.0 - CC_GUARD) {
                for (int i = tid; i < nl; i += CT) {
                    const double v = res[i];
                    if (v > 1.0 + CC_GUARD || v < -1.0 - CC_GUARD) res[i] = 0.0;
                }
                __syncthreads();
                ml = block_maxloc(res, nl, shml);
            }
            maxcc = ml.mx;
            const int ind = ml.imx;
            lg = (ind + 1 + trunc) * Nc - n;
            if (ind == 0 || ind == nl - 1) {
                ss = 0.0;
            } else {
                const double cb4 = res[ind - 1], caf = res[ind + 1], cn = res[ind];
                const double alpha = acos((cb4 + caf) / (2 * cn));
                const double alsi = sin(alpha);
                const double tau = -(atan((cb4 - caf) / (2 * cn * alsi)) / alpha);
                ss = (fabs(tau) > 0.5) ? static_cast<double>(ind) : tau;  // reference quirk, :418-421
            }
        }
        if (tid == 0) {
            const long long o = static_cast<long long>(orow) * N + c;
            cc[o] = maxcc;
            lag[o] = lg;
            sub[o] = ss;
        }
        __syncthreads();
        if (!pairs) c += gridDim.x;
    }
}


// ===================================================================== tensor-core engine
// Events as rank-1 templates of the Hankel GEMM (k1_project.cu, MODE 1), zero-padded events
// as its "chunks".  The GEMM returns the float32 correlation series of every pair; this
// kernel finds the arg-max neighbourhood in it and RE-SCORES those few lags in float64 from
// the original waveforms, so cc / lag / subsamp come out with float64 accuracy.

// padded "chunk" of event c: per channel [P zeros | x_c | zeros], P = ns - trunc - 1, so that
// lag m of the Hankel output is the shift kappa = m + trunc + 1 - ns of _CCX2.
template <typename T>
__global__ void __launch_bounds__(256)
ccx_pad_kernel(const T* __restrict__ X, int n, int Nc, int c0, int P, int Lc, double* __restrict__ out) {
    const int c = c0 + blockIdx.y;
    const long long Lm = static_cast<long long>(Lc) * Nc;
    const T* x = X + static_cast<long long>(c) * n;
    double* o = out + static_cast<long long>(blockIdx.y) * Lm;
    for (long long i = blockIdx.x * 256 + threadIdx.x; i < Lm; i += static_cast<long long>(gridDim.x) * 256) {
        const long long j = i - static_cast<long long>(P) * Nc;
        o[i] = (j >= 0 && j < n) ? static_cast<double>(x[j]) : 0.0;
    }
}

template <typename T>
__device__ __forceinline__ double ccx_exact(const T* __restrict__ x1, const T* __restrict__ x2, int n, int Nc,
                                            int kappa, double sum1, double a, double sb, double std1, int lane) {
    const int sh = kappa * Nc;
    double acc = 0.0;
    const int lo = max(0, -sh), hi = min(n, n - sh);
    for (int i = lo + lane; i < hi; i += 32) acc = fma(static_cast<double>(x1[i]), static_cast<double>(x2[i + sh]), acc);
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    return (acc - sum1 * a) / (static_cast<double>(n) * sb * std1);
}

// Lags kappa-1, kappa, kappa+1 in ONE pass over x1 (the arg-max and the two neighbours the cosine fit
// needs): a third of the L2 traffic of three ccx_exact calls.  acc[j] = sum_i x1[i] * x2[i + (kappa-1+j) Nc].
template <typename T>
__device__ __forceinline__ void ccx_dot3(const T* __restrict__ x1, const T* __restrict__ x2, int n, int Nc,
                                         int kappa, int lane, double acc[3]) {
    const int sh = kappa * Nc;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0;
    const int lo = max(0, -(sh + Nc)), hi = min(n, n - (sh - Nc));   // union of the three index ranges
    for (int i = lo + lane; i < hi; i += 32) {
        const double v = static_cast<double>(x1[i]);
        const int j0 = i + sh - Nc, j1 = i + sh, j2 = i + sh + Nc;
        if (j0 >= 0 && j0 < n) a0 = fma(v, static_cast<double>(x2[j0]), a0);
        if (j1 >= 0 && j1 < n) a1 = fma(v, static_cast<double>(x2[j1]), a1);
        if (j2 >= 0 && j2 < n) a2 = fma(v, static_cast<double>(x2[j2]), a2);
    }
    for (int o = 16; o > 0; o >>= 1) {
        a0 += __shfl_xor_sync(0xffffffffu, a0, o);
        a1 += __shfl_xor_sync(0xffffffffu, a1, o);
        a2 += __shfl_xor_sync(0xffffffffu, a2, o);
    }
    acc[0] = a0; acc[1] = a1; acc[2] = a2;
}

constexpr int CCX_MAXCAND = 8;
// Candidate band of a pair: every lag whose float32 series value is within `band` of the series maximum
// is re-scored in float64, so the exact maximum is found as long as band >= 2 * (error of the series).
// band0 is that bound for operands of unit amplification (3e-5 for the fp16x3 series, 2.2e-3 for the
// one-MMA screening series); rounding an operand moves the dot product by <= eps * |u| * |w|, i.e. the
// normalised value by eps * (|x1| / |x1 - mean|) * (|w| / |w - mean(w)|): a1 from the template's
// statistics, a2 = the worst window of the padded signal (k0_norm, ratio mode).
__device__ __forceinline__ float ccx_band(float band0, double sum1, double std1, int n, unsigned ratio_bits) {
    const double mean1 = sum1 / n;
    const float a1 = sqrtf(1.f + static_cast<float>((mean1 * mean1) / (std1 * std1)));
    const float a2 = sqrtf(fmaxf(1.f, __uint_as_float(ratio_bits)));
    const float b = band0 * a1 * a2;
    return b == b ? b : INFINITY;
}

// one warp per (signal chunk ci, template row r)
template <typename T>
__global__ void __launch_bounds__(256)
ccx_post_kernel(const float* __restrict__ DS, const ChunkDesc* __restrict__ chunks, int c0, const T* __restrict__ X,
                int N, int n, int Nc, int trunc, int nl, const int* __restrict__ rows, int nrows,
                const double* __restrict__ wa, const double* __restrict__ wb, const double* __restrict__ evsum,
                const double* __restrict__ evstd, double* __restrict__ cc, int* __restrict__ lag,
                double* __restrict__ sub, int* __restrict__ nflag, int2* __restrict__ flagged, int flag_cap,
                const int4* __restrict__ karg, const unsigned* __restrict__ ratio_bits, float band0) {
    const int ci = blockIdx.y;
    const int c = c0 + ci;
    const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= nrows) return;
    const int b = rows[r];
    if (b >= c) return;
    // with the tiled re-scoring this kernel only takes the pairs the scan left for it (several
    // near-maxima, a maximum at either end of the lag range, an out-of-range series)
    if (karg && karg[static_cast<long long>(ci) * nrows + r].x != -1) return;
    const ChunkDesc cd = chunks[ci];
    const float* row = DS + cd.ds_off + static_cast<long long>(r) * cd.Tpad;
    const long long o = static_cast<long long>(r) * N + c;
    const double std1 = evstd[b], std2 = evstd[c];
    if (!(std1 > 0.0) || !(std2 > 0.0)) {  // zeroed-out waveform: the reference's all-NaN branch
        if (lane == 0) { cc[o] = 0.0; lag[o] = 0; sub[o] = 0.0; }
        return;
    }
    float mx = -INFINITY, mn = INFINITY;
    int cnt = 0;
    for (int m = lane; m < nl; m += 32) {
        const float v = row[m];
        if (!isnan(v)) { mx = fmaxf(mx, v); mn = fminf(mn, v); ++cnt; }
    }
    for (int s = 16; s > 0; s >>= 1) {
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, s));
        mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, s));
        cnt += __shfl_xor_sync(0xffffffffu, cnt, s);
    }
    if (cnt == 0) {
        if (lane == 0) { cc[o] = 0.0; lag[o] = 0; sub[o] = 0.0; }
        return;
    }
    const float band = ccx_band(band0, evsum[b], std1, n, ratio_bits[ci]);
    const float lim = 1.f + fmaxf(1e-3f, band);       // beyond: zero-variance windows (inf), not round-off
    bool fallback = (mx > lim) || (mn < -lim);
    // candidates: every lag whose float32 value is within the band of the maximum
    int cand[CCX_MAXCAND];
    int ncand = 0;
    if (!fallback) {
        const float thr = mx - band;
        for (int m0 = 0; m0 < nl && !fallback; m0 += 32) {
            const int m = m0 + lane;
            const bool hit = m < nl && row[m] >= thr;
            unsigned bal = __ballot_sync(0xffffffffu, hit);
            while (bal) {
                const int l = __ffs(bal) - 1;
                bal &= bal - 1;
                if (ncand < CCX_MAXCAND) cand[ncand++] = m0 + l;
                else fallback = true;
            }
        }
    }
    const T* x1 = X + static_cast<long long>(b) * n;
    const T* x2 = X + static_cast<long long>(c) * n;
    const int ns = n / Nc;
    const double sum1 = evsum[b];
    const double* wac = wa + static_cast<long long>(c) * nl;
    const double* wbc = wb + static_cast<long long>(c) * nl;
    double best = 0.0;
    int ind = -1;
    double cb4 = 0.0, caf = 0.0;
    bool have_nb = false;
    const double dn = static_cast<double>(n);
    if (!fallback && ncand == 1 && cand[0] > 0 && cand[0] < nl - 1) {
        // the common case: one clear maximum away from the ends -- it and its neighbours in one pass
        const int m = cand[0];
        double acc[3];
        ccx_dot3(x1, x2, n, Nc, m + trunc + 1 - ns, lane, acc);
        const double v = (acc[1] - sum1 * wac[m]) / (dn * wbc[m] * std1);
        if (!isnan(v)) {
            best = v; ind = m;
            cb4 = (acc[0] - sum1 * wac[m - 1]) / (dn * wbc[m - 1] * std1);
            caf = (acc[2] - sum1 * wac[m + 1]) / (dn * wbc[m + 1] * std1);
            have_nb = true;
        }
        if (ind < 0 || best > 1.0 + CC_GUARD) fallback = true;
    } else if (!fallback) {
        for (int k = 0; k < ncand; ++k) {
            const int m = cand[k];
            const double v = ccx_exact(x1, x2, n, Nc, m + trunc + 1 - ns, sum1, wac[m], wbc[m], std1, lane);
            if (isnan(v)) continue;
            if (ind < 0 || v > best) { best = v; ind = m; }
        }
        if (ind < 0 || best > 1.0 + CC_GUARD) fallback = true;
    }
    if (fallback) {
        if (lane == 0) {
            const int k = atomicAdd(nflag, 1);
            if (k < flag_cap) flagged[k] = make_int2(r, c);   // (slot, event c)
        }
        return;
    }
    double ss = 0.0;
    if (ind != 0 && ind != nl - 1) {
        if (!have_nb) {
            cb4 = ccx_exact(x1, x2, n, Nc, ind - 1 + trunc + 1 - ns, sum1, wac[ind - 1], wbc[ind - 1], std1, lane);
            caf = ccx_exact(x1, x2, n, Nc, ind + 1 + trunc + 1 - ns, sum1, wac[ind + 1], wbc[ind + 1], std1, lane);
        }
        const double alpha = acos((cb4 + caf) / (2 * best));
        const double alsi = sin(alpha);
        const double tau = -(atan((cb4 - caf) / (2 * best * alsi)) / alpha);
        ss = (fabs(tau) > 0.5) ? static_cast<double>(ind) : tau;
    }
    if (lane == 0) {
        cc[o] = best;
        lag[o] = (ind + 1 + trunc) * Nc - n;
        sub[o] = ss;
    }
}


// Scan of the float32 correlation series of every pair (one warp per pair, many warps per SM, 128-bit
// loads: the 4 KB rows stream from HBM / L2).  Per pair an int4 record:
//   x >= 0 : the lags (ascending, up to 4, unused = -1) whose value lies within the pair's band of the
//            series maximum -- the tiled kernel scores them in float64, keeps the first largest, and
//            re-scores its two neighbours for the cosine fit;
//   x = -1 : more than 4 such lags or an out-of-range series (zero-variance windows): ccx_post_kernel;
//   x = -2 : nothing to do (b >= c, or a zeroed-out waveform, whose result (0, 0, 0) is written here).
constexpr int SCAN_MAXC = 4;
__global__ void __launch_bounds__(256)
ccx_scan_kernel(const float* __restrict__ DS, const ChunkDesc* __restrict__ chunks, int c0, int N, int nl,
                const int* __restrict__ rows, int nrows, const double* __restrict__ evsum,
                const double* __restrict__ evstd, int n, const unsigned* __restrict__ ratio_bits, float band0,
                double* __restrict__ cc, int* __restrict__ lag, double* __restrict__ sub, int4* __restrict__ karg) {
    const int ci = blockIdx.y;
    const int c = c0 + ci;
    const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= nrows) return;
    const int b = rows[r];
    int4* out = karg + static_cast<long long>(ci) * nrows + r;
    if (b >= c) {
        if (lane == 0) *out = make_int4(-2, -1, -1, -1);
        return;
    }
    const long long o = static_cast<long long>(r) * N + c;
    if (!(evstd[b] > 0.0) || !(evstd[c] > 0.0)) {        // zeroed-out waveform: the reference's all-NaN branch
        if (lane == 0) { cc[o] = 0.0; lag[o] = 0; sub[o] = 0.0; *out = make_int4(-2, -1, -1, -1); }
        return;
    }
    const ChunkDesc cd = chunks[ci];
    const float4* row4 = reinterpret_cast<const float4*>(DS + cd.ds_off + static_cast<long long>(r) * cd.Tpad);
    const int nq = (nl + 3) / 4;                         // rows are padded to a multiple of TILE_T floats
    // rows of up to 1024 lags (one MODE-1 tile: every CCX of n <= 2048 per channel pair) are read ONCE, all
    // eight 128-bit loads of a lane in flight together, and both passes run on registers
    constexpr int SCAN_REG = 8;
    const bool inreg = nq <= 32 * SCAN_REG;
    float4 vr[SCAN_REG];
    if (inreg) {
#pragma unroll
        for (int i = 0; i < SCAN_REG; ++i) {
            const int q = lane + 32 * i;
            vr[i] = q < nq ? __ldcs(row4 + q) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    float mx = -INFINITY, mn = INFINITY;
    int cnt = 0;
    auto stat4 = [&](const float4& v4, int q) {
        const float v[4] = {v4.x, v4.y, v4.z, v4.w};
        if (4 * q + 3 < nl) {                            // a whole quad: fmaxf / fminf skip NaN by themselves
#pragma unroll
            for (int e = 0; e < 4; ++e) { mx = fmaxf(mx, v[e]); mn = fminf(mn, v[e]); cnt += (v[e] == v[e]); }
        } else {
#pragma unroll
            for (int e = 0; e < 4; ++e)
                if (4 * q + e < nl && !isnan(v[e])) { mx = fmaxf(mx, v[e]); mn = fminf(mn, v[e]); ++cnt; }
        }
    };
    if (inreg) {
#pragma unroll
        for (int i = 0; i < SCAN_REG; ++i)
            if (lane + 32 * i < nq) stat4(vr[i], lane + 32 * i);
    } else {
        for (int q = lane; q < nq; q += 32) stat4(__ldcs(row4 + q), q);
    }
    for (int s = 16; s > 0; s >>= 1) {
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, s));
        mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, s));
        cnt += __shfl_xor_sync(0xffffffffu, cnt, s);
    }
    if (cnt == 0) {
        if (lane == 0) { cc[o] = 0.0; lag[o] = 0; sub[o] = 0.0; *out = make_int4(-2, -1, -1, -1); }
        return;
    }
    const float band = ccx_band(band0, evsum[b], evstd[b], n, ratio_bits[ci]);
    const float lim = 1.f + fmaxf(1e-3f, band);          // beyond: zero-variance windows (inf), not round-off
    int cand[SCAN_MAXC] = {-1, -1, -1, -1};
    int nc = 0;
    bool slow = mx > lim || mn < -lim;
    if (!slow) {
        const float thr = mx - band;
        auto hits = [&](const float4& v4, int q) -> unsigned {
            const float v[4] = {v4.x, v4.y, v4.z, v4.w};
            unsigned m4 = 0;
            if (fmaxf(fmaxf(v[0], v[1]), fmaxf(v[2], v[3])) >= thr) {   // rare: a handful of quads per row
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    if (4 * q + e < nl && v[e] >= thr) m4 |= 1u << e;
            }
            return m4;
        };
        auto collect = [&](unsigned m4, int q0) {         // warp-uniform: lags in ascending order
            unsigned any = __ballot_sync(0xffffffffu, m4 != 0);
            while (any && !slow) {
                const int l = __ffs(any) - 1;
                any &= any - 1;
                unsigned mm = __shfl_sync(0xffffffffu, m4, l);
                while (mm) {
                    const int e = __ffs(mm) - 1;
                    mm &= mm - 1;
                    if (nc < SCAN_MAXC) cand[nc++] = 4 * (q0 + l) + e;
                    else slow = true;
                }
            }
        };
        if (inreg) {
#pragma unroll
            for (int i = 0; i < SCAN_REG; ++i) {
                if (32 * i < nq && !slow) {
                    const int q = lane + 32 * i;
                    collect(q < nq ? hits(vr[i], q) : 0u, 32 * i);
                }
            }
        } else {
            for (int q0 = 0; q0 < nq && !slow; q0 += 32) {   // second read of the row: L1 / L2 hits
                const int q = q0 + lane;
                collect(q < nq ? hits(row4[q], q) : 0u, q0);
            }
        }
    }
    if (lane == 0) *out = slow ? make_int4(-1, -1, -1, -1) : make_int4(cand[0], cand[1], cand[2], cand[3]);
}

// ---------------------------------------------------------------------------------------------
// Tiled re-scoring.  ccx_post_kernel above streams both float64 waveforms of every pair from L2
// (48 KB per pair at n = 3000; 8.4 M pairs of configs[2] = 73 ms, L2-bandwidth bound).  Here a CTA
// keeps TC signals (events c) de-multiplexed in shared memory and walks over a group of template rows
// b: row b's waveform is staged once per CTA and shared by the 8 warps, warp w scores the pair
// (b, signal w).  L2 traffic per pair drops to n*8/TC bytes + the pair's float32 series.
// Lane l owns the samples [l*m, (l+1)*m) of every channel, m odd (conflict-free 64-bit shared loads),
// and rolls a three-sample window of the signal, so the arg-max lag and its two neighbours cost one
// shared load per tap each.
// Warps per pair: 1.  (Measured on B200, profiles/r02_ccx_rescoring.md: splitting a pair's taps over two
// warps -- 16 warps per CTA -- is SLOWER, 48 ms against 37 ms per 8.4 M pairs, and so are 12 independent
// partial sums per warp, 40 ms: the kernel is bound by the float64 FMA count, not by latency.)
constexpr int POST_PARTS = 1;

constexpr int POST_SIGS = 8;     // signals per CTA at most
constexpr int POST_WARPS = POST_SIGS * POST_PARTS;

__device__ __forceinline__ double sm_at(const double* s, int idx, int ns) {
    return (static_cast<unsigned>(idx) < static_cast<unsigned>(ns)) ? s[idx] : 0.0;
}

// acc[j] = sum_c sum_i x1_c[i] * x2_c[i + kappa - 1 + j], j = 0..2 (x2 zero outside [0, ns)).
// Only 8 warps share an SM (the tile fills its shared memory), so the inner loop must be lean: the taps
// whose three signal samples all lie inside [0, ns) run without any bounds check (two shared loads and
// three FMAs per tap, the signal window rolls through registers); at most two taps on either side of
// that range have a sample outside and take the checked path; taps with no sample inside are skipped.
__device__ __forceinline__ void sm_dot3(const double* __restrict__ s1, const double* __restrict__ s2, int ns, int Nc,
                                        int m, int kappa, int vlane, double acc[3]) {
    double a0 = 0.0, a1 = 0.0, a2 = 0.0;
    const int j0 = min(ns, vlane * m), j1 = min(ns, j0 + m);   // vlane: 0..63 over the two warps of a pair
    const int alo = max(j0, -kappa - 1), ahi = min(j1, ns - kappa + 1);      // some sample inside
    const int ilo = max(alo, 1 - kappa), ihi = min(ahi, ns - 1 - kappa);     // all three inside
    for (int c = 0; c < Nc; ++c) {
        const double* x1 = s1 + c * ns;
        const double* x2 = s2 + c * ns + kappa;
        for (int j = alo; j < min(ahi, ilo); ++j) {                           // head (<= 2 taps)
            const double v = x1[j];
            a0 = fma(v, sm_at(x2 - kappa, j + kappa - 1, ns), a0);
            a1 = fma(v, sm_at(x2 - kappa, j + kappa, ns), a1);
            a2 = fma(v, sm_at(x2 - kappa, j + kappa + 1, ns), a2);
        }
        if (ilo < ihi) {
            double w0 = x2[ilo - 1], w1 = x2[ilo];
DTX_UNROLL(DTX_CCX_UNROLL3)
            for (int j = ilo; j < ihi; ++j) {
                const double w2 = x2[j + 1];
                const double v = x1[j];
                a0 = fma(v, w0, a0);
                a1 = fma(v, w1, a1);
                a2 = fma(v, w2, a2);
                w0 = w1;
                w1 = w2;
            }
        }
        for (int j = max(alo, max(ihi, ilo)); j < ahi; ++j) {                 // tail (<= 2 taps)
            const double v = x1[j];
            a0 = fma(v, sm_at(x2 - kappa, j + kappa - 1, ns), a0);
            a1 = fma(v, sm_at(x2 - kappa, j + kappa, ns), a1);
            a2 = fma(v, sm_at(x2 - kappa, j + kappa + 1, ns), a2);
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        a0 += __shfl_xor_sync(0xffffffffu, a0, o);
        a1 += __shfl_xor_sync(0xffffffffu, a1, o);
        a2 += __shfl_xor_sync(0xffffffffu, a2, o);
    }
    acc[0] = a0; acc[1] = a1; acc[2] = a2;
}

// sum_c sum_i x1_c[i] * x2_c[i + kappa]: a single lag, no bounds checks at all (two partial sums)
__device__ __forceinline__ double sm_dot1(const double* __restrict__ s1, const double* __restrict__ s2, int ns, int Nc,
                                          int m, int kappa, int lane) {
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    const int j0 = lane * m, j1 = min(ns, j0 + m);
    const int lo = max(j0, -kappa), hi = min(j1, ns - kappa);
    for (int c = 0; c < Nc; ++c) {
        const double* x1 = s1 + c * ns;
        const double* x2 = s2 + c * ns + kappa;
        int j = lo;
        for (; j + 3 < hi; j += 4) {
            a0 = fma(x1[j], x2[j], a0);
            a1 = fma(x1[j + 1], x2[j + 1], a1);
            a2 = fma(x1[j + 2], x2[j + 2], a2);
            a3 = fma(x1[j + 3], x2[j + 3], a3);
        }
        for (; j < hi; ++j) a0 = fma(x1[j], x2[j], a0);
    }
    a0 = (a0 + a1) + (a2 + a3);
    for (int o = 16; o > 0; o >>= 1) a0 += __shfl_xor_sync(0xffffffffu, a0, o);
    return a0;
}

// Two lags in one pass over the template row (bounds-checked loads: of the variants measured --
// four lags per pass, one check-free pass per lag, check-free two-lag pass with edge loops -- this one is
// the fastest, profiles/r02_ccx_rescoring.md).
__device__ __forceinline__ void sm_dot2(const double* __restrict__ s1, const double* __restrict__ s2, int ns, int Nc,
                                        int m, int kap0, int kap1, int lane, double& r0, double& r1) {
    double a0 = 0.0, a1 = 0.0;
    const int j0 = lane * m, j1 = min(ns, j0 + m);
    for (int c = 0; c < Nc; ++c) {
        const double* x1 = s1 + c * ns;
        const double* x2 = s2 + c * ns;
DTX_UNROLL(DTX_CCX_UNROLL4)
        for (int j = j0; j < j1; ++j) {
            const double v = x1[j];
            a0 = fma(v, sm_at(x2, j + kap0, ns), a0);
            a1 = fma(v, sm_at(x2, j + kap1, ns), a1);
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        a0 += __shfl_xor_sync(0xffffffffu, a0, o);
        a1 += __shfl_xor_sync(0xffffffffu, a1, o);
    }
    r0 = a0;
    r1 = a1;
}

// Several lags within the band (rare path, kept out of the main loop's register allocation): their
// float64 values, two lags per pass; the first largest wins (np.nanargmax), NaN values are skipped.
// Returns the winning lag or -1 (every value NaN).
__device__ __noinline__ int sm_best_of(const double* s1, const double* s2, int ns, int Nc, int m, int4 kc, int koff,
                                       int lane, double sum1, double std1, double dn, const double* wac,
                                       const double* wbc) {
    double bestv = 0.0;
    int bestk = -1;
    auto take = [&](int k, double a) {
        const double v = (a - sum1 * wac[k]) / (dn * wbc[k] * std1);
        if (!isnan(v) && (bestk < 0 || v > bestv)) { bestv = v; bestk = k; }
    };
    double a0, a1;
    sm_dot2(s1, s2, ns, Nc, m, kc.x + koff, kc.y + koff, lane, a0, a1);
    take(kc.x, a0);
    take(kc.y, a1);
    if (kc.z >= 0) {
        const int k3 = kc.w >= 0 ? kc.w : kc.z;
        sm_dot2(s1, s2, ns, Nc, m, kc.z + koff, k3 + koff, lane, a0, a1);
        take(kc.z, a0);
        if (kc.w >= 0) take(kc.w, a1);
    }
    return bestk;
}

__device__ __forceinline__ double sm_exact(const double* __restrict__ s1, const double* __restrict__ s2, int ns, int Nc,
                                           int m, int kappa, int lane) {
    double a = 0.0;
    const int j0 = lane * m, j1 = min(ns, j0 + m);
    for (int c = 0; c < Nc; ++c)
        for (int j = j0; j < j1; ++j) a = fma(s1[c * ns + j], sm_at(s2 + c * ns, j + kappa, ns), a);
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    return a;
}

template <typename T>
__device__ __forceinline__ void stage_demux(const T* __restrict__ x, double* __restrict__ dst, int n, int Nc, int ns,
                                            int tid, int nthreads) {
    for (int i = tid; i < n; i += nthreads) dst[(i % Nc) * ns + i / Nc] = static_cast<double>(x[i]);
}

template <typename T>
__global__ void __launch_bounds__(POST_WARPS * 32)
ccx_post_tiled_kernel(const int4* __restrict__ karg, int c0, int nsig,
                      const T* __restrict__ X, int N, int n, int Nc, int trunc, int nl,
                      const int* __restrict__ rows, int nrows, int TC, int RG,
                      const double* __restrict__ wa, const double* __restrict__ wb,
                      const double* __restrict__ evsum, const double* __restrict__ evstd,
                      double* __restrict__ cc, int* __restrict__ lag, double* __restrict__ sub,
                      int* __restrict__ nflag, int2* __restrict__ flagged, int flag_cap) {
    extern __shared__ double sm[];   // [TC] signals, then the current template row; each [Nc][ns]
    __shared__ double pbuf[POST_SIGS][3];                // partial sums of a pair's second warp
    const int ns = n / Nc;
    const int m32 = ((ns + 31) / 32) | 1;                // taps per lane and channel when ONE warp covers a row (odd:
    const int m64 = ((ns + 32 * POST_PARTS - 1) / (32 * POST_PARTS)) | 1;   // conflict-free 64-bit shared loads); when POST_PARTS warps do
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int sig = warp / POST_PARTS, part = warp % POST_PARTS;   // POST_PARTS warps share a signal: taps split among them
    const int vlane = part * 32 + lane;
    const int sig0 = blockIdx.y * TC;
    const int nmine = min(TC, nsig - sig0);              // signals of this CTA
    const int r0 = blockIdx.x * RG, r1 = min(nrows, r0 + RG);
    const int c_last = c0 + sig0 + nmine - 1;
    if (r0 >= nrows || rows[r0] >= c_last) return;       // nothing with b < c in this tile
    for (int s = 0; s < nmine; ++s)
        stage_demux(X + static_cast<long long>(c0 + sig0 + s) * n, sm + static_cast<size_t>(s) * n, n, Nc, ns, tid,
                    POST_WARPS * 32);
    double* s1 = sm + static_cast<size_t>(TC) * n;
    const bool have_sig = sig < nmine;
    const int ci = sig0 + (have_sig ? sig : 0);          // this warp's signal (index inside the batch)
    const int c = c0 + ci;
    const double* s2 = sm + static_cast<size_t>(have_sig ? sig : 0) * n;
    const double* wac = wa + static_cast<long long>(c) * nl;
    const double* wbc = wb + static_cast<long long>(c) * nl;
    const double dn = static_cast<double>(n);
    // the next template row is fetched into registers while the current one is scored (n <= 4096;
    // longer waveforms are staged directly); the de-multiplexed shared-memory offset of each of a
    // thread's elements is computed once
    constexpr int PF = 16 / POST_PARTS;
    const bool use_pf = n <= PF * POST_WARPS * 32;
    double pf[PF];
    int soff[PF];
#pragma unroll
    for (int k = 0; k < PF; ++k) {
        const int i = tid + k * POST_WARPS * 32;
        soff[k] = i < n ? (i % Nc) * ns + i / Nc : -1;
    }
    int4 k_nx = make_int4(-2, -1, -1, -1);   // the next row's candidate lags and per-event scalars (fetched a row ahead)
    double std_nx = 0.0, sum_nx = 0.0;
    auto prefetch = [&](int r) {
        const int bb = rows[r];
        if (use_pf) {
            const T* x = X + static_cast<long long>(bb) * n;
#pragma unroll
            for (int k = 0; k < PF; ++k) pf[k] = soff[k] >= 0 ? static_cast<double>(x[tid + k * POST_WARPS * 32]) : 0.0;
        }
        k_nx = (have_sig && bb < c) ? karg[static_cast<long long>(ci) * nrows + r] : make_int4(-2, -1, -1, -1);
        std_nx = evstd[bb];
        sum_nx = evsum[bb];
    };
    // Finalisation is deferred twice.  (i) The first warp of a pair keeps its three partial dot products
    // until the second warp's have arrived through shared memory (next barrier).  (ii) Lane (count % 32)
    // of the first warp then keeps the complete sums of a pair, and every 32 pairs all lanes normalise, fit
    // the cosine and store in parallel -- the float64 divisions and acos / sin / atan would otherwise be
    // executed by the whole warp for every single pair.
    double q_a0 = 0.0, q_a1 = 0.0, q_a2 = 0.0, q_sum = 0.0, q_std = 0.0;
    int q_k = 0, q_r = -1;
    double p_a0 = 0.0, p_a1 = 0.0, p_a2 = 0.0, p_sum = 0.0, p_std = 0.0;
    long long p_o = -1;
    int p_k = 0, p_b = 0, npend = 0;
    auto flush = [&]() {
        if (p_o >= 0) {
            const int k = p_k;
            const double v = (p_a1 - p_sum * wac[k]) / (dn * wbc[k] * p_std);
            if (isnan(v) || v > 1.0 + CC_GUARD) {
                const int q = atomicAdd(nflag, 1);
                if (q < flag_cap) flagged[q] = make_int2(p_b, c);   // p_b holds the slot r
            } else {
                double ss = 0.0;                          // maximum at either end of the lag range: 0 (construct.py:402)
                if (k > 0 && k < nl - 1) {
                    const double cb4 = (p_a0 - p_sum * wac[k - 1]) / (dn * wbc[k - 1] * p_std);
                    const double caf = (p_a2 - p_sum * wac[k + 1]) / (dn * wbc[k + 1] * p_std);
                    const double alpha = acos((cb4 + caf) / (2 * v));
                    const double alsi = sin(alpha);
                    const double tau = -(atan((cb4 - caf) / (2 * v * alsi)) / alpha);
                    ss = (fabs(tau) > 0.5) ? static_cast<double>(k) : tau;   // reference quirk, construct.py:418-421
                }
                cc[p_o] = v;
                lag[p_o] = (k + 1 + trunc) * Nc - n;
                sub[p_o] = ss;
            }
            p_o = -1;
        }
        npend = 0;
    };
    auto complete = [&]() {      // first warp of a pair, after a barrier: add the second warp's partial sums
        if (q_r >= 0) {
            if (lane == npend) {
                if (POST_PARTS > 1) { q_a0 += pbuf[sig][0]; q_a1 += pbuf[sig][1]; q_a2 += pbuf[sig][2]; }
                p_a0 = q_a0; p_a1 = q_a1; p_a2 = q_a2;
                p_sum = q_sum; p_std = q_std; p_k = q_k; p_b = q_r;
                p_o = static_cast<long long>(q_r) * N + c;
            }
            q_r = -1;
            if (++npend == 32) flush();
        }
    };
    prefetch(r0);
    for (int r = r0; r < r1; ++r) {
        const int b = rows[r];
        if (b >= c_last) break;                          // rows ascend: no later row has a pair here
        __syncthreads();                                 // previous row's readers are done with s1, pbuf is written
        if (part == 0) complete();
        if (use_pf) {
#pragma unroll
            for (int k = 0; k < PF; ++k)
                if (soff[k] >= 0) s1[soff[k]] = pf[k];
        } else {
            stage_demux(X + static_cast<long long>(b) * n, s1, n, Nc, ns, tid, POST_WARPS * 32);
        }
        __syncthreads();
        const int4 kc = k_nx;                             // candidate lags of the float32 series
        const double std1 = std_nx, sum1 = sum_nx;
        if (r + 1 < r1) prefetch(r + 1);
        if (kc.x < 0) continue;                           // b >= c, or left to ccx_post_kernel
        int k = kc.x;
        if (kc.y >= 0) {
            // both warps of the pair evaluate the candidates over all taps and reach the same winner
            k = sm_best_of(s1, s2, ns, Nc, m32, kc, trunc + 1 - ns, lane, sum1, std1, dn, wac, wbc);
            if (k < 0) {                                  // every candidate NaN: let the float64 kernel decide
                if (lane == 0 && part == 0) {
                    const int q = atomicAdd(nflag, 1);
                    if (q < flag_cap) flagged[q] = make_int2(r, c);
                }
                continue;
            }
        }
        double acc[3];
        sm_dot3(s1, s2, ns, Nc, m64, k + trunc + 1 - ns, vlane, acc);
        if (part == 1) {
            if (lane < 3) pbuf[sig][lane] = acc[lane == 0 ? 0 : (lane == 1 ? 1 : 2)];
        } else {
            q_a0 = acc[0]; q_a1 = acc[1]; q_a2 = acc[2];
            q_sum = sum1; q_std = std1; q_k = k; q_r = r;
            if (POST_PARTS == 1) complete();              // nothing to wait for
        }
    }
    __syncthreads();
    if (part == 0) {
        complete();
        flush();
    }
}


// ---------------------------------------------------------------------------------------------
// Ring re-scoring (round 2, default).  The tiled kernel above synchronises the whole CTA twice per
// template row, and every row takes as long as its slowest pair: with the one-MMA screening series about
// half of the pairs carry 2-4 candidate lags, so nearly every row of 8 pairs waits for a slow one (ncu:
// 2.7 barrier-stall cycles per issued instruction, shared-memory pipe 41 % busy).  Here
//   * the events are de-multiplexed once per call into a float64 copy, so a waveform is ONE contiguous
//     cp.async.bulk of n*8 bytes (no cooperative staging, no CTA barrier);
//   * a producer warp keeps a ring of NB template rows full (mbarrier full / empty per slot; the row's
//     candidate records and scalars are parked next to it), TC signals stay resident;
//   * the consumer warps (11 by default) take (row, signal) tasks from a shared counter in row order -- a warp that drew a
//     one-candidate pair simply takes the next task, up to NB - 1 rows ahead of the slowest warp;
//   * a pair with several candidates is scored in ONE pass: the template sample is loaded once and every
//     candidate rolls its own three-sample window of the signal (1 + C shared loads and 3C FMAs per tap;
//     the two-pass scheme cost 3 + 2 loads for C = 2 and 6 + 2 for C = 4), which also yields the winner's
//     neighbours for the cosine fit.
// Sums are accumulated in the same order as sm_dot3, so results are bit-identical to the tiled kernel's.
// DTX_RING_TOOL_SYNC = 1 expresses the same producer / consumer protocol per THREAD -- every consumer lane arrives on
// the slot's `empty` barrier itself, lane 0 of the producer parks the records it releases, the load gate is read with
// atomics -- which is what compute-sanitizer's racecheck can follow: 0 hazards (profiles/r02_sanitizer.md).  The
// default lets lane 0 arrive for its warp after a __syncwarp (32 x fewer shared-memory atomics per task: 83 ms
// instead of 97 ms per 4096-event call); racecheck, which tracks barriers per thread, reports that form as hazards.
#ifndef DTX_RING_TOOL_SYNC
#define DTX_RING_TOOL_SYNC 0
#endif
constexpr int RING_THREADS = 512;                 // at most; warp 0 = producer, the others consume (default 12 warps)
constexpr int RING_TC_MAX = 16;                   // resident signals at most
constexpr int RING_NB_MAX = 4;                    // ring slots at most

struct RingRec {
    int4 kc[RING_TC_MAX];
    double std1, sum1;
};

// Three lags (k-1, k, k+1) of each of C candidate lags in one pass over the template row; bounds-checked
// signal loads (zero outside [0, ns)), per-lane partition and summation order of sm_dot3.
template <int C>
__device__ __forceinline__ void sm_multi3(const double* __restrict__ s1, const double* __restrict__ s2, int ns, int Nc,
                                          int m, const int (&kap)[C], int lane, double (&a)[C][3]) {
#pragma unroll
    for (int q = 0; q < C; ++q) a[q][0] = a[q][1] = a[q][2] = 0.0;
    const int j0 = min(ns, lane * m), j1 = min(ns, j0 + m);
    for (int c = 0; c < Nc; ++c) {
        const double* x1 = s1 + c * ns;
        const double* x2 = s2 + c * ns;
        double w0[C], w1[C];
#pragma unroll
        for (int q = 0; q < C; ++q) {
            w0[q] = sm_at(x2, j0 + kap[q] - 1, ns);
            w1[q] = sm_at(x2, j0 + kap[q], ns);
        }
DTX_UNROLL(2)
        for (int j = j0; j < j1; ++j) {
            const double v = x1[j];
#pragma unroll
            for (int q = 0; q < C; ++q) {
                const double w2 = sm_at(x2, j + kap[q] + 1, ns);
                a[q][0] = fma(v, w0[q], a[q][0]);
                a[q][1] = fma(v, w1[q], a[q][1]);
                a[q][2] = fma(v, w2, a[q][2]);
                w0[q] = w1[q];
                w1[q] = w2;
            }
        }
    }
#pragma unroll
    for (int q = 0; q < C; ++q)
#pragma unroll
        for (int e = 0; e < 3; ++e)
            for (int o = 16; o > 0; o >>= 1) a[q][e] += __shfl_xor_sync(0xffffffffu, a[q][e], o);
}

// Winner among C candidates (first largest float64 value, NaN skipped: np.nanargmax) and its three sums.
// Returns the winning lag or -1.
template <int C>
__device__ __noinline__ int sm_best3(const double* s1, const double* s2, int ns, int Nc, int m, int4 kc, int koff,
                                     int lane, double sum1, double std1, double dn, const double* wac,
                                     const double* wbc, double acc[3]) {
    const int ks[4] = {kc.x, kc.y, kc.z, kc.w};
    int kap[C];
#pragma unroll
    for (int q = 0; q < C; ++q) kap[q] = ks[q] + koff;
    double a[C][3];
    sm_multi3<C>(s1, s2, ns, Nc, m, kap, lane, a);
    double bestv = 0.0;
    int bestk = -1;
#pragma unroll
    for (int q = 0; q < C; ++q) {
        const int k = ks[q];
        const double v = (a[q][1] - sum1 * wac[k]) / (dn * wbc[k] * std1);
        if (!isnan(v) && (bestk < 0 || v > bestv)) {
            bestv = v; bestk = k;
            acc[0] = a[q][0]; acc[1] = a[q][1]; acc[2] = a[q][2];
        }
    }
    return bestk;
}

__global__ void __launch_bounds__(RING_THREADS, 1)
ccx_post_ring_kernel(const int4* __restrict__ karg, int c0, int nsig, const double* __restrict__ Xd, int N, int n,
                     int Nc, int trunc, int nl, const int* __restrict__ rows, int nrows, int TC, int NB, int RG,
                     const double* __restrict__ wa, const double* __restrict__ wb,
                     const double* __restrict__ evsum, const double* __restrict__ evstd,
                     double* __restrict__ cc, int* __restrict__ lag, double* __restrict__ sub,
                     int* __restrict__ nflag, int2* __restrict__ flagged, int flag_cap) {
    extern __shared__ __align__(16) double sm[];   // [TC] signals, [NB] template rows; each [Nc][ns]
    __shared__ __align__(8) uint64_t bar_full[RING_NB_MAX], bar_empty[RING_NB_MAX], bar_sig;
    __shared__ RingRec rec[RING_NB_MAX];
    __shared__ int next_task;
    __shared__ int issued;                               // rows whose load has been issued (see the consumer's wait)
    const int ns = n / Nc;
    const int m32 = ((ns + 31) / 32) | 1;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int sig0 = blockIdx.y * TC;
    const int nmine = min(TC, nsig - sig0);              // signals of this CTA
    const int r0 = blockIdx.x * RG;
    const int c_last = c0 + sig0 + nmine - 1;
    if (r0 >= nrows || rows[r0] >= c_last) return;       // nothing with b < c in this tile (CTA-uniform)
    // rows ascend: the rows of this tile that have a pair here are a prefix
    int lo = r0, hi = min(nrows, r0 + RG);
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (rows[mid] < c_last) lo = mid + 1; else hi = mid;
    }
    const int nr = lo - r0;                               // >= 1
    const uint32_t row_bytes = static_cast<uint32_t>(n) * 8u;
    if (tid == 0) {
        for (int i = 0; i < NB; ++i) {
            mbar_init(&bar_full[i], 1);
            mbar_init(&bar_empty[i], DTX_RING_TOOL_SYNC ? 32 * nmine : nmine);
        }
        mbar_init(&bar_sig, 1);
        next_task = 0;
        issued = 0;
        fence_barrier_init();
    }
    __syncthreads();
    double* srow = sm + static_cast<size_t>(TC) * n;
    if (warp == 0) {
        // ------------------------------------------------------------------ producer
        if (lane == 0) {
            mbar_arrive_expect_tx(&bar_sig, row_bytes * nmine);
            for (int s = 0; s < nmine; ++s)
                bulk_g2s(sm + static_cast<size_t>(s) * n, Xd + static_cast<long long>(c0 + sig0 + s) * n, row_bytes,
                         &bar_sig);
        }
        for (int i = 0; i < nr; ++i) {
            const int slot = i % NB;
            if (i >= NB) mbar_wait(&bar_empty[slot], ((i / NB) - 1) & 1);
            const int r = r0 + i;
            const int b = rows[r];
            // the lanes fetch the row's candidate records together
            int4 kcl = make_int4(-2, -1, -1, -1);
            if (lane < nmine && b < c0 + sig0 + lane) kcl = karg[static_cast<long long>(sig0 + lane) * nrows + r];
#if DTX_RING_TOOL_SYNC
            // lane 0 alone parks them and releases them with its own arrive
#pragma unroll
            for (int s = 0; s < RING_TC_MAX; ++s) {
                int4 v;
                v.x = __shfl_sync(0xffffffffu, kcl.x, s); v.y = __shfl_sync(0xffffffffu, kcl.y, s);
                v.z = __shfl_sync(0xffffffffu, kcl.z, s); v.w = __shfl_sync(0xffffffffu, kcl.w, s);
                if (lane == 0 && s < nmine) rec[slot].kc[s] = v;
            }
#else
            if (lane < nmine) rec[slot].kc[lane] = kcl;   // ordered before lane 0's arrive by the __syncwarp
#endif
            if (lane == 0) { rec[slot].std1 = evstd[b]; rec[slot].sum1 = evsum[b]; }
            __syncwarp();
            if (lane == 0) {
                mbar_arrive_expect_tx(&bar_full[slot], row_bytes);
                bulk_g2s(srow + static_cast<size_t>(slot) * n, Xd + static_cast<long long>(b) * n, row_bytes,
                         &bar_full[slot]);
                __threadfence_block();
                atomicExch(&issued, i + 1);
            }
        }
        return;
    }
    // ---------------------------------------------------------------------- consumers
    const double dn = static_cast<double>(n);
    const int koff = trunc + 1 - ns;
    // deferred finalisation: lane (count % 32) keeps a finished pair's sums; every 32 pairs all lanes normalise,
    // fit the cosine and store in parallel
    double p_a0 = 0.0, p_a1 = 0.0, p_a2 = 0.0, p_sum = 0.0, p_std = 0.0;
    long long p_o = -1;
    int p_k = 0, p_r = 0, p_c = 0, npend = 0;
    auto flush = [&]() {
        if (p_o >= 0) {
            const int k = p_k;
            const double* wac = wa + static_cast<long long>(p_c) * nl;
            const double* wbc = wb + static_cast<long long>(p_c) * nl;
            const double v = (p_a1 - p_sum * wac[k]) / (dn * wbc[k] * p_std);
            if (isnan(v) || v > 1.0 + CC_GUARD) {
                const int q = atomicAdd(nflag, 1);
                if (q < flag_cap) flagged[q] = make_int2(p_r, p_c);
            } else {
                double ss = 0.0;                          // maximum at either end of the lag range: 0 (construct.py:402)
                if (k > 0 && k < nl - 1) {
                    const double cb4 = (p_a0 - p_sum * wac[k - 1]) / (dn * wbc[k - 1] * p_std);
                    const double caf = (p_a2 - p_sum * wac[k + 1]) / (dn * wbc[k + 1] * p_std);
                    const double alpha = acos((cb4 + caf) / (2 * v));
                    const double alsi = sin(alpha);
                    const double tau = -(atan((cb4 - caf) / (2 * v * alsi)) / alpha);
                    ss = (fabs(tau) > 0.5) ? static_cast<double>(k) : tau;   // reference quirk, construct.py:418-421
                }
                cc[p_o] = v;
                lag[p_o] = (k + 1 + trunc) * Nc - n;
                sub[p_o] = ss;
            }
            p_o = -1;
        }
        npend = 0;
    };
    mbar_wait(&bar_sig, 0);
    const int ntask = nr * nmine;
    for (;;) {
        int id = 0;
        if (lane == 0) id = atomicAdd(&next_task, 1);
        id = __shfl_sync(0xffffffffu, id, 0);
        if (id >= ntask) break;
        const int i = id / nmine, s = id - i * nmine;
        const int slot = i % NB;
        // A parity wait is only meaningful once the barrier has reached this row's phase: with few tasks per row a
        // warp can draw a task two uses of the slot ahead, where the parity test would alias to a finished phase.
#if DTX_RING_TOOL_SYNC
        if (lane == 0)
            while (atomicAdd(&issued, 0) <= i) __nanosleep(32);
        __syncwarp();
#else
        while (*static_cast<volatile int*>(&issued) <= i) __nanosleep(32);
#endif
        mbar_wait(&bar_full[slot], (i / NB) & 1);
        const int4 kc = rec[slot].kc[s];
        if (kc.x >= 0) {                                  // else: b >= c, or left to ccx_post_kernel
            const double std1 = rec[slot].std1, sum1 = rec[slot].sum1;
            const int r = r0 + i, c = c0 + sig0 + s;
            const double* s1 = srow + static_cast<size_t>(slot) * n;
            const double* s2 = sm + static_cast<size_t>(s) * n;
            double acc[3];
            int k = kc.x;
            if (kc.y >= 0) {
                const double* wac = wa + static_cast<long long>(c) * nl;
                const double* wbc = wb + static_cast<long long>(c) * nl;
                if (kc.z < 0) k = sm_best3<2>(s1, s2, ns, Nc, m32, kc, koff, lane, sum1, std1, dn, wac, wbc, acc);
                else if (kc.w < 0) k = sm_best3<3>(s1, s2, ns, Nc, m32, kc, koff, lane, sum1, std1, dn, wac, wbc, acc);
                else k = sm_best3<4>(s1, s2, ns, Nc, m32, kc, koff, lane, sum1, std1, dn, wac, wbc, acc);
            } else {
                sm_dot3(s1, s2, ns, Nc, m32, k + koff, lane, acc);
            }
            if (k < 0) {                                  // every candidate NaN: let the float64 kernel decide
                if (lane == 0) {
                    const int q = atomicAdd(nflag, 1);
                    if (q < flag_cap) flagged[q] = make_int2(r, c);
                }
            } else {
                if (lane == npend) {
                    p_a0 = acc[0]; p_a1 = acc[1]; p_a2 = acc[2];
                    p_sum = sum1; p_std = std1; p_k = k; p_r = r; p_c = c;
                    p_o = static_cast<long long>(r) * N + c;
                }
                if (++npend == 32) flush();
            }
        }
#if DTX_RING_TOOL_SYNC
        mbar_arrive(&bar_empty[slot]);                   // this lane no longer reads the slot
#else
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_empty[slot]);   // this warp no longer reads the slot
#endif
    }
    flush();
}

// events de-multiplexed into float64 [N][Nc][ns] (one bulk copy per waveform in the ring kernel)
template <typename T>
__global__ void __launch_bounds__(256)
ccx_demux_kernel(const T* __restrict__ X, int n, int Nc, double* __restrict__ Xd) {
    const int ns = n / Nc;
    const T* x = X + static_cast<long long>(blockIdx.x) * n;
    double* d = Xd + static_cast<long long>(blockIdx.x) * n;
    for (int i = threadIdx.x; i < n; i += 256) d[(i % Nc) * ns + i / Nc] = static_cast<double>(x[i]);
}

}  // namespace

void launch_ccx_stats(const void* d_X, int dtype_f32, int N, int n, int Nc, double* wa, double* wb, double* es,
                      double* ed, cudaStream_t st) {
    const int ns = n / Nc;
    const int trunc = n / (2 * Nc) - 1;
    const int nl = 2 * ns - 1 - 2 * trunc;
    const size_t sm_stats = sizeof(double) * 2 * Nc * (ns + 1);
    if (dtype_f32) {
        cudaFuncSetAttribute(ccx_stats<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sm_stats));
        ccx_stats<float><<<N, CT, sm_stats, st>>>(static_cast<const float*>(d_X), N, n, Nc, trunc, nl, wa, wb, es, ed);
    } else {
        cudaFuncSetAttribute(ccx_stats<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sm_stats));
        ccx_stats<double><<<N, CT, sm_stats, st>>>(static_cast<const double*>(d_X), N, n, Nc, trunc, nl, wa, wb, es, ed);
    }
}

void launch_ccx_fp64(const void* d_X, int dtype_f32, int N, int n, int Nc, int row_begin, int nrows, const int* d_rows,
                     const double* wa, const double* wb, const double* es, const double* ed, double* d_cc,
                     int* d_lag, double* d_sub, int num_sms, cudaStream_t st) {
    const int ns = n / Nc;
    const int trunc = n / (2 * Nc) - 1;
    const int nl = 2 * ns - 1 - 2 * trunc;
    const size_t sm_res = sizeof(double) * nl;
    const int rows = nrows;
    int gx = (2 * num_sms + rows - 1) / rows;
    if (gx < 1) gx = 1;
    if (gx > N) gx = N;
    const dim3 grid(gx, rows);
    if (dtype_f32) {
        cudaFuncSetAttribute(ccx_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sm_res));
        ccx_kernel<float><<<grid, CT, sm_res, st>>>(static_cast<const float*>(d_X), N, n, Nc, trunc, nl, row_begin,
                                                    d_rows, wa, wb, es, ed, d_cc, d_lag, d_sub, nullptr, nullptr, 0);
    } else {
        cudaFuncSetAttribute(ccx_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sm_res));
        ccx_kernel<double><<<grid, CT, sm_res, st>>>(static_cast<const double*>(d_X), N, n, Nc, trunc, nl, row_begin,
                                                     d_rows, wa, wb, es, ed, d_cc, d_lag, d_sub, nullptr, nullptr, 0);
    }
}

// The float64 kernel over a device-side list of (slot, event c) pairs (no host round trip).
void launch_ccx_fp64_pairs(const void* d_X, int dtype_f32, int N, int n, int Nc, const int* d_rows, const double* wa,
                           const double* wb, const double* es, const double* ed, double* d_cc, int* d_lag,
                           double* d_sub, const int2* d_pairs, const int* d_npairs, int pair_cap, int num_sms,
                           cudaStream_t st) {
    const int ns = n / Nc;
    const int trunc = n / (2 * Nc) - 1;
    const int nl = 2 * ns - 1 - 2 * trunc;
    const size_t sm_res = sizeof(double) * nl;
    const dim3 grid(4 * num_sms, 1);
    if (dtype_f32) {
        cudaFuncSetAttribute(ccx_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sm_res));
        ccx_kernel<float><<<grid, CT, sm_res, st>>>(static_cast<const float*>(d_X), N, n, Nc, trunc, nl, 0, d_rows, wa, wb,
                                                    es, ed, d_cc, d_lag, d_sub, d_pairs, d_npairs, pair_cap);
    } else {
        cudaFuncSetAttribute(ccx_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sm_res));
        ccx_kernel<double><<<grid, CT, sm_res, st>>>(static_cast<const double*>(d_X), N, n, Nc, trunc, nl, 0, d_rows, wa,
                                                     wb, es, ed, d_cc, d_lag, d_sub, d_pairs, d_npairs, pair_cap);
    }
}

// Events [row_begin, row_begin + rows) as rank-1 templates of the tensor-core engine:
// x / ||x - mean|| in float64 (zero rows for zeroed-out waveforms).  One block per event.
template <typename T>
__global__ void __launch_bounds__(256)
ccx_templates_kernel(const T* __restrict__ X, int n, const int* __restrict__ rows, double* __restrict__ U) {
    const T* x = X + static_cast<long long>(rows[blockIdx.x]) * n;
    double* u = U + static_cast<long long>(blockIdx.x) * n;
    __shared__ double sh[8];
    __shared__ double bc;
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    auto block_sum = [&](double v) {
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        __syncthreads();
        if (l == 0) sh[w] = v;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0;
            for (int i = 0; i < 8; ++i) t += sh[i];
            bc = t;
        }
        __syncthreads();
        return bc;
    };
    double s1 = 0;
    for (int i = threadIdx.x; i < n; i += 256) s1 += static_cast<double>(x[i]);
    const double mean = block_sum(s1) / n;
    double s2 = 0;
    for (int i = threadIdx.x; i < n; i += 256) {
        const double v = static_cast<double>(x[i]) - mean;
        s2 += v * v;
    }
    const double nrm = sqrt(block_sum(s2));
    for (int i = threadIdx.x; i < n; i += 256) u[i] = nrm > 0 ? static_cast<double>(x[i]) / nrm : 0.0;
}

void launch_ccx_templates(const void* d_X, int dtype_f32, int n, const int* d_rows, int rows, double* d_U,
                          cudaStream_t st) {
    if (dtype_f32)
        ccx_templates_kernel<float><<<rows, 256, 0, st>>>(static_cast<const float*>(d_X), n, d_rows, d_U);
    else
        ccx_templates_kernel<double><<<rows, 256, 0, st>>>(static_cast<const double*>(d_X), n, d_rows, d_U);
}

void launch_ccx_pad(const void* d_X, int dtype_f32, int n, int Nc, int c0, int nsig, int P, int Lc, double* out,
                    cudaStream_t st) {
    const dim3 grid(32, nsig);
    if (dtype_f32)
        ccx_pad_kernel<float><<<grid, 256, 0, st>>>(static_cast<const float*>(d_X), n, Nc, c0, P, Lc, out);
    else
        ccx_pad_kernel<double><<<grid, 256, 0, st>>>(static_cast<const double*>(d_X), n, Nc, c0, P, Lc, out);
}

void launch_ccx_demux(const void* d_X, int dtype_f32, int N, int n, int Nc, double* d_Xd, cudaStream_t st) {
    if (dtype_f32) ccx_demux_kernel<float><<<N, 256, 0, st>>>(static_cast<const float*>(d_X), n, Nc, d_Xd);
    else ccx_demux_kernel<double><<<N, 256, 0, st>>>(static_cast<const double*>(d_X), n, Nc, d_Xd);
}

static int env_int(const char* name, int dflt) {
    const char* v = std::getenv(name);
    return (v && *v) ? std::atoi(v) : dflt;
}

void launch_ccx_post(const float* DS, const ChunkDesc* d_chunks, int c0, int nsig, const void* d_X, int dtype_f32,
                     const double* d_Xd, int N, int n, int Nc, const int* d_rows, int nrows, const double* wa,
                     const double* wb, const double* es, const double* ed, double* d_cc, int* d_lag, double* d_sub,
                     int* d_nflag, int2* d_flagged, int flag_cap, int4* d_karg, const unsigned* d_ratio_bits,
                     float band0, cudaStream_t st) {
    const int ns = n / Nc;
    const int trunc = n / (2 * Nc) - 1;
    const int nl = 2 * ns - 1 - 2 * trunc;
    const int rows = nrows;
    const long long per_wave = static_cast<long long>(n) * 8;
    const int4* karg = nullptr;
    // ring variant (default): TC resident signals + a ring of NB template rows, float64, in shared memory
    const int slots = static_cast<int>(std::min<long long>(RING_TC_MAX + RING_NB_MAX, (227 * 1024 - 2048 - 128) / per_wave));
    const bool ring = d_Xd && d_karg && n % 2 == 0 && slots >= 3 && !env_int("DTX_CCX_POST_TILED", 0) &&
                      !std::getenv("DTX_CCX_POST_UNTILED");
    // tiled variant: TC signals + 1 template row, float64, in shared memory
    const int TC = static_cast<int>(std::min<long long>(POST_SIGS, (220 * 1024 - 128) / per_wave - 1));
    if (ring) {
        const dim3 sg((rows + 7) / 8, nsig);
        ccx_scan_kernel<<<sg, 256, 0, st>>>(DS, d_chunks, c0, N, nl, d_rows, nrows, es, ed, n, d_ratio_bits, band0, d_cc,
                                            d_lag, d_sub, d_karg);
        int NB = env_int("DTX_CCX_RING_NB", slots >= 8 ? 3 : 2);
        NB = std::max(2, std::min(NB, std::min(RING_NB_MAX, slots - 1)));
        int RTC = env_int("DTX_CCX_RING_TC", slots - NB);
        RTC = std::max(1, std::min(RTC, std::min(RING_TC_MAX, slots - NB)));
        const int RG = std::max(1, env_int("DTX_CCX_RING_RG", 64));
        const size_t smem = static_cast<size_t>(RTC + NB) * per_wave + 128;
        const dim3 tg((rows + RG - 1) / RG, (nsig + RTC - 1) / RTC);
        cudaFuncSetAttribute(ccx_post_ring_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        const int nthreads = 32 * std::max(2, std::min(RING_THREADS / 32, env_int("DTX_CCX_RING_WARPS", 12)));
        ccx_post_ring_kernel<<<tg, nthreads, smem, st>>>(d_karg, c0, nsig, d_Xd, N, n, Nc, trunc, nl, d_rows, nrows,
                                                             RTC, NB, RG, wa, wb, es, ed, d_cc, d_lag, d_sub, d_nflag,
                                                             d_flagged, flag_cap);
        karg = d_karg;   // ccx_post_kernel below: only the pairs the scan left for it
    } else if (TC >= 1 && d_karg && !std::getenv("DTX_CCX_POST_UNTILED")) {
        const dim3 sg((rows + 7) / 8, nsig);
        ccx_scan_kernel<<<sg, 256, 0, st>>>(DS, d_chunks, c0, N, nl, d_rows, nrows, es, ed, n, d_ratio_bits, band0, d_cc,
                                            d_lag, d_sub, d_karg);
        const int RG = 64;
        const size_t smem = static_cast<size_t>(TC + 1) * per_wave + 128;   // + slack behind the last row
        const dim3 tg((rows + RG - 1) / RG, (nsig + TC - 1) / TC);
        if (dtype_f32) {
            cudaFuncSetAttribute(ccx_post_tiled_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 static_cast<int>(smem));
            ccx_post_tiled_kernel<float><<<tg, POST_WARPS * 32, smem, st>>>(
                d_karg, c0, nsig, static_cast<const float*>(d_X), N, n, Nc, trunc, nl, d_rows, nrows, TC, RG, wa,
                wb, es, ed, d_cc, d_lag, d_sub, d_nflag, d_flagged, flag_cap);
        } else {
            cudaFuncSetAttribute(ccx_post_tiled_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 static_cast<int>(smem));
            ccx_post_tiled_kernel<double><<<tg, POST_WARPS * 32, smem, st>>>(
                d_karg, c0, nsig, static_cast<const double*>(d_X), N, n, Nc, trunc, nl, d_rows, nrows, TC, RG, wa,
                wb, es, ed, d_cc, d_lag, d_sub, d_nflag, d_flagged, flag_cap);
        }
        karg = d_karg;   // ccx_post_kernel below: only the pairs the scan left for it
    }
    const dim3 grid((rows + 7) / 8, nsig);
    if (dtype_f32)
        ccx_post_kernel<float><<<grid, 256, 0, st>>>(DS, d_chunks, c0, static_cast<const float*>(d_X), N, n, Nc, trunc,
                                                     nl, d_rows, nrows, wa, wb, es, ed, d_cc, d_lag, d_sub,
                                                     d_nflag, d_flagged, flag_cap, karg, d_ratio_bits, band0);
    else
        ccx_post_kernel<double><<<grid, 256, 0, st>>>(DS, d_chunks, c0, static_cast<const double*>(d_X), N, n, Nc,
                                                      trunc, nl, d_rows, nrows, wa, wb, es, ed, d_cc, d_lag,
                                                      d_sub, d_nflag, d_flagged, flag_cap, karg, d_ratio_bits, band0);
}

// dense [nslots][N] rows -> SciPy condensed order; one block row per event b, threads over c > b
__global__ void __launch_bounds__(256)
ccx_pack_kernel(const double* __restrict__ cc, const int* __restrict__ lag, const double* __restrict__ sub,
                const int* __restrict__ slot_of_row, int N, double* __restrict__ o_cc, int* __restrict__ o_lag,
                double* __restrict__ o_sub) {
    const int b = blockIdx.y;
    const int c = b + 1 + blockIdx.x * 256 + threadIdx.x;
    if (c >= N) return;
    const long long src = static_cast<long long>(slot_of_row[b]) * N + c;
    const long long dst = static_cast<long long>(b) * N - static_cast<long long>(b) * (b + 1) / 2 + (c - b - 1);
    o_cc[dst] = cc[src];
    o_lag[dst] = lag[src];
    o_sub[dst] = sub[src];
}

// A rank's own template rows -> their places in the condensed arrays of the WHOLE matrix.  The output pointers may
// be page-locked host memory mapped into the device (zero-copy stores over this GPU's own PCIe link; every row's
// entries are contiguous, so a warp writes 256 / 128 contiguous bytes).  One block row per slot, threads over c > b.
__global__ void __launch_bounds__(256)
ccx_pack_rows_kernel(const double* __restrict__ cc, const int* __restrict__ lag, const double* __restrict__ sub,
                     const int* __restrict__ rows, int N, double* __restrict__ o_cc, int* __restrict__ o_lag,
                     double* __restrict__ o_sub) {
    const int r = blockIdx.y;
    const int b = rows[r];
    const int c = b + 1 + blockIdx.x * 256 + threadIdx.x;
    if (c >= N) return;
    const long long src = static_cast<long long>(r) * N + c;
    const long long dst = static_cast<long long>(b) * N - static_cast<long long>(b) * (b + 1) / 2 + (c - b - 1);
    o_cc[dst] = cc[src];
    o_lag[dst] = lag[src];
    o_sub[dst] = sub[src];
}

void launch_ccx_pack_rows(const double* d_cc, const int* d_lag, const double* d_sub, const int* d_rows, int nrows, int N,
                          double* o_cc, int* o_lag, double* o_sub, cudaStream_t st) {
    const dim3 grid((N - 1 + 255) / 256, nrows);
    ccx_pack_rows_kernel<<<grid, 256, 0, st>>>(d_cc, d_lag, d_sub, d_rows, N, o_cc, o_lag, o_sub);
}

void launch_ccx_pack(const double* d_cc, const int* d_lag, const double* d_sub, const int* d_slot_of_row, int N,
                     double* o_cc, int* o_lag, double* o_sub, cudaStream_t st) {
    const dim3 grid((N - 1 + 255) / 256, N - 1);
    ccx_pack_kernel<<<grid, 256, 0, st>>>(d_cc, d_lag, d_sub, d_slot_of_row, N, o_cc, o_lag, o_sub);
}

}  // namespace dtx

// ------------------------------------------------------------------ zero-lag Pearson matrix
// `SubSpace.validateClusters` (reference detex/subspace.py:738-773) calls
// construct.fast_normcorr(t, s) on every pair of ALIGNED, trimmed waveforms of a cluster; at
// equal lengths that is the single zero-lag Pearson coefficient.  One CTA per pair, float64.
namespace dtx {
namespace {
__global__ void __launch_bounds__(256)
corr0_kernel(const double* __restrict__ X, int N, int n, double* __restrict__ out) {
    const int b = blockIdx.y, c = blockIdx.x;
    if (c < b) return;
    const double* x1 = X + static_cast<long long>(b) * n;
    const double* x2 = X + static_cast<long long>(c) * n;
    double s1 = 0, s2 = 0, q1 = 0, q2 = 0, p = 0;
    for (int i = threadIdx.x; i < n; i += 256) {
        const double a = x1[i], d = x2[i];
        s1 += a; s2 += d; q1 += a * a; q2 += d * d; p += a * d;
    }
    __shared__ double sh[8][5];
    double v[5] = {s1, s2, q1, q2, p};
    for (int k = 0; k < 5; ++k)
        for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
    if ((threadIdx.x & 31) == 0)
        for (int k = 0; k < 5; ++k) sh[threadIdx.x >> 5][k] = v[k];
    __syncthreads();
    if (threadIdx.x == 0) {
        double t[5] = {0, 0, 0, 0, 0};
        for (int w = 0; w < 8; ++w)
            for (int k = 0; k < 5; ++k) t[k] += sh[w][k];
        const double nn = n;
        const double cov = t[4] - t[0] * t[1] / nn;
        const double v1 = t[2] - t[0] * t[0] / nn, v2 = t[3] - t[1] * t[1] / nn;
        const double r = cov / sqrt(v1 * v2);
        out[static_cast<long long>(b) * N + c] = r;
        out[static_cast<long long>(c) * N + b] = r;
    }
}
}  // namespace

void launch_corr0(const double* d_X, int N, int n, double* d_out, cudaStream_t st) {
    const dim3 grid(N, N);
    corr0_kernel<<<grid, 256, 0, st>>>(d_X, N, n, d_out);
}
}  // namespace dtx
