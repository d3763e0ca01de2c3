// k3_post.cu -- K3 / K5: per (chunk, subspace) row of the detection statistic ->
//   MaxDS with the reference's inf-zeroing rule   (detex/detect.py:275-281)
//   400-bin histogram                             (detect.py:178-181; fas.py:79)
//   compacted candidates DS >= threshold          (first stage of _CreateCoeffArray, :410)
//   FAS sufficient statistics N, sum x, sum x^2, sum log x, sum log1p(-x)  (fas.py:81-84,
//     the quantities scipy.stats.beta.fit(floc=0, fscale=1) and beta.nnlf reduce to)
// plus the centred LTA mean of |DS| at each candidate (detect.py:501-524).
//
// HBM-bound: k3_fast_kernel streams each row once (four 128-bit loads in flight per thread,
// float32 pre-binning, per-thread run-length histogram); rows with NaN / inf or too many
// candidates are redone by the two-pass k3_kernel (second pass hits L2 for rows up to ~1.5 MB).
// One CTA per row, shared-memory privatised histogram.
#include "dtx_kernels.cuh"

namespace dtx {
namespace {

__device__ __forceinline__ float warp_max(float v) {
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// np.histogram(x, bins=np.linspace(lo, hi, 401)): bin i holds edges[i] <= x < edges[i+1],
// last bin closed; edges[i] = lo + i*step (float64), edges[400] = hi exactly.
__device__ __forceinline__ int hist_bin(float xf, double lo, double hi, double step, int nb) {
    const double x = static_cast<double>(xf);
    if (!(x >= lo) || !(x <= hi)) return -1;
    int i = static_cast<int>((x - lo) / step);
    if (i > nb - 1) i = nb - 1;
    while (i > 0 && x < lo + i * step) --i;
    while (i < nb - 1 && x >= lo + (i + 1) * step) ++i;
    return i;
}

constexpr int K3_THREADS = 256;

__global__ void __launch_bounds__(K3_THREADS)
k3_kernel(const float* __restrict__ DS, const ChunkDesc* __restrict__ chunks, int S,
          const float* __restrict__ thr, float* __restrict__ rowmax, int* __restrict__ rowflags,
          unsigned long long* __restrict__ hist, double hlo, double hhi,
          Candidate* __restrict__ cand, int cand_cap, int* __restrict__ ncand,
          double* __restrict__ fas, int only_flagged, int nb, int row_base) {
    const ChunkDesc cd = chunks[blockIdx.y];
    const int s = blockIdx.x;
    const int row = row_base + blockIdx.y * S + s;
    if (only_flagged && !(rowflags[row] & 4)) return;  // the single-pass kernel already did this row
    const float* x = DS + cd.ds_off + static_cast<long long>(s) * cd.Tpad + cd.t_lo;   // core lags only
    const int T = cd.t_hi - cd.t_lo;
    const int tid = threadIdx.x, w = tid >> 5, l = tid & 31;
    __shared__ int sh_hist[HIST_MAX_BINS];
    __shared__ float sh_f[8][2];
    __shared__ int sh_i[8][2];
    __shared__ double sh_d[8][4];
    __shared__ float b_maxfin;
    __shared__ int b_nan, b_inf, b_posinf;

    // ---- pass 1: max / nan / inf
    float mfin = -INFINITY;
    int nnan = 0, ninf = 0, npinf = 0;
    const int T4 = T & ~3;
    for (int i = tid * 4; i < T4; i += K3_THREADS * 4) {
        const float4 v = *reinterpret_cast<const float4*>(x + i);
        const float a[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (isnan(a[k])) nnan = 1;
            else if (isinf(a[k])) { ninf = 1; if (a[k] > 0) npinf = 1; }
            else mfin = fmaxf(mfin, a[k]);
        }
    }
    for (int i = T4 + tid; i < T; i += K3_THREADS) {
        const float a = x[i];
        if (isnan(a)) nnan = 1;
        else if (isinf(a)) { ninf = 1; if (a > 0) npinf = 1; }
        else mfin = fmaxf(mfin, a);
    }
    mfin = warp_max(mfin);
    nnan = __any_sync(0xffffffffu, nnan);
    ninf = __any_sync(0xffffffffu, ninf);
    npinf = __any_sync(0xffffffffu, npinf);
    if (l == 0) { sh_f[w][0] = mfin; sh_i[w][0] = nnan | (ninf << 1) | (npinf << 2); }
    for (int i = tid; i < nb; i += K3_THREADS) sh_hist[i] = 0;
    __syncthreads();
    if (tid == 0) {
        float m = -INFINITY; int f = 0;
        for (int i = 0; i < 8; ++i) { m = fmaxf(m, sh_f[i][0]); f |= sh_i[i][0]; }
        b_maxfin = m; b_nan = f & 1; b_inf = (f >> 1) & 1; b_posinf = (f >> 2) & 1;
    }
    __syncthreads();
    const bool has_nan = b_nan != 0;
    // reference: MaxDS = ssd.max(); if MaxDS > 1.1: ssd[isinf(ssd)] = 0; MaxDS = ssd.max()
    const float max_all = b_posinf ? INFINITY : b_maxfin;
    const bool zero_inf = !has_nan && (max_all > 1.1f) && b_inf;
    float eff_max = max_all;
    if (zero_inf) eff_max = fmaxf(b_maxfin, 0.f);
    if (has_nan) eff_max = nanf("");
    if (tid == 0) {
        rowmax[row] = eff_max;
        rowflags[row] = (has_nan ? 1 : 0) | (zero_inf ? 2 : 0);
    }
    if (has_nan) return;  // np.histogram raises -> chunk's histogram skipped; NaN max never triggers

    // ---- pass 2: histogram, candidates, FAS sums
    const double step = (hhi - hlo) / nb;
    const float th = thr ? thr[s] : INFINITY;
    const bool trig = eff_max > th;  // _evalTrigCon: strict >
    double f1 = 0, f2 = 0, f3 = 0, f4 = 0;
    for (int i = tid; i < T; i += K3_THREADS) {
        float a = x[i];
        if (zero_inf && isinf(a)) a = 0.f;
        const int b = hist_bin(a, hlo, hhi, step, nb);
        if (b >= 0) atomicAdd(&sh_hist[b], 1);
        if (trig && a >= th) {
            const int k = atomicAdd(ncand, 1);
            if (k < cand_cap) {
                Candidate c; c.row = row; c.t = cd.t_lo + i; c.ds = a; c.lta = 0.f;
                cand[k] = c;
            }
        }
        if (fas) {
            const double d = static_cast<double>(a);
            f1 += d; f2 += d * d;
            // float32 DS can round to exactly 0 (or 1) where float64 would not: keep the logs finite
            f3 += static_cast<double>(logf(fmaxf(a, 1e-30f)));
            f4 += static_cast<double>(log1pf(-fminf(a, 0.99999994f)));
        }
    }
    __syncthreads();
    for (int i = tid; i < nb; i += K3_THREADS) {
        const int c = sh_hist[i];
        if (c) atomicAdd(&hist[static_cast<long long>(s) * HIST_MAX_BINS + i], static_cast<unsigned long long>(c));
    }
    if (fas) {
        f1 = warp_sum(f1); f2 = warp_sum(f2); f3 = warp_sum(f3); f4 = warp_sum(f4);
        if (l == 0) { sh_d[w][0] = f1; sh_d[w][1] = f2; sh_d[w][2] = f3; sh_d[w][3] = f4; }
        __syncthreads();
        if (tid < 4) {
            double t = 0;
            for (int i = 0; i < 8; ++i) t += sh_d[i][tid];
            atomicAdd(&fas[s * 5 + 1 + tid], t);
        }
        if (tid == 4) atomicAdd(&fas[s * 5], static_cast<double>(T));
    }
}


// Single-pass version for the common case (no NaN / inf in the row): 128-bit loads, per-thread
// run-length histogram (noise DS sits in one or two bins, so almost no shared atomics),
// candidates staged in shared memory.  Rows with NaN / inf or more than K3_STAGE candidates
// are flagged (bit 2) and redone by k3_kernel, which carries the reference's corner rules.
constexpr int K3_STAGE = 1024;
template <bool FAS>
__global__ void __launch_bounds__(K3_THREADS)
k3_fast_kernel(const float* __restrict__ DS, const ChunkDesc* __restrict__ chunks, int S,
               const float* __restrict__ thr, float* __restrict__ rowmax, int* __restrict__ rowflags,
               unsigned long long* __restrict__ hist, double hlo, double hhi,
               Candidate* __restrict__ cand, int cand_cap, int* __restrict__ ncand,
               double* __restrict__ fas, int nb, int row_base) {
    const ChunkDesc cd = chunks[blockIdx.y];
    const int s = blockIdx.x;
    const int row = row_base + blockIdx.y * S + s;
    const float* x = DS + cd.ds_off + static_cast<long long>(s) * cd.Tpad + cd.t_lo;   // core lags only (t_lo % 4 == 0)
    const int T = cd.t_hi - cd.t_lo;
    const int tid = threadIdx.x, w = tid >> 5, l = tid & 31;
    __shared__ int sh_hist[HIST_MAX_BINS];
    __shared__ int2 sh_cand[K3_STAGE];
    __shared__ int sh_ncand, sh_bad, sh_base;
    __shared__ float sh_max[8];
    __shared__ double sh_d[8][4];
    for (int i = tid; i < nb; i += K3_THREADS) sh_hist[i] = 0;
    if (tid == 0) { sh_ncand = 0; sh_bad = 0; }
    __syncthreads();
    const double step = (hhi - hlo) / nb;
    const float flo = static_cast<float>(hlo), finv = static_cast<float>(nb / (hhi - hlo));
    const float th = thr ? thr[s] : INFINITY;
    float mfin = -INFINITY;
    int bad = 0, cur = -1, cnt = 0;
    double f1 = 0, f2 = 0, f3 = 0, f4 = 0;
    auto fas_add = [&](float a) {
        const double d = static_cast<double>(a);
        f1 += d; f2 += d * d;
        // float32 DS can round to exactly 0 (or 1) where float64 would not: keep the logs finite
        f3 += static_cast<double>(logf(fmaxf(a, 1e-30f)));
        f4 += static_cast<double>(log1pf(-fminf(a, 0.99999994f)));
    };
    auto one = [&](float a, int i) {
        if (isnan(a) || isinf(a)) { bad = 1; return; }
        mfin = fmaxf(mfin, a);
        const int b = hist_bin_fast(a, flo, finv, hlo, hhi, step, nb);
        if (b == cur) ++cnt;
        else {
            if (cnt > 0 && cur >= 0) atomicAdd(&sh_hist[cur], cnt);
            cur = b;
            cnt = 1;
        }
        if (a >= th) {
            const int k = atomicAdd(&sh_ncand, 1);
            if (k < K3_STAGE) sh_cand[k] = make_int2(i, __float_as_int(a));
        }
        if (FAS) fas_add(a);
    };
    // Four values at once, branch free: if all four are finite, sit well inside ONE bin (more than
    // 1e-3 of a bin away from its edges, so float32 rounding cannot move them), that bin is the
    // current run's and none reaches the threshold, the run just grows by four.  Noise DS does
    // that almost always; everything else takes the per-value path above.
    const float toff = -flo * finv;
    auto quad = [&](const float4 v, int i) {
        const float curf = static_cast<float>(cur);
        const float t0 = fmaf(v.x, finv, toff), t1 = fmaf(v.y, finv, toff), t2 = fmaf(v.z, finv, toff),
                    t3 = fmaf(v.w, finv, toff);
        const float g0 = floorf(t0), g1 = floorf(t1), g2 = floorf(t2), g3 = floorf(t3);
        const bool inside = fabsf(t0 - g0 - 0.5f) < 0.499f && fabsf(t1 - g1 - 0.5f) < 0.499f &&
                            fabsf(t2 - g2 - 0.5f) < 0.499f && fabsf(t3 - g3 - 0.5f) < 0.499f;
        const bool same = cur >= 0 && g0 == curf && g1 == curf && g2 == curf && g3 == curf;
        const float m4 = fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w));
        if (inside && same && m4 < th) {
            cnt += 4;
            mfin = fmaxf(mfin, m4);
            if (FAS) {
                // quad partial sums in float32 (4 terms), accumulated in float64
                f1 += static_cast<double>((v.x + v.y) + (v.z + v.w));
                f2 += static_cast<double>(fmaf(v.x, v.x, v.y * v.y) + fmaf(v.z, v.z, v.w * v.w));
                f3 += static_cast<double>((logf(fmaxf(v.x, 1e-30f)) + logf(fmaxf(v.y, 1e-30f))) +
                                          (logf(fmaxf(v.z, 1e-30f)) + logf(fmaxf(v.w, 1e-30f))));
                f4 += static_cast<double>((log1pf(-fminf(v.x, 0.99999994f)) + log1pf(-fminf(v.y, 0.99999994f))) +
                                          (log1pf(-fminf(v.z, 0.99999994f)) + log1pf(-fminf(v.w, 0.99999994f))));
            }
        } else {
            one(v.x, i); one(v.y, i + 1); one(v.z, i + 2); one(v.w, i + 3);
        }
    };
    const int T4 = T & ~3;
    constexpr int STRIDE = K3_THREADS * 4;
    int i0 = tid * 4;
    // four independent 128-bit loads in flight per thread
    for (; i0 + 3 * STRIDE < T4; i0 += 4 * STRIDE) {
        const float4 v0 = __ldcs(reinterpret_cast<const float4*>(x + i0));
        const float4 v1 = __ldcs(reinterpret_cast<const float4*>(x + i0 + STRIDE));
        const float4 v2 = __ldcs(reinterpret_cast<const float4*>(x + i0 + 2 * STRIDE));
        const float4 v3 = __ldcs(reinterpret_cast<const float4*>(x + i0 + 3 * STRIDE));
        quad(v0, i0);
        quad(v1, i0 + STRIDE);
        quad(v2, i0 + 2 * STRIDE);
        quad(v3, i0 + 3 * STRIDE);
    }
    for (int i = i0; i < T4; i += STRIDE) quad(*reinterpret_cast<const float4*>(x + i), i);
    for (int i = T4 + tid; i < T; i += K3_THREADS) one(x[i], i);
    if (cnt > 0 && cur >= 0) atomicAdd(&sh_hist[cur], cnt);
    mfin = warp_max(mfin);
    if (__any_sync(0xffffffffu, bad) && l == 0) sh_bad = 1;
    if (l == 0) sh_max[w] = mfin;
    if (FAS) {
        f1 = warp_sum(f1); f2 = warp_sum(f2); f3 = warp_sum(f3); f4 = warp_sum(f4);
        if (l == 0) { sh_d[w][0] = f1; sh_d[w][1] = f2; sh_d[w][2] = f3; sh_d[w][3] = f4; }
    }
    __syncthreads();
    const int nc = sh_ncand;
    const int badrow = sh_bad;
    __syncthreads();                        // (sh_base below shares a vector load with these two)
    if (badrow || nc > K3_STAGE) {          // hand the row to the two-pass kernel
        if (tid == 0) rowflags[row] = 4;
        return;
    }
    float mx = sh_max[0];
    for (int i = 1; i < 8; ++i) mx = fmaxf(mx, sh_max[i]);
    const bool trig = mx > th;               // _evalTrigCon: strict >
    if (tid == 0) {
        rowmax[row] = mx;
        rowflags[row] = 0;
        sh_base = (trig && nc > 0) ? atomicAdd(ncand, nc) : 0;
    }
    for (int i = tid; i < nb; i += K3_THREADS) {
        const int c = sh_hist[i];
        if (c) atomicAdd(&hist[static_cast<long long>(s) * HIST_MAX_BINS + i], static_cast<unsigned long long>(c));
    }
    if (fas) {
        if (tid < 4) {
            double t = 0;
            for (int i = 0; i < 8; ++i) t += sh_d[i][tid];
            atomicAdd(&fas[s * 5 + 1 + tid], t);
        }
        if (tid == 4) atomicAdd(&fas[s * 5], static_cast<double>(T));
    }
    __syncthreads();
    if (trig)
        for (int i = tid; i < nc; i += K3_THREADS) {
            const int k = sh_base + i;
            if (k < cand_cap) {
                Candidate c; c.row = row; c.t = cd.t_lo + sh_cand[i].x; c.ds = __int_as_float(sh_cand[i].y); c.lta = 0.f;
                cand[k] = c;
            }
        }
}

// One warp per candidate: centred rolling mean of |DS| with pandas' window placement and
// _replaceNanWithMean's edge rule (detect.py:517-524).  `x` is the DS row, `i` the lag.
__device__ __forceinline__ float centred_abs_mean(const float* __restrict__ x, int T, int W, int i,
                                                  bool zero_inf, int l) {
    if (T < W) return nanf("");
    const int off = (W - 1) / 2;
    const int first = W - 1 - off, last = T - 1 - off;  // valid centres
    if (i < first) i = (first + 1 <= last) ? first + 1 : first;
    if (i > last) i = last;
    const int a0 = i - (W - 1) + off;
    double acc = 0;
    for (int j = l; j < W; j += 32) {
        float v = x[a0 + j];
        if (zero_inf && isinf(v)) v = 0.f;
        acc += fabs(static_cast<double>(v));
    }
    acc = warp_sum(acc);
    return static_cast<float>(acc / W);
}

// cand.lta = the denominator of DS_STALTA (`_getStaLtaArray`, detect.py:501-515):
//   Wsta == 0 (reference default, STA = |DS|):  LTA mean;
//   Wsta  > 0:  LTA mean * |DS[t]| / STA mean, so that |DS[t]| / lta == STA / LTA in both cases.
__global__ void __launch_bounds__(256)
lta_kernel(const float* __restrict__ DS, const ChunkDesc* __restrict__ chunks, int S,
           const int* __restrict__ rowflags, Candidate* __restrict__ cand,
           const int* __restrict__ ncand, const int* __restrict__ ncand_before, int cand_cap, int row_base, int W,
           int Wsta) {
    const int n = min(*ncand, cand_cap);
    const int wid = (ncand_before ? *ncand_before : 0) + ((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int l = threadIdx.x & 31;
    if (wid >= n) return;
    Candidate c = cand[wid];
    const int lrow = c.row - row_base;          // row inside the current batch
    const ChunkDesc cd = chunks[lrow / S];
    const float* x = DS + cd.ds_off + static_cast<long long>(lrow % S) * cd.Tpad;
    // An inf anywhere in a row makes its MaxDS inf > 1.1, so the reference has zeroed every inf of the
    // row before it forms the LTA (detect.py:275-288): infs count as 0 here whether they sit in this
    // segment's core (rowflags bit 1) or only in its halo.
    const bool zero_inf = true;
    (void)rowflags;
    float out = centred_abs_mean(x, cd.T, W, c.t, zero_inf, l);
    if (Wsta > 0) {
        const float sta = centred_abs_mean(x, cd.T, Wsta, c.t, zero_inf, l);
        out = static_cast<float>(static_cast<double>(out) * fabs(static_cast<double>(c.ds)) /
                                 static_cast<double>(sta));
    }
    if (l == 0) cand[wid].lta = out;
}


// Dense STA/LTA of one DS row (`_getStaLtaArray`, STA = 0 -> |DS|; detect.py:501-515):
// out[i] = |DS[i]| / LTA[i_eff], LTA the centred rolling mean of |DS| over W samples with
// the `_replaceNanWithMean` edge rule.  Only the CorDF.STALTA column needs it; the trigger
// path uses lta_kernel on the sparse candidates.  One block per 1024 outputs, |DS| staged
// in shared memory as a float64 prefix sum.
constexpr int SL_TILE = 1024;
__global__ void __launch_bounds__(256)
stalta_dense_kernel(const float* __restrict__ x, int T, int W, int zero_inf, int mean_only,
                    float* __restrict__ out) {
    extern __shared__ double pre[];  // [SL_TILE + W + 1]
    const int off = (W - 1) / 2;
    const int first = W - 1 - off, last = T - 1 - off;
    const int i0 = blockIdx.x * SL_TILE;
    // centres needed by this tile after the edge rule: clamp(i) for i in [i0, i0+SL_TILE)
    int clo = i0, chi = min(i0 + SL_TILE, T) - 1;
    auto eff = [&](int i) {
        if (i < first) return (first + 1 <= last) ? first + 1 : first;
        if (i > last) return last;
        return i;
    };
    // span of centres this tile can touch (eff(i) >= clamp(i); i < first maps to first+1)
    const int cmin = min(max(clo, first), last);
    int cmax = min(max(chi, first), last);
    if (clo < first) cmax = max(cmax, min(first + 1, last));
    const int a0 = cmin - (W - 1) + off;  // first sample needed
    const int a1 = cmax + off;            // last sample needed
    const int cnt = a1 - a0 + 1;
    for (int j = threadIdx.x; j < cnt; j += 256) {
        float v = x[a0 + j];
        if (zero_inf && isinf(v)) v = 0.f;
        pre[j + 1] = fabs(static_cast<double>(v));
    }
    if (threadIdx.x == 0) pre[0] = 0.0;
    __syncthreads();
    {   // inclusive prefix sum of pre[1..cnt]: serial inside a thread's segment, block scan of the totals
        __shared__ double seg_tot[256];
        const int per = (cnt + 255) / 256;
        const int j0 = 1 + threadIdx.x * per, j1 = min(cnt + 1, j0 + per);
        double run = 0.0;
        for (int j = j0; j < j1; ++j) { run += pre[j]; pre[j] = run; }
        seg_tot[threadIdx.x] = run;
        __syncthreads();
        if (threadIdx.x < 32) {   // one warp scans the 256 totals, 8 per lane
            double t[8], acc = 0.0;
            for (int k = 0; k < 8; ++k) { acc += seg_tot[threadIdx.x * 8 + k]; t[k] = acc; }
            double incl = acc;
            for (int o = 1; o < 32; o <<= 1) {
                const double v = __shfl_up_sync(0xffffffffu, incl, o);
                if (threadIdx.x >= o) incl += v;
            }
            const double excl = incl - acc;
            for (int k = 0; k < 8; ++k) seg_tot[threadIdx.x * 8 + k] = excl + t[k];
        }
        __syncthreads();
        const double off = threadIdx.x ? seg_tot[threadIdx.x - 1] : 0.0;
        for (int j = j0; j < j1; ++j) pre[j] += off;
    }
    __syncthreads();
    for (int k = threadIdx.x; k < SL_TILE; k += 256) {
        const int i = i0 + k;
        if (i >= T) break;
        const int c = eff(i);
        const int s0 = c - (W - 1) + off - a0;
        const double lta = (pre[s0 + W] - pre[s0]) / W;
        float v = x[i];
        if (zero_inf && isinf(v)) v = 0.f;
        out[i] = mean_only ? static_cast<float>(lta) : static_cast<float>(fabs(static_cast<double>(v)) / lta);
    }
}

// STA / LTA from the two dense rolling means (triggerSTATime != 0)
__global__ void __launch_bounds__(256)
ratio_kernel(const float* num, const float* den, int T, float* out) {  // out may alias den
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i < T) out[i] = static_cast<float>(static_cast<double>(num[i]) / static_cast<double>(den[i]));
}

}  // namespace

void launch_k3(const float* DS, const ChunkDesc* d_chunks, int nchunks, int S, const float* d_thr,
               float* d_rowmax, int* d_rowflags, unsigned long long* d_hist, double hist_lo,
               double hist_hi, int nbins, Candidate* d_cand, int cand_cap, int* d_ncand, double* d_fas,
               int row_base, cudaStream_t st) {
    const dim3 grid(S, nchunks);
    if (d_fas)
        k3_fast_kernel<true><<<grid, K3_THREADS, 0, st>>>(DS, d_chunks, S, d_thr, d_rowmax, d_rowflags, d_hist,
                                                          hist_lo, hist_hi, d_cand, cand_cap, d_ncand, d_fas, nbins, row_base);
    else
        k3_fast_kernel<false><<<grid, K3_THREADS, 0, st>>>(DS, d_chunks, S, d_thr, d_rowmax, d_rowflags, d_hist,
                                                           hist_lo, hist_hi, d_cand, cand_cap, d_ncand, nullptr, nbins, row_base);
    k3_kernel<<<grid, K3_THREADS, 0, st>>>(DS, d_chunks, S, d_thr, d_rowmax, d_rowflags, d_hist, hist_lo, hist_hi,
                                           d_cand, cand_cap, d_ncand, d_fas, 1, nbins, row_base);
}

void launch_lta(const float* DS, const ChunkDesc* d_chunks, int S, const int* d_rowflags,
                Candidate* d_cand, const int* d_ncand, const int* d_ncand_before, int cand_cap, int row_base, int W,
                int Wsta, cudaStream_t st) {
    // grid sized for the capacity; warps beyond *ncand exit immediately
    const int warps = cand_cap;
    const int grid = (warps * 32 + 255) / 256;
    lta_kernel<<<grid, 256, 0, st>>>(DS, d_chunks, S, d_rowflags, d_cand, d_ncand, d_ncand_before, cand_cap, row_base,
                                     W, Wsta);
}

static void launch_rolling(const float* row, int T, int W, int zero_inf, int mean_only, float* out,
                           cudaStream_t st) {
    const int grid = (T + SL_TILE - 1) / SL_TILE;
    const size_t sm = sizeof(double) * (SL_TILE + 2 * W + 2);
    cudaFuncSetAttribute(stalta_dense_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sm));
    stalta_dense_kernel<<<grid, 256, sm, st>>>(row, T, W, zero_inf, mean_only, out);
}

// largest rolling window (samples) the dense STA/LTA kernel can stage in shared memory
int stalta_dense_max_window() { return (227 * 1024 / static_cast<int>(sizeof(double)) - SL_TILE - 2) / 2; }

// Wsta == 0: out = |DS| / LTA.  Wsta > 0: out = STA / LTA, `tmp` (T floats) holds the STA means.
void launch_stalta_dense(const float* row, int T, int W, int Wsta, int zero_inf, float* out, float* tmp,
                         cudaStream_t st) {
    if (T < W || T < Wsta) return;
    if (Wsta <= 0) {
        launch_rolling(row, T, W, zero_inf, 0, out, st);
        return;
    }
    launch_rolling(row, T, Wsta, zero_inf, 1, tmp, st);
    launch_rolling(row, T, W, zero_inf, 1, out, st);
    ratio_kernel<<<(T + 255) / 256, 256, 0, st>>>(tmp, out, T, out);
}

}  // namespace dtx
