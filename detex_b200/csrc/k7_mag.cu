// k7_mag.cu -- N1: per-detection magnitude / SNR estimates, `_SSDetex._estMag`
// (reference detex/detect.py:447-499) with `_estPEMag` / `_estSTDMag` (:637-664) and
// `construct.fast_normcorr` (construct.py:469-483).
//
// Sparse work (one CTA per trigger), all float64, operating on the raw chunk that is already
// resident for the detection run:
//   ConDat = MPcon[t*Nc : t*Nc+n]
//   ssCon  = U^T (U ConDat)                     (the reference's UtU . ConDat without the n x n UtU)
//   proEn  = var(ssCon) / var(WFU_i)            per event
//   noise  = median(rolling_std(pe, n)), pe = the 5n samples before the trigger
//            (or the 7n samples from it if there are not 5n before)      -> SNR = std(ConDat)/noise
//   cor_i  = Pearson(event_i, ConDat)           (fast_normcorr at equal lengths)
//   ProEnMag = sum_i (mag_i + log10 sqrt(proEn_i)) cor_i^2 / sum cor_i^2      (mag_i > -15)
//   Mag      = sum_i (mag_i + log10(std(ConDat)/std(event_i))) cor_i^2 / sum cor_i^2
// Singles (detect.py:489-498): ProEnMag = mag + (ConDat.WFU0)/(WFU0.WFU0),
//                              Mag = mag + log10(std(ConDat)/std(WFU0)).
#include "dtx_kernels.cuh"

namespace dtx {
namespace {

constexpr int MT = 256;

__device__ double block_sum(double v, double* sh) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0;
    for (int i = 0; i < MT / 32; ++i) t += sh[i];
    return t;
}

// k-th smallest (0-based) of n non-negative doubles by MSB-first 8-bit radix select
__device__ double kth_smallest(const double* v, int n, int k, unsigned* hist /*[256] smem*/) {
    unsigned long long prefix = 0, mask = 0;
    for (int shift = 56; shift >= 0; shift -= 8) {
        for (int i = threadIdx.x; i < 256; i += MT) hist[i] = 0;
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += MT) {
            const unsigned long long key = static_cast<unsigned long long>(__double_as_longlong(v[i]));
            if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 0xff], 1u);
        }
        __syncthreads();
        // all threads walk the histogram identically
        int d = 0;
        unsigned acc = 0;
        for (; d < 256; ++d) {
            if (acc + hist[d] > static_cast<unsigned>(k)) break;
            acc += hist[d];
        }
        k -= static_cast<int>(acc);
        prefix |= static_cast<unsigned long long>(d) << shift;
        mask |= 0xffull << shift;
        __syncthreads();
    }
    return __longlong_as_double(static_cast<long long>(prefix));
}

template <typename T>
__global__ void __launch_bounds__(MT)
mag_kernel(const T* __restrict__ raw, const ChunkDesc* __restrict__ chunks, const double* __restrict__ sum,
           const MagTrigger* __restrict__ trig, const MagSubspace* __restrict__ subs, const double* __restrict__ U,
           int n, int Nc, double* __restrict__ scratch, int scratch_stride, double* __restrict__ out) {
    const MagTrigger tg = trig[blockIdx.x];
    const MagSubspace sp = subs[tg.subspace];
    const ChunkDesc cd = chunks[tg.chunk];
    const T* x = raw + cd.raw_off;
    const long long o = static_cast<long long>(tg.t) * Nc;
    const int tid = threadIdx.x;
    __shared__ double sh[MT / 32];
    __shared__ unsigned hist[256];
    __shared__ double coef[MAG_MAX_RANK];
    double* scr = scratch + static_cast<long long>(blockIdx.x) * scratch_stride;
    const double nn = static_cast<double>(n);

    // ---- ConDat statistics
    double a = 0;
    for (int j = tid; j < n; j += MT) a += static_cast<double>(x[o + j]);
    const double mean_s = block_sum(a, sh) / nn;
    a = 0;
    for (int j = tid; j < n; j += MT) {
        const double d = static_cast<double>(x[o + j]) - mean_s;
        a += d * d;
    }
    const double std_s = sqrt(block_sum(a, sh) / nn);  // np.std (population)

    // ---- noise level: median of rolling_std (ddof = 1) over the pre-event (or post-event) span
    long long p0;
    int plen;
    if (o > 5LL * n) { p0 = o - 5LL * n; plen = 5 * n; }
    else { p0 = o; plen = 7 * n; }
    if (p0 + plen > cd.L) plen = static_cast<int>(cd.L - p0);   // reference slices silently truncate
    double noise = nan("");
    const int nwin = plen - n + 1;
    if (nwin >= 1) {
        const double cm = sum[tg.chunk] / static_cast<double>(cd.L);  // conditioning only
        // window sums by chunks of MT windows: thread w handles windows [w*per, (w+1)*per)
        const int per = (nwin + MT - 1) / MT;
        const int w0 = tid * per, w1 = min(nwin, w0 + per);
        if (w0 < w1) {
            double s1 = 0, s2 = 0;
            for (int j = 0; j < n; ++j) {
                const double d = static_cast<double>(x[p0 + w0 + j]) - cm;
                s1 += d; s2 += d * d;
            }
            for (int wv = w0; wv < w1; ++wv) {
                double var = (s2 - s1 * s1 / nn) / (nn - 1.0);
                if (var < 0) var = 0;
                scr[wv] = sqrt(var);
                if (wv + 1 < w1) {
                    const double din = static_cast<double>(x[p0 + wv + n]) - cm;
                    const double dout = static_cast<double>(x[p0 + wv]) - cm;
                    s1 += din - dout; s2 += din * din - dout * dout;
                }
            }
        }
        __syncthreads();
        __threadfence_block();
        if (nwin & 1) noise = kth_smallest(scr, nwin, nwin / 2, hist);
        else noise = 0.5 * (kth_smallest(scr, nwin, nwin / 2 - 1, hist) + kth_smallest(scr, nwin, nwin / 2, hist));
    }
    const double snr = std_s / noise;

    double peMag = nan(""), stMag = nan("");
    if (sp.is_single) {
        const double* wfu = sp.ewf;  // WFU[0] == the single's trimmed waveform
        double d1 = 0, d2 = 0, m = 0;
        for (int j = tid; j < n; j += MT) {
            const double wv = wfu[j];
            d1 += static_cast<double>(x[o + j]) * wv; d2 += wv * wv; m += wv;
        }
        d1 = block_sum(d1, sh); d2 = block_sum(d2, sh); m = block_sum(m, sh) / nn;
        double v = 0;
        for (int j = tid; j < n; j += MT) { const double d = wfu[j] - m; v += d * d; }
        const double std_w = sqrt(block_sum(v, sh) / nn);
        const double mg = sp.mags[0];
        if (!(isnan(mg) || mg < -15)) {
            peMag = mg + d1 / d2;
            stMag = mg + log10(std_s / std_w);
        }
    } else {
        // coefficients a_k = U_k . ConDat, ssCon = sum_k a_k U_k
        const double* Us = U + static_cast<long long>(sp.row0) * n;
        for (int k = 0; k < sp.rank; ++k) {
            double d = 0;
            for (int j = tid; j < n; j += MT) d += Us[static_cast<long long>(k) * n + j] * static_cast<double>(x[o + j]);
            d = block_sum(d, sh);
            if (tid == 0) coef[k] = d;
        }
        __syncthreads();
        double s1 = 0, s2 = 0;
        for (int j = tid; j < n; j += MT) {
            double ss = 0;
            for (int k = 0; k < sp.rank; ++k) ss += coef[k] * Us[static_cast<long long>(k) * n + j];
            s1 += ss; s2 += ss * ss;
        }
        s1 = block_sum(s1, sh); s2 = block_sum(s2, sh);
        const double var_ss = s2 / nn - (s1 / nn) * (s1 / nn);
        double num_pe = 0, num_st = 0, den = 0;
        bool any = false;
        for (int i = 0; i < sp.nev; ++i) {
            const double* e = sp.ewf + static_cast<long long>(i) * n;
            double d = 0;
            for (int j = tid; j < n; j += MT) d += e[j] * static_cast<double>(x[o + j]);
            d = block_sum(d, sh);
            const double cor = (d - nn * sp.ev_mean[i] * mean_s) / (nn * sp.ev_std[i] * std_s);
            const double we = cor * cor;
            if (sp.mags[i] > -15) {
                any = true;
                den += we;
                num_pe += (sp.mags[i] + log10(sqrt(var_ss / sp.wfu_var[i]))) * we;
                num_st += (sp.mags[i] + log10(std_s / sp.ev_std[i])) * we;
            }
        }
        if (any) { peMag = num_pe / den; stMag = num_st / den; }
    }
    if (tid == 0) {
        out[blockIdx.x * 3 + 0] = peMag;   // ProEnMag
        out[blockIdx.x * 3 + 1] = stMag;   // Mag
        out[blockIdx.x * 3 + 2] = snr;     // SNR
    }
}

}  // namespace

void launch_mag(const void* raw, int dtype_f32, const ChunkDesc* d_chunks, const double* d_sum,
                const MagTrigger* d_trig, int ntrig, const MagSubspace* d_subs, const double* d_U, int n, int Nc,
                double* d_scratch, int scratch_stride, double* d_out, cudaStream_t st) {
    if (ntrig < 1) return;
    if (dtype_f32)
        mag_kernel<float><<<ntrig, MT, 0, st>>>(static_cast<const float*>(raw), d_chunks, d_sum, d_trig, d_subs, d_U, n,
                                                Nc, d_scratch, scratch_stride, d_out);
    else
        mag_kernel<double><<<ntrig, MT, 0, st>>>(static_cast<const double*>(raw), d_chunks, d_sum, d_trig, d_subs, d_U,
                                                 n, Nc, d_scratch, scratch_stride, d_out);
}

}  // namespace dtx
