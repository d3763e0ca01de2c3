// k8_preproc.cu -- N2: the array part of `construct._applyFilter` (reference
// detex/construct.py:990-1030) followed by `multiplex` (:928-987), on the device, so that raw
// per-channel traces go H2D once and come out as the multiplexed chunk K0/K1 consume:
//   st.detrend('linear')                      -> scipy.signal.detrend(type='linear') per trace
//   st.filter('bandpass', corners, zerophase) -> ObsPy 1.0.2 obspy/signal/filter.py::bandpass:
//        sos = zpk2sos(iirfilter(corners, [lo, hi], 'band', 'butter')); y = sosfilt(sos, x);
//        zerophase: y = sosfilt(sos, y[::-1])[::-1]      (zero initial state, no edge padding)
//   multiplex: trim to the shortest channel, interleave [c0[0], c1[0], c2[0], c0[1], ...]
// The SOS coefficients are designed on the host (scalar work) and passed in.
//
// A biquad (direct form II transposed) is the linear recurrence s' = A s + B x, y = b0 x + s0.
// It is parallelised over segments of SEG samples: pass 1 runs every segment from the zero
// state to get its end state, a tiny sequential pass propagates the true initial states with
// A^SEG, pass 2 re-runs every segment from its true initial state.  float64 throughout; the
// result equals the sequential recursion up to round-off (1e-15 relative).
#include "dtx_kernels.cuh"

namespace dtx {
namespace {

constexpr int SEG = 256;

struct Biquad { double b0, b1, b2, a1, a2; };

// ------------------------------------------------------------------ detrend (linear LSQ)
__global__ void __launch_bounds__(256)
pp_detrend_stats(const double* __restrict__ buf, const long long* __restrict__ off, const int* __restrict__ len,
                 double* __restrict__ stats /*[ntr][2]*/) {
    const int tr = blockIdx.y;
    const int N = len[tr];
    const double* y = buf + off[tr];
    const double tbar = 0.5 * (N + 1.0) / N;   // mean of t_i = (i+1)/N
    double s0 = 0, s1 = 0;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < N; i += gridDim.x * 256) {
        const double v = y[i];
        s0 += v;
        s1 += ((i + 1.0) / N - tbar) * v;
    }
    for (int o = 16; o > 0; o >>= 1) {
        s0 += __shfl_xor_sync(0xffffffffu, s0, o);
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    }
    __shared__ double sh[8][2];
    if ((threadIdx.x & 31) == 0) { sh[threadIdx.x >> 5][0] = s0; sh[threadIdx.x >> 5][1] = s1; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0, b = 0;
        for (int i = 0; i < 8; ++i) { a += sh[i][0]; b += sh[i][1]; }
        atomicAdd(&stats[tr * 2 + 0], a);
        atomicAdd(&stats[tr * 2 + 1], b);
    }
}

__global__ void __launch_bounds__(256)
pp_detrend_apply(double* __restrict__ buf, const long long* __restrict__ off, const int* __restrict__ len,
                 const double* __restrict__ stats) {
    const int tr = blockIdx.y;
    const int N = len[tr];
    double* y = buf + off[tr];
    const double tbar = 0.5 * (N + 1.0) / N;
    const double stt = (static_cast<double>(N) * N - 1.0) / (12.0 * N);   // sum (t_i - tbar)^2
    const double slope = N > 1 ? stats[tr * 2 + 1] / stt : 0.0;
    const double icpt = stats[tr * 2 + 0] / N - slope * tbar;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < N; i += gridDim.x * 256)
        y[i] -= slope * ((i + 1.0) / N) + icpt;
}

// -------------------------------------------------------------------------- biquad passes
// direction dir = +1: samples 0..N-1 ; dir = -1: samples N-1..0 (the zero-phase backward pass)
__device__ __forceinline__ int sidx(int k, int N, int dir) { return dir > 0 ? k : N - 1 - k; }

__global__ void __launch_bounds__(128)
pp_biquad_pass1(const double* __restrict__ buf, const long long* __restrict__ off, const int* __restrict__ len,
                Biquad q, int dir, double* __restrict__ segstate /*[ntr][maxseg][2]*/, int maxseg) {
    const int tr = blockIdx.y;
    const int N = len[tr];
    const int sg = blockIdx.x * 128 + threadIdx.x;
    if (sg * SEG >= N) return;
    const double* x = buf + off[tr];
    double z1 = 0, z2 = 0;
    const int k1 = min(N, (sg + 1) * SEG);
    for (int k = sg * SEG; k < k1; ++k) {
        const double xv = x[sidx(k, N, dir)];
        const double yv = q.b0 * xv + z1;
        z1 = q.b1 * xv - q.a1 * yv + z2;
        z2 = q.b2 * xv - q.a2 * yv;
    }
    segstate[(static_cast<long long>(tr) * maxseg + sg) * 2 + 0] = z1;
    segstate[(static_cast<long long>(tr) * maxseg + sg) * 2 + 1] = z2;
}

// one thread per trace: true initial state of every segment.  A^SEG is built by squaring.
__global__ void pp_biquad_scan(const int* __restrict__ len, Biquad q, double* __restrict__ segstate, int maxseg,
                               int ntr) {
    const int tr = blockIdx.x * blockDim.x + threadIdx.x;
    if (tr >= ntr) return;
    const int nseg = (len[tr] + SEG - 1) / SEG;
    // A = [[-a1, 1], [-a2, 0]]
    double m00 = -q.a1, m01 = 1.0, m10 = -q.a2, m11 = 0.0;
    for (int s = 1; s < SEG; s <<= 1) {   // SEG is a power of two
        const double n00 = m00 * m00 + m01 * m10, n01 = m00 * m01 + m01 * m11;
        const double n10 = m10 * m00 + m11 * m10, n11 = m10 * m01 + m11 * m11;
        m00 = n00; m01 = n01; m10 = n10; m11 = n11;
    }
    double* st = segstate + static_cast<long long>(tr) * maxseg * 2;
    double i1 = 0, i2 = 0;   // state entering segment 0
    for (int s = 0; s < nseg; ++s) {
        const double e1 = st[s * 2], e2 = st[s * 2 + 1];   // zero-state end state of segment s
        st[s * 2] = i1;
        st[s * 2 + 1] = i2;
        const double t1 = m00 * i1 + m01 * i2 + e1;        // full segments only matter (last is unused)
        const double t2 = m10 * i1 + m11 * i2 + e2;
        i1 = t1; i2 = t2;
    }
}

__global__ void __launch_bounds__(128)
pp_biquad_pass2(double* __restrict__ buf, const long long* __restrict__ off, const int* __restrict__ len, Biquad q,
                int dir, const double* __restrict__ segstate, int maxseg) {
    const int tr = blockIdx.y;
    const int N = len[tr];
    const int sg = blockIdx.x * 128 + threadIdx.x;
    if (sg * SEG >= N) return;
    double* x = buf + off[tr];
    double z1 = segstate[(static_cast<long long>(tr) * maxseg + sg) * 2 + 0];
    double z2 = segstate[(static_cast<long long>(tr) * maxseg + sg) * 2 + 1];
    const int k1 = min(N, (sg + 1) * SEG);
    for (int k = sg * SEG; k < k1; ++k) {
        const int i = sidx(k, N, dir);
        const double xv = x[i];
        const double yv = q.b0 * xv + z1;
        z1 = q.b1 * xv - q.a1 * yv + z2;
        z2 = q.b2 * xv - q.a2 * yv;
        x[i] = yv;   // in place: a segment only touches its own samples
    }
}

// ------------------------------------------------------------------------------ decimate
// data[::factor] of every trace (ObsPy Trace.decimate after its low-pass) into a second buffer
__global__ void __launch_bounds__(256)
pp_decimate(const double* __restrict__ src, const long long* __restrict__ off, double* __restrict__ dst,
            const long long* __restrict__ off2, const int* __restrict__ len2, int factor) {
    const int tr = blockIdx.y;
    const int n = len2[tr];
    const double* s = src + off[tr];
    double* d = dst + off2[tr];
    for (int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256)
        d[i] = s[static_cast<long long>(i) * factor];
}

// ------------------------------------------------------------------------------ multiplex
__global__ void __launch_bounds__(256)
pp_multiplex(const double* __restrict__ buf, const long long* __restrict__ off, const int* __restrict__ minlen,
             const long long* __restrict__ out_off, int Nc, double* __restrict__ out) {
    const int ch = blockIdx.y;   // chunk
    const int n = minlen[ch];
    double* o = out + out_off[ch];
    const long long total = static_cast<long long>(n) * Nc;
    for (long long j = blockIdx.x * 256LL + threadIdx.x; j < total; j += gridDim.x * 256LL) {
        const int c = static_cast<int>(j % Nc);
        const int i = static_cast<int>(j / Nc);
        o[j] = buf[off[ch * Nc + c] + i];
    }
}

}  // namespace

void launch_preproc(double* d_buf, const long long* d_off, const int* d_len, int ntr, int maxlen, const double* sos,
                    int nsos, int zerophase, int detrend, double* d_stats, double* d_segstate, cudaStream_t st) {
    if (detrend) {
        cudaMemsetAsync(d_stats, 0, sizeof(double) * 2 * ntr, st);
        const dim3 g(32, ntr);
        pp_detrend_stats<<<g, 256, 0, st>>>(d_buf, d_off, d_len, d_stats);
        pp_detrend_apply<<<g, 256, 0, st>>>(d_buf, d_off, d_len, d_stats);
    }
    const int maxseg = (maxlen + SEG - 1) / SEG;
    const dim3 gs((maxseg + 127) / 128, ntr);
    for (int dir = 1; dir >= (zerophase ? -1 : 1); dir -= 2)
        for (int s = 0; s < nsos; ++s) {
            // sos row = [b0 b1 b2 a0 a1 a2], a0 == 1
            Biquad q{sos[s * 6 + 0] / sos[s * 6 + 3], sos[s * 6 + 1] / sos[s * 6 + 3], sos[s * 6 + 2] / sos[s * 6 + 3],
                     sos[s * 6 + 4] / sos[s * 6 + 3], sos[s * 6 + 5] / sos[s * 6 + 3]};
            pp_biquad_pass1<<<gs, 128, 0, st>>>(d_buf, d_off, d_len, q, dir, d_segstate, maxseg);
            pp_biquad_scan<<<(ntr + 63) / 64, 64, 0, st>>>(d_len, q, d_segstate, maxseg, ntr);
            pp_biquad_pass2<<<gs, 128, 0, st>>>(d_buf, d_off, d_len, q, dir, d_segstate, maxseg);
        }
}

void launch_multiplex(const double* d_buf, const long long* d_off, const int* d_minlen, const long long* d_out_off,
                      int nchunks, int Nc, int maxlen, double* d_out, cudaStream_t st) {
    const dim3 g(64, nchunks);
    pp_multiplex<<<g, 256, 0, st>>>(d_buf, d_off, d_minlen, d_out_off, Nc, d_out);
}

void launch_decimate(const double* d_src, const long long* d_off, double* d_dst, const long long* d_off2,
                     const int* d_len2, int ntr, int factor, cudaStream_t st) {
    const dim3 g(64, ntr);
    pp_decimate<<<g, 256, 0, st>>>(d_src, d_off, d_dst, d_off2, d_len2, factor);
}

int preproc_seg() { return SEG; }

}  // namespace dtx
