// afftdn -- FFT spectral-subtraction denoiser (libavfilter/af_afftdn.c), fltp path:
// "afftdn=nr=12:nt=w|custom:bn=<15 dB>:tn=0|1[:nf=<dB>]" (reference: filters.go:830-861,
// adaptive.go:133-170; 15 band centres analyser_noise_bands.go:15-17).
// window = 3 * (fs/80) samples of wscale*sin^2, hop fs/80, 2048-point real transform (at 44.1/48 kHz),
// per-bin a-priori-SNR gain with a Bark-band masking floor; output delayed by window-hop.
//
// The filter's only cross-hop state is (a) prior[bin], a contraction (factor <= 0.39 per hop),
// (b) prior_band_excit[band], a contraction (factor beta <= 0.78), (c) the tracked noise floor
// (tn=1), which depends on the INPUT spectra only.  So the per-hop work is cut along its data
// dependences instead of being walked hop by hop:
//   F1  forward FFT, two hops per transform (hop 2p as the real part, hop 2p+1 as the imaginary part of
//       one complex Stockham radix-4 FFT in shared memory, separated by symmetry); with tn=1 also each
//       hop's flatness test and floor candidate.  Fully parallel over hop pairs.
//   F2  noise-floor recurrence over hops as an affine scan (tn=1)
//   F3a a-priori-SNR gain recursion: one thread per BIN walks a chunk of hops (+128 warm-up hops),
//       coalesced across bins -> gain[hop][bin], clean[hop][bin]
//   F3b band excitation: per-band sums (bin order, as the scalar code), the per-band recursion over hops
//       (+128 warm-up hops) and the spreading matrix -> amt[hop][band]
//   F3c gain limiting, spectrum scaling and the inverse FFT, again two hops per transform -> windowed frames
//   F4  overlap-add of the three frames covering each output sample (f64, hop order)
#include "jt_internal.h"
#include "jt_device.cuh"
#include "jt_fft.cuh"
#include <cstdio>
#include <algorithm>
#include <cstring>
#include <vector>

#define AF_THREADS 512
#define AF_MAXBANDS 64
#define AF_MAXOWN 5
#define AF_C (M_LN10 * 0.1)

struct AfConst {
    int A, W, FL, FL2, bins, nbands;
    double floor_, gain_scale, max_gain, ratio, floor_offset;
};

// block-wide sum / max of NV doubles per thread at once (one barrier pair)
template <int NV, bool MAX>
__device__ __forceinline__ void af_block_reduce(double (&v)[NV], double (*red)[8])
{
#pragma unroll
    for (int i = 0; i < NV; i++) v[i] = MAX ? jt_warp_max(v[i]) : jt_warp_sum(v[i]);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int i = 0; i < NV; i++) red[threadIdx.x >> 5][i] = v[i];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NV; i++) {
        double s = 0;
        for (int w = 0; w < AF_THREADS / 32; w++) s = MAX ? fmax(s, red[w][i]) : s + red[w][i];
        v[i] = s;
    }
}

// F1: forward transform of hop pairs; optional noise-floor candidates (track_noise)
__global__ void __launch_bounds__(AF_THREADS, 2)
k_afftdn_fwd(const float *__restrict__ x, int64_t n, int64_t n_hops, AfConst K, const double *__restrict__ window,
             const float2 *__restrict__ tw_g, float2 *__restrict__ spectra, int track, double *__restrict__ cand /* 2 per hop: new_floor, flag */)
{
    extern __shared__ float2 sm[];
    __shared__ double red[AF_THREADS / 32][8];
    const int N = K.FL, t = threadIdx.x;
    float2 *bufA = sm, *bufB = sm + N, *tw = sm + 2 * N;
    for (int i = t; i < N; i += AF_THREADS) tw[i] = tw_g[i];
    const int64_t n_pairs = (n_hops + 1) / 2;
    for (int64_t pr = blockIdx.x; pr < n_pairs; pr += gridDim.x) {
        const int64_t h0 = 2 * pr, h1 = h0 + 1;
        const bool has1 = h1 < n_hops;
        const int64_t wa = (h0 - 2) * (int64_t)K.A, ava = min(n, (h0 + 1) * (int64_t)K.A);
        const int64_t wb = (h1 - 2) * (int64_t)K.A, avb = min(n, (h1 + 1) * (int64_t)K.A);
        __syncthreads();
        for (int m = t; m < N; m += AF_THREADS) {
            float a = 0.f, b = 0.f;
            if (m < K.W) {
                const double wv = window[m];
                const int64_t sa = wa + m, sb = wb + m;
                if (sa >= 0 && sa < ava) a = (float)(wv * (double)x[sa] * 8388608.0);
                if (has1 && sb >= 0 && sb < avb) b = (float)(wv * (double)x[sb] * 8388608.0);
            }
            bufA[m] = make_float2(a, b);
        }
        __syncthreads();
        const float2 *Z = af_fft<false>(bufA, bufB, tw, N);
        // separate the two real transforms: A[k] = (Z[k] + conj(Z[N-k])) / 2, B[k] = (Z[k] - conj(Z[N-k])) / 2i
        double mga[AF_MAXOWN], mgb[AF_MAXOWN];
#pragma unroll
        for (int j = 0; j < AF_MAXOWN; j++) {
            const int k = t + j * AF_THREADS;
            mga[j] = mgb[j] = -1.0;
            if (k < K.bins) {
                const float2 zk = Z[k], zn = Z[(N - k) & (N - 1)];
                const float2 fa = make_float2(0.5f * (zk.x + zn.x), 0.5f * (zk.y - zn.y));
                const float2 fb = make_float2(0.5f * (zk.y + zn.y), 0.5f * (zn.x - zk.x));
                spectra[h0 * (int64_t)K.bins + k] = fa;
                if (has1) spectra[h1 * (int64_t)K.bins + k] = fb;
                if (track) { mga[j] = hypot((double)fa.x, (double)fa.y); mgb[j] = hypot((double)fb.x, (double)fb.y); }
            }
        }
        if (track) {
            double sums[6] = {0, 0, 0, 0, 0, 0};          // num, den, count for each of the two hops
#pragma unroll
            for (int j = 0; j < AF_MAXOWN; j++) {
                if (mga[j] > K.floor_) { sums[0] += log(mga[j]); sums[1] += mga[j]; sums[2] += 1; }
                if (mgb[j] > K.floor_) { sums[3] += log(mgb[j]); sums[4] += mgb[j]; sums[5] += 1; }
            }
            af_block_reduce<6, false>(sums, red);
            const double size_a = fmax(sums[2], 1.0), size_b = fmax(sums[5], 1.0);
            const double num_a = exp(sums[0] / size_a), den_a = sums[1] / size_a;
            const double num_b = exp(sums[3] / size_b), den_b = sums[4] / size_b;
            double off[2] = {0, 0};
#pragma unroll
            for (int j = 0; j < AF_MAXOWN; j++) {
                if (mga[j] >= 0) off[0] = fmax(off[0], fabs(mga[j] - den_a));
                if (mgb[j] >= 0) off[1] = fmax(off[1], fabs(mgb[j] - den_b));
            }
            af_block_reduce<2, true>(off, red);
            if (t == 0) {
                double nf = 10.0 * log10(den_a) - 100.0 + K.floor_offset * (off[0] / den_a);
                cand[2 * h0] = fmin(fmax(nf, -90.), -20.); cand[2 * h0 + 1] = (num_a / den_a > 0.8) ? 1.0 : 0.0;
                if (has1) {
                    nf = 10.0 * log10(den_b) - 100.0 + K.floor_offset * (off[1] / den_b);
                    cand[2 * h1] = fmin(fmax(nf, -90.), -20.); cand[2 * h1 + 1] = (num_b / den_b > 0.8) ? 1.0 : 0.0;
                }
            }
        }
    }
}

// F2: max_var before / after each hop's floor update
// The floor recurrence nf' = flag ? 0.1*cand + 0.9*nf : nf is affine in nf, so one CTA scans it:
// each thread composes its slice of hops into (A, B), thread 0 chains the 1024 slices, every
// thread replays its slice from its true entry value.
__global__ void __launch_bounds__(1024)
k_afftdn_floor(const double *__restrict__ cand, int64_t n_hops, double nf0, double floor_, int track,
               double *__restrict__ mv_pre, double *__restrict__ mv_post)
{
    __shared__ double sA[1024], sB[1024], sIn[1024];
    const int t = threadIdx.x;
    const int64_t per = (n_hops + 1023) / 1024, h0 = t * per, h1 = min(h0 + per, n_hops);
    double A = 1.0, B = 0.0;
    if (track) for (int64_t h = h0; h < h1; h++) if (cand[2 * h + 1] != 0.0) { A *= 0.9; B = 0.1 * cand[2 * h] + B * 0.9; }
    sA[t] = A; sB[t] = B;
    __syncthreads();
    if (t == 0) { double nf = nf0; for (int i = 0; i < 1024; i++) { sIn[i] = nf; nf = sA[i] * nf + sB[i]; } }
    __syncthreads();
    double nf = sIn[t];
    for (int64_t h = h0; h < h1; h++) {
        mv_pre[h] = floor_ * exp((100.0 + nf) * AF_C);
        if (track && cand[2 * h + 1] != 0.0) nf = 0.1 * cand[2 * h] + nf * 0.9;
        mv_post[h] = floor_ * exp((100.0 + nf) * AF_C);
    }
}

__device__ __forceinline__ double af_limit_gain(double a, double b)
{
    if (a > 1.0) return (b * a - 1.0) / (b + a - 2.0);
    if (a < 1.0) return (b * a - 2.0 * a + 1.0) / (b - a);
    return 1.0;
}

// F3a: a-priori SNR recursion, one thread per bin over a chunk of hops (warm-up hops first)
#define AF_GAIN_THREADS 128
__global__ void __launch_bounds__(AF_GAIN_THREADS)
k_afftdn_gain(const float2 *__restrict__ spectra, int64_t n_hops, int chunk, int warm, AfConst K,
              const double *__restrict__ rel_var, const double *__restrict__ mv_pre,
              double *__restrict__ gain, double *__restrict__ clean)
{
    const int i = blockIdx.x * AF_GAIN_THREADS + threadIdx.x;
    if (i >= K.bins) return;
    const int64_t h_out0 = (int64_t)blockIdx.y * chunk;
    if (h_out0 >= n_hops) return;
    const int64_t h_out1 = min(h_out0 + chunk, n_hops), h_begin = max((int64_t)0, h_out0 - warm);
    const double rv = rel_var[i];
    double prior = 0.0;
    auto step = [&](int64_t h, float2 sp) {
        const double ratio = (h == 0) ? 1.0 : K.ratio, rratio = 1.0 - ratio;
        const double mag = hypot((double)sp.x, (double)sp.y), power = mag * mag;
        const double abs_var = fmax(mv_pre[h] * rv, 1.0);
        const double mav = power / abs_var;
        const double nmav = ratio * prior + rratio * fmax(mav - 1.0, 0.0);
        const double g = nmav / (1.0 + nmav), sg = g * g;
        prior = mav * sg;
        if (h >= h_out0) { gain[h * (int64_t)K.bins + i] = g; clean[h * (int64_t)K.bins + i] = power * sg; }
    };
    int64_t h = h_begin;
    for (; h + 4 <= h_out1; h += 4) {             // four loads in flight ahead of the carried chain
        float2 sp[4];
#pragma unroll
        for (int u = 0; u < 4; u++) sp[u] = spectra[(h + u) * (int64_t)K.bins + i];
#pragma unroll
        for (int u = 0; u < 4; u++) step(h + u, sp[u]);
    }
    for (; h < h_out1; h++) step(h, spectra[h * (int64_t)K.bins + i]);
}

// F3b-1: band excitation before the recursion: raw[hop][band] = sum of clean over the band's bins, in bin order
__global__ void __launch_bounds__(256)
k_afftdn_bandsum(const double *__restrict__ clean, int64_t n_hops, AfConst K, const int *__restrict__ band_lo /* nbands+1 */,
                 double *__restrict__ raw)
{
    const int lane = threadIdx.x & 31;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t h = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; h < n_hops; h += warps) {
        const double *row = clean + h * (int64_t)K.bins;
        for (int b = lane; b < K.nbands; b += 32) {
            double be = 0.0;
            const int lo = band_lo[b], hi = band_lo[b + 1];
            for (int i = lo; i < hi; i++) be += row[i];
            raw[h * (int64_t)K.nbands + b] = be;
        }
    }
}

// F3b-2: per-band recursion over a chunk of hops (warm-up hops first), then the spreading matrix
__global__ void __launch_bounds__(256)
k_afftdn_bandrec(const double *__restrict__ raw, int64_t n_hops, int chunk, int warm, AfConst K,
                 const double *__restrict__ band_alpha, const double *__restrict__ band_beta, const double *__restrict__ spread,
                 double *__restrict__ amt)
{
    extern __shared__ double sbe[];                         // (chunk + warm) x nbands, then nbands x nbands
    const int nb = K.nbands, t = threadIdx.x;
    double *ssp = sbe + (size_t)(chunk + warm) * nb;
    const int64_t h_out0 = (int64_t)blockIdx.x * chunk;
    if (h_out0 >= n_hops) return;
    const int64_t h_out1 = min(h_out0 + chunk, n_hops), h_begin = max((int64_t)0, h_out0 - warm);
    const int nh = (int)(h_out1 - h_begin);
    for (int i = t; i < nh * nb; i += blockDim.x) sbe[i] = raw[h_begin * nb + i];
    for (int i = t; i < nb * nb; i += blockDim.x) ssp[i] = spread[i];
    __syncthreads();
    if (t < nb) {
        const double alpha = band_alpha[t], beta = band_beta[t];
        double prev = 0.0;
        for (int h = 0; h < nh; h++) {
            double be = sbe[h * nb + t];
            be = fmax(be, alpha * be + beta * prev);
            prev = be;
            sbe[h * nb + t] = be;
        }
    }
    __syncthreads();
    const int skip = (int)(h_out0 - h_begin);
    for (int i = t; i < (nh - skip) * nb; i += blockDim.x) {
        const int h = skip + i / nb, j = i % nb;
        double a = 0.0;
        for (int k = 0; k < nb; k++) a += ssp[j * nb + k] * sbe[h * nb + k];
        amt[(h_begin + h) * nb + j] = a;
    }
}

// F3c: gain limiting against the masking threshold, spectrum scaling, inverse transform of hop pairs
__global__ void __launch_bounds__(AF_THREADS)
k_afftdn_synth(const float2 *__restrict__ spectra, const double *__restrict__ gain, const double *__restrict__ amt,
               int64_t n_hops, AfConst K, const double *__restrict__ rel_var, const int *__restrict__ bin2band,
               const double *__restrict__ mv_post, const float2 *__restrict__ tw_g, float *__restrict__ frames)
{
    extern __shared__ float2 sm[];
    const int N = K.FL, t = threadIdx.x;
    float2 *bufA = sm, *bufB = sm + N, *tw = sm + 2 * N;
    for (int i = t; i < N; i += AF_THREADS) tw[i] = tw_g[i];
    const int64_t n_pairs = (n_hops + 1) / 2;
    for (int64_t pr = blockIdx.x; pr < n_pairs; pr += gridDim.x) {
        const int64_t h0 = 2 * pr;
        const bool has1 = h0 + 1 < n_hops;
        __syncthreads();
        for (int k = t; k < K.bins; k += AF_THREADS) {
            const double rv = rel_var[k];
            const int band = bin2band[k];
            float2 X[2];
#pragma unroll
            for (int u = 0; u < 2; u++) {
                X[u] = make_float2(0.f, 0.f);
                const int64_t h = h0 + u;
                if (u == 1 && !has1) continue;
                const float2 sp = spectra[h * (int64_t)K.bins + k];
                double g = gain[h * (int64_t)K.bins + k];
                const double am = amt[h * (int64_t)K.nbands + band];
                const double abs_var = fmax(mv_post[h] * rv, 1.0), min_abs_var = K.gain_scale * abs_var;
                if (am > abs_var) g = 1.0;
                else if (am > min_abs_var) g = af_limit_gain(g, sqrt(abs_var / am));
                else g = af_limit_gain(g, K.max_gain);
                const float ng = (float)g;
                X[u] = make_float2(__fmul_rn(sp.x, ng), __fmul_rn(sp.y, ng));
                if (k == 0 || k == K.FL2) X[u].y = 0.f;
            }
            // Z = Xa + i*Xb on the lower half, conj(Xa) + i*conj(Xb) mirrored: Re(ifft) = frame a, Im(ifft) = frame b
            bufA[k] = make_float2(X[0].x - X[1].y, X[0].y + X[1].x);
            if (k > 0 && k < K.FL2) bufA[N - k] = make_float2(X[0].x + X[1].y, X[1].x - X[0].y);
        }
        __syncthreads();
        const float2 *z = af_fft<true>(bufA, bufB, tw, N);
        for (int m = t; m < K.W; m += AF_THREADS) {
            const float2 v = z[m];
            frames[h0 * (int64_t)K.W + m] = v.x;
            if (has1) frames[(h0 + 1) * (int64_t)K.W + m] = v.y;
        }
    }
}

// F4: overlap-add, out[q] = sum over the (up to) three hops covering q, in hop order, f64
__global__ void __launch_bounds__(256)
k_afftdn_ola(const float *__restrict__ frames, const double *__restrict__ window, int64_t n, int64_t n_hops, int A, int W,
             float *__restrict__ y)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += stride) {
        const int64_t g0 = q / A;
        double acc = 0.0;
        for (int64_t g = max((int64_t)0, g0 - 2); g <= g0 && g < n_hops; g++) {
            const int m = (int)(q - g * A);
            if (m < W) acc += window[m] * (double)frames[g * (int64_t)W + m] / 8388608.0;
        }
        y[q] = (float)acc;
    }
}

// ---------------------------------------------------------------------------------------
// host: tables (af_afftdn.c config_input / set_parameters / set_band_parameters)
// ---------------------------------------------------------------------------------------
static const int kBandCentre[15] = {80, 125, 195, 290, 440, 660, 1000, 1500, 2250, 3350, 5000, 7500, 11200, 16000, 24000};
static double freq2bark(double x) { double d = x / 7500.0; return 13.0 * atan(7.6E-4 * x) + 3.5 * atan(d * d); }

static double band_noise_extrapolated(const double *bn, double sample_rate)
{
    // least-squares degree-4 polynomial through the 15 band values, evaluated past the last band
    double ma[25], mb[75], vb[5];
    for (int j = 0; j < 5; j++) for (int k = 0; k < 5; k++) { ma[j + k * 5] = 0; for (int m = 0; m < 15; m++) ma[j + k * 5] += pow(m, j + k); }
    for (int i = 0; i < 4; i++) for (int j = i + 1; j < 5; j++) { double d = ma[j + i * 5] / ma[i + i * 5]; ma[j + i * 5] = d; for (int k = i + 1; k < 5; k++) ma[j + k * 5] -= d * ma[i + k * 5]; }
    { int i = 0; for (int j = 0; j < 5; j++) for (int k = 0; k < 15; k++) mb[i++] = pow(k, j); }
    { int i = 0; for (int j = 0; j < 5; j++) { double s = 0; for (int k = 0; k < 15; k++) s += mb[i++] * bn[k]; vb[j] = s; } }
    for (int i = 0; i < 4; i++) for (int j = i + 1; j < 5; j++) vb[j] -= ma[j + i * 5] * vb[i];
    vb[4] /= ma[24];
    for (int i = 3; i >= 0; i--) { double d = vb[i]; for (int j = i + 1; j < 5; j++) d -= ma[i + j * 5] * vb[j]; vb[i] = d / ma[i + i * 5]; }
    double f = (0.5 * sample_rate) / kBandCentre[14];
    f = 15.0 + log(f / 1.5) / log(1.5);
    double sum = 0, prod = 1; for (int j = 0; j < 5; j++) { sum += prod * vb[j]; prod *= f; }
    return sum;
}

// One chunk's view of the noise-floor recurrence nf' = flag ? 0.1*cand + 0.9*nf : nf: the affine map (A, B) of a run
// of hops, composed on the host from the per-hop candidates.
static void floor_affine(const double *cand, int64_t h0, int64_t h1, double &A, double &B)
{
    A = 1.0; B = 0.0;
    for (int64_t h = h0; h < h1; h++) if (cand[2 * h + 1] != 0.0) { A *= 0.9; B = 0.1 * cand[2 * h] + B * 0.9; }
}
struct FloorCarryRec { int64_t key; double A, B; int64_t valid; };

// transform geometry, window and floor constant: functions of the rate (and of ad / fo, which ride along in K)
static void afftdn_geometry(int rate, const AfftdnParams &P, AfConst &K, std::vector<double> &window, std::vector<float2> &tw)
{
    const double sample_rate = (float)rate;
    K.A = (int)(sample_rate / 80); K.W = 3 * K.A;
    K.FL = 1; while (K.FL <= K.W) K.FL <<= 1;
    K.FL2 = K.FL / 2; K.bins = K.FL2 + 1;
    if (K.FL < 512 || K.bins > AF_MAXOWN * AF_THREADS || K.FL > 8192) JT_THROW(JT_ERR_UNSUPPORTED, "afftdn at %d Hz (transform length %d)", rate, K.FL);
    K.ratio = P.ad; K.floor_offset = P.fo; K.nbands = 0; K.gain_scale = K.max_gain = 0;
    window.resize(K.W); double sum = 0;
    { const double wscale = sqrt(8.0 / (9.0 * K.FL)); for (int i = 0; i < K.W; i++) { double d = sin(i * M_PI / K.W); d *= wscale * d; window[i] = d; sum += d * d; } }
    K.floor_ = (double)(1LL << 48) * exp(-23.025558369790467) * (0.5 * sum);
    tw.resize(K.FL);                                // full circle: the radix-4 passes use w, w^2, w^3
    for (int k = 0; k < K.FL; k++) { const double a = -2.0 * M_PI * k / K.FL; tw[k] = make_float2((float)cos(a), (float)sin(a)); }
}
static int afftdn_fft_grid(jt_ctx *c, const AfConst &K, int64_t n_hops, size_t &smem_fft)
{
    smem_fft = sizeof(float2) * 3 * (size_t)K.FL;
    const int64_t n_pairs = (n_hops + 1) / 2;
    const int fft_per_sm = (int)std::max<size_t>(1, std::min<size_t>(4, (200 * 1024) / (smem_fft + 1024)));
    // CTAs that live for ~8 transform pairs, not for a sixteenth of the stream: the block scheduler can only hand SM slots to a
    // higher-priority stream (the band graphs the host is waiting for while this kernel runs in Pass 2's head) when CTAs retire
    return (int)std::min<int64_t>(n_pairs, (int64_t)c->num_sms * fft_per_sm * 32);
}

void jt_afftdn_forward(jt_ctx *c, const Sig &in, const AfftdnParams &P, AfftdnFwd &out)
{
    out = AfftdnFwd();
    if (in.fmt != JT_FMT_FLT || in.n <= 0) return;
    AfConst K; std::vector<double> window; std::vector<float2> tw;
    afftdn_geometry(in.rate, P, K, window, tw);
    const int64_t n_hops = (in.n + K.A - 1) / K.A;
    const double *d_window = jt_dev_table(c, "afftdn_window", window);
    const float2 *d_tw = jt_dev_table(c, "afftdn_tw", tw);
    float2 *d_spec = jt_dalloc<float2>(c, (size_t)n_hops * K.bins);
    double *d_cand = jt_dalloc<double>(c, (size_t)n_hops * 2);
    size_t smem_fft; const int grid_fft = afftdn_fft_grid(c, K, n_hops, smem_fft);
    jt_smem_optin((const void *)k_afftdn_fwd, smem_fft);
    { JtLaunch L(c, "afftdn:fwd");
      k_afftdn_fwd<<<grid_fft, AF_THREADS, smem_fft, c->stream>>>((const float *)in.d, in.n, n_hops, K, d_window, d_tw, d_spec, P.tn, d_cand); }
    out.d_spec = d_spec; out.d_cand = d_cand; out.src = in.d; out.n = in.n; out.n_hops = n_hops; out.rate = in.rate; out.tn = P.tn; out.fo = P.fo;
}

Sig jt_afftdn(jt_ctx *c, const Sig &in, const AfftdnParams &P, const AfftdnCarry *carry, const AfftdnFwd *fwd)
{
    if (in.fmt != JT_FMT_FLT) JT_THROW(JT_ERR_INVALID_ARG, "afftdn expects float input");
    Sig o = in; o.d = jt_dalloc<float>(c, in.n);
    if (in.n <= 0) return o;
    AfConst K; std::vector<double> window; std::vector<float2> tw;
    afftdn_geometry(in.rate, P, K, window, tw);
    const double sample_rate = (float)in.rate;
    std::vector<int> bin2band(K.bins);
    for (int i = 0; i < K.bins; i++) bin2band[i] = (int)lrint(P.bm * freq2bark((0.5 * i * sample_rate) / K.FL2));
    K.nbands = bin2band[K.bins - 1] + 1;
    if (K.nbands > AF_MAXBANDS) JT_THROW(JT_ERR_UNSUPPORTED, "afftdn band_multiplier %g gives %d bands", P.bm, K.nbands);
    const int nb = K.nbands;
    std::vector<int> band_lo(nb + 1, K.bins);
    for (int i = K.bins - 1; i >= 0; i--) band_lo[bin2band[i]] = i;
    for (int b = nb - 1; b >= 0; b--) if (band_lo[b] == K.bins) band_lo[b] = band_lo[b + 1];
    band_lo[nb] = K.bins;
    // band noise profile
    double bn[15] = {0};
    if (P.nt == 3 && P.has_bn) for (int i = 0; i < 15; i++) { double v = (float)P.bn[i]; bn[i] = v < -24. ? -24. : v > 24. ? 24. : v; }
    else if (P.nt == 1 || P.nt == 2) for (int i = 0; i < 15; i++) {
        const double a = P.nt == 1 ? 50.0 : 1.0, b = P.nt == 1 ? 500.5 : 500.0, cc = P.nt == 1 ? 2125.0 : 1.0E10;
        double d1 = a / kBandCentre[i]; d1 = 10.0 * log(1.0 + d1 * d1) / M_LN10;
        double d2 = b / kBandCentre[i]; d2 = 10.0 * log(1.0 + d2 * d2) / M_LN10;
        double d3 = kBandCentre[i] / cc; d3 = 10.0 * log(1.0 + d3 * d3) / M_LN10;
        bn[i] = -d1 + d2 - d3;
    }
    { double mean = 0; for (double v : bn) mean += v; mean /= 15; for (double &v : bn) v -= mean; }
    // spread function, normalised as config_input does
    std::vector<double> spread((size_t)nb * nb), cnt(nb, 0.0), pbe(nb, 0.0), be(nb);
    { const double p1 = pow(0.1, 2.5 / P.bm), p2 = pow(0.1, 1.0 / P.bm); int j = 0;
      for (int m = 0; m < nb; m++) for (int n2 = 0; n2 < nb; n2++) spread[j++] = n2 < m ? pow(p2, m - n2) : n2 > m ? pow(p1, n2 - m) : 1.0;
      for (int m = 0; m < K.bins; m++) cnt[bin2band[m]] += 1.0;
      j = 0; for (int m = 0; m < nb; m++) for (int n2 = 0; n2 < nb; n2++) pbe[m] += spread[j++] * cnt[n2];
      const double mn = pow(0.1, 2.5), mx = pow(0.1, 1.0);
      for (int i = 0; i < nb; i++) { double v = i < lrint(12.0 * P.bm) ? pow(0.1, 1.45 + 0.1 * i / P.bm) : pow(0.1, 2.5 - 0.2 * (i / P.bm - 14.0)); be[i] = v < mn ? mn : v > mx ? mx : v; }
      j = 0; for (int i = 0; i < nb; i++) for (int k = 0; k < nb; k++) spread[j++] *= be[i] / pbe[i]; }
    std::vector<double> alpha(nb, 0.0), beta(nb, 0.0);
    { int j = 0; const double sar = K.A / sample_rate;
      for (int i = 0; i < K.bins; i++) if (i == K.FL2 || bin2band[i] > j) {
          const double d6 = (i - 1) * sample_rate / K.FL, d7 = std::fmin(0.008 + 2.2 / d6, 0.03);
          alpha[j] = exp(-sar / d7); beta[j] = 1.0 - alpha[j]; j = bin2band[i]; } }
    // rel_var (set_band_parameters)
    std::vector<double> rel_var(K.bins);
    { double band_noise = bn[0], d2 = 1, d5 = 0; int i = 0, j = 0, k = 0;
      for (int m = 0; m < K.bins; m++) {
          if (m == j) {
              i = j; d5 = band_noise;
              j = k >= 15 ? K.bins : (int)(K.FL * kBandCentre[k] / sample_rate);
              d2 = j - i;
              band_noise = k < 15 ? bn[k] : band_noise_extrapolated(bn, sample_rate);
              k++;
          }
          const double d3 = (j - m) / d2, d4 = (m - i) / d2;
          rel_var[m] = exp((d5 * d3 + band_noise * d4) * AF_C);
      } }
    K.max_gain = exp(P.nr * (0.5 * AF_C)); K.gain_scale = 1.0 / (K.max_gain * K.max_gain);

    const int64_t n_hops = (in.n + K.A - 1) / K.A;
    const double *d_window = jt_dev_table(c, "afftdn_window", window);
    const float2 *d_tw = jt_dev_table(c, "afftdn_tw", tw);
    const double *d_rel = jt_upload_params(c, rel_var);          // follows the file's band-noise profile: a per-call table
    const int *d_blo = jt_dev_table(c, "afftdn_bandlo", band_lo);
    const int *d_b2b = jt_dev_table(c, "afftdn_bin2band", bin2band);
    const double *d_alpha = jt_dev_table(c, "afftdn_alpha", alpha), *d_beta = jt_dev_table(c, "afftdn_beta", beta);
    const double *d_spread = jt_dev_table(c, "afftdn_spread", spread);
    // forward transforms already made from this very signal (jt_afftdn_forward)?
    const bool have_fwd = fwd && fwd->d_spec && fwd->src == in.d && fwd->n == in.n && fwd->n_hops == n_hops && fwd->rate == in.rate && fwd->tn == P.tn && fwd->fo == P.fo;
    float2 *d_spec = have_fwd ? (float2 *)fwd->d_spec : jt_dalloc<float2>(c, (size_t)n_hops * K.bins);
    double *d_gain = jt_dalloc<double>(c, (size_t)n_hops * K.bins), *d_clean = jt_dalloc<double>(c, (size_t)n_hops * K.bins);
    double *d_raw = jt_dalloc<double>(c, (size_t)n_hops * nb), *d_amt = jt_dalloc<double>(c, (size_t)n_hops * nb);
    float *d_frames = jt_dalloc<float>(c, (size_t)n_hops * K.W);
    double *d_cand = have_fwd ? fwd->d_cand : jt_dalloc<double>(c, (size_t)n_hops * 2), *d_pre = jt_dalloc<double>(c, n_hops), *d_post = jt_dalloc<double>(c, n_hops);
    size_t smem_fft; const int grid_fft = afftdn_fft_grid(c, K, n_hops, smem_fft);
    jt_smem_optin((const void *)k_afftdn_fwd, (size_t)(smem_fft));
    jt_smem_optin((const void *)k_afftdn_synth, (size_t)(smem_fft));
    const int warm = 128;                             // 0.39^128, 0.78^128: both recursions have forgotten their start
    const int chunk_gain = 768, chunk_band = 256;
    const size_t smem_band = sizeof(double) * ((size_t)(chunk_band + warm) * nb + (size_t)nb * nb);
    jt_smem_optin((const void *)k_afftdn_bandrec, (size_t)(smem_band));
    {
        if (!have_fwd) { JtLaunch L(c, "afftdn:fwd");
          k_afftdn_fwd<<<grid_fft, AF_THREADS, smem_fft, c->stream>>>((const float *)in.d, in.n, n_hops, K, d_window, d_tw, d_spec, P.tn, d_cand); }
        double nf_start = P.nf;
        if (carry && P.tn) {
            // the floor entering the owned hops = the stream's initial floor pushed through every earlier chunk's
            // affine carry; ONE exchange of (key, A, B) per chunk, the only data-path collective of the chain
            std::vector<double> h_cand((size_t)n_hops * 2);
            JT_CUDA(cudaMemcpyAsync(h_cand.data(), d_cand, sizeof(double) * 2 * n_hops, cudaMemcpyDeviceToHost, c->stream));
            JT_CUDA(cudaStreamSynchronize(c->stream));
            const int64_t h0 = std::min(std::max<int64_t>(carry->hop0, 0), n_hops), h1 = std::min(std::max(carry->hop1, h0), n_hops);
            FloorCarryRec mine; mine.key = carry->key; mine.valid = 1;
            floor_affine(h_cand.data(), h0, h1, mine.A, mine.B);
            double entry = P.nf;
            if (carry->fn) {
                const int nr = std::max(carry->n_ranks, 1);
                std::vector<FloorCarryRec> all((size_t)nr);
                memset(all.data(), 0, sizeof(FloorCarryRec) * nr);
                if (carry->fn(carry->user, &mine, (int64_t)sizeof(mine), all.data()) != 0) JT_THROW(JT_ERR_INVALID_ARG, "afftdn noise-floor exchange failed");
                std::vector<FloorCarryRec> before;
                for (const FloorCarryRec &r : all) if (r.valid && r.key < mine.key) before.push_back(r);
                std::sort(before.begin(), before.end(), [](const FloorCarryRec &a, const FloorCarryRec &b) { return a.key < b.key; });
                for (const FloorCarryRec &r : before) entry = r.A * entry + r.B;
            } else if (h0 > 0) JT_THROW(JT_ERR_INVALID_ARG, "afftdn track_noise in a mid-stream chunk needs the exchange callback (jt_set_exchange)");
            // start value for the window's first hop such that the recurrence arrives at `entry` on hop h0 (the context
            // hops before h0 only warm up the contractive gain recursions)
            double Ah, Bh; floor_affine(h_cand.data(), 0, h0, Ah, Bh);
            nf_start = Ah > 1e-9 ? (entry - Bh) / Ah : P.nf;
        }
        { JtLaunch L(c, "afftdn:floor");
          k_afftdn_floor<<<1, 1024, 0, c->stream>>>(d_cand, n_hops, nf_start, K.floor_, P.tn, d_pre, d_post); }
        { JtLaunch L(c, "afftdn:gain");
          dim3 g((K.bins + AF_GAIN_THREADS - 1) / AF_GAIN_THREADS, (unsigned)((n_hops + chunk_gain - 1) / chunk_gain));
          k_afftdn_gain<<<g, AF_GAIN_THREADS, 0, c->stream>>>(d_spec, n_hops, chunk_gain, warm, K, d_rel, d_pre, d_gain, d_clean); }
        { JtLaunch L(c, "afftdn:bands", 2);
          k_afftdn_bandsum<<<jt_grid_for(n_hops * 32, 256, c->num_sms, 16), 256, 0, c->stream>>>(d_clean, n_hops, K, d_blo, d_raw);
          k_afftdn_bandrec<<<(int)((n_hops + chunk_band - 1) / chunk_band), 256, smem_band, c->stream>>>(d_raw, n_hops, chunk_band, warm, K, d_alpha, d_beta, d_spread, d_amt); }
        { JtLaunch L(c, "afftdn:synth");
          k_afftdn_synth<<<grid_fft, AF_THREADS, smem_fft, c->stream>>>(d_spec, d_gain, d_amt, n_hops, K, d_rel, d_b2b, d_post, d_tw, d_frames); }
        JtLaunch L(c, "afftdn:ola");
        k_afftdn_ola<<<jt_grid_for(in.n, 256, c->num_sms, 16), 256, 0, c->stream>>>(d_frames, d_window, in.n, n_hops, K.A, K.W, (float *)o.d);
    }
    return o;
}
