// afftdn -- FFT spectral-subtraction denoiser (libavfilter/af_afftdn.c), fltp path:
// "afftdn=nr=12:nt=w|custom:bn=<15 dB>:tn=0|1[:nf=<dB>]" (reference: filters.go:830-861,
// adaptive.go:133-170; 15 band centres analyser_noise_bands.go:15-17).
// window = 3 * (fs/80) samples of wscale*sin^2, hop fs/80, 2048-point real transform (at 44.1/48 kHz),
// per-bin a-priori-SNR gain with a Bark-band masking floor; output delayed by window-hop.
//
// The filter's only cross-hop state is (a) prior[bin], a contraction (factor <= 0.39 per hop),
// (b) prior_band_excit[band], a contraction (factor beta <= 0.78), (c) the tracked noise floor
// (tn=1), which depends on the INPUT spectra only.  So:
//   F1  one CTA per hop: window, forward FFT (shared-memory radix-2), spectrum -> HBM;
//       with tn=1 also the hop's flatness / floor candidate
//   F2  one thread: noise-floor recurrence over hops (tn=1)
//   F3  one CTA per chunk of hops (+128 warm-up hops): the gain recursion, band masking,
//       gain limiting, inverse FFT -> windowed frames in HBM
//   F4  overlap-add of the three frames covering each output sample (f64, hop order)
#include "jt_internal.h"
#include "jt_device.cuh"
#include <cstdio>

#define AF_THREADS 512
#define AF_MAXBANDS 64
#define AF_MAXOWN 5
#define AF_C (M_LN10 * 0.1)

struct AfConst {
    int A, W, FL, FL2, bins, nbands;
    double floor_, gain_scale, max_gain, ratio, floor_offset;
};

__device__ __forceinline__ void af_fft(float2 *s, const float2 *__restrict__ tw, int n, bool inverse)
{   // in-place radix-2 DIT on bit-reversed input (caller stores bit-reversed)
    for (int len = 2; len <= n; len <<= 1) {
        const int half = len >> 1, tstep = n / len;
        for (int b = threadIdx.x; b < n / 2; b += blockDim.x) {
            const int k = b & (half - 1), i = ((b - k) << 1) + k;
            float2 w = tw[k * tstep]; if (inverse) w.y = -w.y;
            const float2 a = s[i], bb = s[i + half];
            const float tr = __fsub_rn(__fmul_rn(bb.x, w.x), __fmul_rn(bb.y, w.y));
            const float ti = __fadd_rn(__fmul_rn(bb.x, w.y), __fmul_rn(bb.y, w.x));
            s[i] = make_float2(a.x + tr, a.y + ti);
            s[i + half] = make_float2(a.x - tr, a.y - ti);
        }
        __syncthreads();
    }
}

__device__ __forceinline__ double af_block_sum(double v, double *red)
{
    v = jt_warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0; for (int i = 0; i < AF_THREADS / 32; i++) s += red[i];
    return s;
}
__device__ __forceinline__ double af_block_max(double v, double *red)
{
    v = jt_warp_max(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0; for (int i = 0; i < AF_THREADS / 32; i++) s = fmax(s, red[i]);
    return s;
}

// F1: forward transform of every hop; optional noise-floor candidate (track_noise)
__global__ void __launch_bounds__(AF_THREADS)
k_afftdn_fwd(const float *__restrict__ x, int64_t n, int64_t n_hops, AfConst K, const double *__restrict__ window,
             const float2 *__restrict__ tw, float2 *__restrict__ spectra, int track, double *__restrict__ cand /* 2 per hop: new_floor, flag */)
{
    extern __shared__ float2 sb[];
    __shared__ double red[AF_THREADS / 32];
    int logn = 0; while ((1 << logn) < K.FL) logn++;
    for (int64_t h = blockIdx.x; h < n_hops; h += gridDim.x) {
        const int64_t w0 = (h - 2) * (int64_t)K.A, avail = min(n, (h + 1) * (int64_t)K.A);
        __syncthreads();
        for (int m = threadIdx.x; m < K.FL; m += AF_THREADS) {
            float v = 0.f;
            if (m < K.W) { const int64_t s = w0 + m; if (s >= 0 && s < avail) v = (float)(window[m] * (double)x[s] * 8388608.0); }
            sb[__brev((unsigned)m) >> (32 - logn)] = make_float2(v, 0.f);
        }
        __syncthreads();
        af_fft(sb, tw, K.FL, false);
        for (int i = threadIdx.x; i < K.bins; i += AF_THREADS) spectra[h * (int64_t)K.bins + i] = sb[i];
        if (track) {
            double num = 0, den = 0, cnt = 0;
            for (int i = threadIdx.x; i < K.bins; i += AF_THREADS) {
                const double v = hypot((double)sb[i].x, (double)sb[i].y);
                if (v > K.floor_) { num += log(v); den += v; cnt += 1; }
            }
            num = af_block_sum(num, red); den = af_block_sum(den, red); cnt = af_block_sum(cnt, red);
            const double size = fmax(cnt, 1.0);
            num = exp(num / size); den /= size;
            double off = 0;
            for (int i = threadIdx.x; i < K.bins; i += AF_THREADS) off = fmax(off, fabs(hypot((double)sb[i].x, (double)sb[i].y) - den));
            off = af_block_max(off, red);
            if (threadIdx.x == 0) {
                const double flat = num / den;
                double nf = 10.0 * log10(den) - 100.0 + K.floor_offset * (off / den);
                nf = fmin(fmax(nf, -90.), -20.);
                cand[2 * h] = nf; cand[2 * h + 1] = (flat > 0.8) ? 1.0 : 0.0;
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

// F3: the recursion over a chunk of hops
__global__ void __launch_bounds__(AF_THREADS)
k_afftdn_core(const float2 *__restrict__ spectra, int64_t n_hops, int chunk, int warm, AfConst K,
              const double *__restrict__ rel_var, const int *__restrict__ band_lo /* nbands+1 */,
              const int *__restrict__ bin2band, const double *__restrict__ band_alpha, const double *__restrict__ band_beta,
              const double *__restrict__ spread, const double *__restrict__ mv_pre, const double *__restrict__ mv_post,
              const float2 *__restrict__ tw, float *__restrict__ frames)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2 *sb = (float2 *)smem_raw;                               // FL complex (inverse transform)
    double *clean = (double *)(smem_raw + sizeof(float2) * K.FL);  // bins
    __shared__ double s_be[AF_MAXBANDS], s_amt[AF_MAXBANDS];
    const int t = threadIdx.x;
    const int64_t h_out0 = (int64_t)blockIdx.x * chunk;
    if (h_out0 >= n_hops) return;
    const int64_t h_out1 = min(h_out0 + chunk, n_hops);
    const int64_t h_begin = max((int64_t)0, h_out0 - warm);
    int logn = 0; while ((1 << logn) < K.FL) logn++;
    // per-thread bins: t, t + 512, ... (bins <= AF_MAXOWN * 512)
    constexpr int nb_own = AF_MAXOWN;
    int bidx[AF_MAXOWN];
    double prior[AF_MAXOWN], rv[AF_MAXOWN], gain[AF_MAXOWN];
    float2 sp[AF_MAXOWN];
#pragma unroll
    for (int k = 0; k < nb_own; k++) { bidx[k] = t + k * AF_THREADS; prior[k] = 0.0; gain[k] = 0.0; sp[k] = make_float2(0.f, 0.f); rv[k] = bidx[k] < K.bins ? rel_var[bidx[k]] : 1.0; }
    double prior_be = 0.0;                                          // thread b < nbands owns band b
    const double alpha = t < K.nbands ? band_alpha[t] : 0, beta = t < K.nbands ? band_beta[t] : 0;
    const int blo = t < K.nbands ? band_lo[t] : 0, bhi = t < K.nbands ? band_lo[t + 1] : 0;

    for (int64_t h = h_begin; h < h_out1; h++) {
        const double ratio = (h == 0) ? 1.0 : K.ratio, rratio = 1.0 - ratio;
        const double mvp = mv_pre[h], mvq = mv_post[h];
#pragma unroll
        for (int k = 0; k < nb_own; k++) {
            const int i = bidx[k]; if (i >= K.bins) continue;
            sp[k] = spectra[h * (int64_t)K.bins + i];
            const double mag = hypot((double)sp[k].x, (double)sp[k].y), power = mag * mag;
            const double abs_var = fmax(mvp * rv[k], 1.0);
            const double mav = power / abs_var;
            const double nmav = ratio * prior[k] + rratio * fmax(mav - 1.0, 0.0);
            const double g = nmav / (1.0 + nmav), sg = g * g;
            prior[k] = mav * sg;
            clean[i] = power * sg;
            gain[k] = g;
        }
        __syncthreads();
        if (t < K.nbands) {
            double be = 0.0;
            for (int i = blo; i < bhi; i++) be += clean[i];
            be = fmax(be, alpha * be + beta * prior_be);
            prior_be = be;
            s_be[t] = be;
        }
        __syncthreads();
        if (h >= h_out0) {
            if (t < K.nbands) {
                double a = 0.0;
                for (int k = 0; k < K.nbands; k++) a += spread[t * K.nbands + k] * s_be[k];
                s_amt[t] = a;
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < nb_own; k++) {
                const int i = bidx[k]; if (i >= K.bins) continue;
                const double amt = s_amt[bin2band[i]];
                const double abs_var = fmax(mvq * rv[k], 1.0), min_abs_var = K.gain_scale * abs_var;
                double g = gain[k];
                if (amt > abs_var) g = 1.0;
                else if (amt > min_abs_var) g = af_limit_gain(g, sqrt(abs_var / amt));
                else g = af_limit_gain(g, K.max_gain);
                const float ng = (float)g;
                float2 v = make_float2(__fmul_rn(sp[k].x, ng), __fmul_rn(sp[k].y, ng));
                if (i == 0 || i == K.FL2) v.y = 0.f;
                sb[__brev((unsigned)i) >> (32 - logn)] = v;
                if (i > 0 && i < K.FL2) sb[__brev((unsigned)(K.FL - i)) >> (32 - logn)] = make_float2(v.x, -v.y);
            }
            __syncthreads();
            af_fft(sb, tw, K.FL, true);
            for (int m = t; m < K.W; m += AF_THREADS) frames[h * (int64_t)K.W + m] = sb[m].x;
        }
        __syncthreads();
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

Sig jt_afftdn(jt_ctx *c, const Sig &in, const AfftdnParams &P)
{
    if (in.fmt != JT_FMT_FLT) JT_THROW(JT_ERR_INVALID_ARG, "afftdn expects float input");
    Sig o = in; o.d = jt_dalloc<float>(c, in.n);
    if (in.n <= 0) return o;
    AfConst K;
    const double sample_rate = (float)in.rate;
    K.A = (int)(sample_rate / 80); K.W = 3 * K.A;
    K.FL = 1; while (K.FL <= K.W) K.FL <<= 1;
    K.FL2 = K.FL / 2; K.bins = K.FL2 + 1;
    if (K.FL < 512 || K.bins > AF_MAXOWN * AF_THREADS) JT_THROW(JT_ERR_UNSUPPORTED, "afftdn at %d Hz (transform length %d)", in.rate, K.FL);
    K.ratio = P.ad; K.floor_offset = P.fo;
    std::vector<double> window(K.W); double sum = 0;
    { const double wscale = sqrt(8.0 / (9.0 * K.FL)); for (int i = 0; i < K.W; i++) { double d = sin(i * M_PI / K.W); d *= wscale * d; window[i] = d; sum += d * d; } }
    K.floor_ = (double)(1LL << 48) * exp(-23.025558369790467) * (0.5 * sum);
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
    std::vector<float2> tw(K.FL / 2);
    for (int k = 0; k < K.FL / 2; k++) { const double a = -2.0 * M_PI * k / K.FL; tw[k] = make_float2((float)cos(a), (float)sin(a)); }

    const int64_t n_hops = (in.n + K.A - 1) / K.A;
    auto up = [&](const void *h, size_t bytes) { void *d = jt_dalloc_bytes(c, bytes); JT_CUDA(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, c->stream)); return d; };
    double *d_window = (double *)up(window.data(), sizeof(double) * K.W);
    float2 *d_tw = (float2 *)up(tw.data(), sizeof(float2) * tw.size());
    double *d_rel = (double *)up(rel_var.data(), sizeof(double) * K.bins);
    int *d_blo = (int *)up(band_lo.data(), sizeof(int) * (nb + 1));
    int *d_b2b = (int *)up(bin2band.data(), sizeof(int) * K.bins);
    double *d_alpha = (double *)up(alpha.data(), sizeof(double) * nb), *d_beta = (double *)up(beta.data(), sizeof(double) * nb);
    double *d_spread = (double *)up(spread.data(), sizeof(double) * nb * nb);
    JT_CUDA(cudaStreamSynchronize(c->stream));          // host vectors above are locals
    float2 *d_spec = jt_dalloc<float2>(c, (size_t)n_hops * K.bins);
    float *d_frames = jt_dalloc<float>(c, (size_t)n_hops * K.W);
    double *d_cand = jt_dalloc<double>(c, (size_t)n_hops * 2), *d_pre = jt_dalloc<double>(c, n_hops), *d_post = jt_dalloc<double>(c, n_hops);
    const size_t smem1 = sizeof(float2) * K.FL, smem3 = sizeof(float2) * K.FL + sizeof(double) * (K.bins + 1);
    JT_CUDA(cudaFuncSetAttribute(k_afftdn_core, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem3));
    const int chunk = 384, warm = 128;
    {
        JtLaunch L(c, "afftdn", 4);
        k_afftdn_fwd<<<jt_grid_for(n_hops, 1, c->num_sms, 16), AF_THREADS, smem1, c->stream>>>((const float *)in.d, in.n, n_hops, K, d_window, d_tw, d_spec, P.tn, d_cand);
        k_afftdn_floor<<<1, 1024, 0, c->stream>>>(d_cand, n_hops, P.nf, K.floor_, P.tn, d_pre, d_post);
        k_afftdn_core<<<(int)((n_hops + chunk - 1) / chunk), AF_THREADS, smem3, c->stream>>>(d_spec, n_hops, chunk, warm, K, d_rel, d_blo, d_b2b, d_alpha, d_beta, d_spread, d_pre, d_post, d_tw, d_frames);
        k_afftdn_ola<<<jt_grid_for(in.n, 256, c->num_sms, 16), 256, 0, c->stream>>>(d_frames, d_window, in.n, n_hops, K.A, K.W, (float *)o.d);
    }
    return o;
}
