// highpass / lowpass (libavfilter/af_biquads.c, RBJ cookbook, poles=2, width_type=q) in the
// link's sample format: "highpass=f=80:poles=2:width_type=q:width=0.707:normalize=1:a=tdii",
// "lowpass=f=20500:..." (reference: filters.go:740-769) and the 17-band region RMS batch
// "highpass=f=lo:p=2,lowpass=f=hi:p=2,astats" (reference: analyser_bands.go:33,
// analyser_noise_bands.go:15-52).
// A biquad is a linear recurrence whose state decays like r^n (r = pole radius), so the
// stream is cut into segments, one sequential lane each, started 48 time-constants early
// from zero state: the result is the sequential one to below the format's rounding.
#include "jt_internal.h"
#include "jt_device.cuh"
#include "jt_tiles.cuh"
#include <cstdlib>
#include <cstdio>
#include <cstring>

BiquadCoef jt_biquad_design(bool highpass, double freq, double q, int rate, bool normalize)
{
    const double w0 = 2 * M_PI * freq / rate;
    // af_biquads.c config_filter(): a corner at or past Nyquist (or a non-positive width) puts the filter in bypass
    // ("Invalid frequency and/or width!") -- the top afftdn band (analyser_noise_bands.go:15-17) at 44.1 / 48 kHz
    if (w0 > M_PI || w0 <= 0.0 || q <= 0.0) return BiquadCoef{1.0, 0.0, 0.0, 0.0, 0.0};
    const double alpha = sin(w0) / (2 * q);
    double a0 = 1 + alpha, a1 = -2 * cos(w0), a2 = 1 - alpha, b0, b1, b2;
    if (highpass) { b0 = (1 + cos(w0)) / 2; b1 = -(1 + cos(w0)); b2 = (1 + cos(w0)) / 2; }
    else { b0 = (1 - cos(w0)) / 2; b1 = 1 - cos(w0); b2 = (1 - cos(w0)) / 2; }
    a1 /= a0; a2 /= a0; b0 /= a0; b1 /= a0; b2 /= a0; a0 = 1;
    if (normalize && fabs(b0 + b1 + b2) > 1e-6) {
        const double factor = (a0 + a1 + a2) / (b0 + b1 + b2);
        b0 *= factor; b1 *= factor; b2 *= factor;
    }
    return BiquadCoef{b0, b1, b2, a1, a2};
}

static int biquad_warmup(const BiquadCoef &k)
{
    // pole radius: complex pair -> sqrt(a2); real poles -> larger root magnitude
    double r;
    const double disc = k.a1 * k.a1 - 4 * k.a2;
    if (disc < 0) r = sqrt(fabs(k.a2));
    else r = std::max(fabs((-k.a1 + sqrt(disc)) / 2), fabs((-k.a1 - sqrt(disc)) / 2));
    if (!(r < 1.0)) return 1 << 20;
    const double tau = -1.0 / log(std::max(r, 1e-9));
    double w = 48.0 * tau + 64;
    if (w < 1024) w = 1024;
    if (w > (1 << 20)) w = 1 << 20;
    return (int)w;
}

template <class F> __device__ __forceinline__ F mul_rn(F a, F b);
template <> __device__ __forceinline__ float mul_rn<float>(float a, float b) { return __fmul_rn(a, b); }
template <> __device__ __forceinline__ double mul_rn<double>(double a, double b) { return __dmul_rn(a, b); }
template <class F> __device__ __forceinline__ F add_rn(F a, F b);
template <> __device__ __forceinline__ float add_rn<float>(float a, float b) { return __fadd_rn(a, b); }
template <> __device__ __forceinline__ double add_rn<double>(double a, double b) { return __dadd_rn(a, b); }

template <class T, class F> struct BqIO;
template <> struct BqIO<float, float> { static __device__ __forceinline__ float out(float v) { return v; } };
template <> struct BqIO<double, double> { static __device__ __forceinline__ double out(double v) { return v; } };
template <> struct BqIO<int16_t, float> {       // need_clipping path: clamp, then C float->int16 conversion (truncation)
    static __device__ __forceinline__ int16_t out(float v) { if (v < -32768.f) return -32768; if (v > 32767.f) return 32767; return (int16_t)(int)v; }
};

template <> struct BqIO<int32_t, double> {      // BIQUAD_FILTER(s32, int32_t, double, INT32_MIN, INT32_MAX, 1): clamp, then truncation
    static __device__ __forceinline__ int32_t out(double v) { if (v < -2147483648.0) return (int32_t)0x80000000; if (v > 2147483647.0) return 2147483647; return (int32_t)v; }
};

template <class F> struct BqState { F s0 = 0, s1 = 0, s2 = 0, s3 = 0; };

// one step; TDII: state (w1,w2).  DI: state (i1,i2,o1,o2).  Unfused mul/add = the C code
template <class F, bool TDII>
__device__ __forceinline__ F bq_step(BqState<F> &s, F in, F b0, F b1, F b2, F na1, F na2, F wet, F dry)
{
    F out;
    if (TDII) {
        out = add_rn(mul_rn(b0, in), s.s0);
        s.s0 = add_rn(add_rn(mul_rn(b1, in), s.s1), mul_rn(na1, out));
        s.s1 = add_rn(mul_rn(b2, in), mul_rn(na2, out));
    } else {
        out = add_rn(add_rn(add_rn(add_rn(mul_rn(s.s1, b2), mul_rn(s.s0, b1)), mul_rn(in, b0)), mul_rn(s.s3, na2)), mul_rn(s.s2, na1));
        s.s1 = s.s0; s.s0 = in; s.s3 = s.s2; s.s2 = out;
    }
    return add_rn(mul_rn(out, wet), mul_rn(in, dry));
}

template <class T, class F, bool TDII>
__global__ void __launch_bounds__(64)
k_biquad(const T *__restrict__ x, T *__restrict__ y, int64_t n, int seg, int warm,
         F b0, F b1, F b2, F na1, F na2, F wet, F dry)
{
    const int64_t lane = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t s0 = lane * seg; if (s0 >= n) return;
    const int64_t s1 = min(s0 + (int64_t)seg, n);
    BqState<F> st;
    int64_t i = max((int64_t)0, s0 - warm);
    for (; i < s0; i++) (void)bq_step<F, TDII>(st, (F)x[i], b0, b1, b2, na1, na2, wet, dry);
#pragma unroll 4
    for (; i < s1; i++) y[i] = BqIO<T, F>::out(bq_step<F, TDII>(st, (F)x[i], b0, b1, b2, na1, na2, wet, dry));
}

// The same lanes on tensor-map tiles (jt_tiles.cuh): one UTMALDG brings the next 64 samples of all 32 lanes of a warp, one
// UTMASTG takes 64 results of each away.  The plain kernel above reads and writes 4 bytes per lane per step at 32 far-apart
// addresses (ncu r2j: 5.8 % issue-active, 18 stall cycles per issue on global loads, 2.6 ms for an hour at 48 kHz) although the
// recurrence itself is three dependent f32 operations per sample.  Same segments, same unfused operations per sample; the warm-up
// is rounded up to whole tiles.
#define BQT_CH 2
#define BQT_NS 4
typedef LaneTileIn<float, BQT_CH, BQT_NS> BqTIn;
typedef LaneTileOut<float, BQT_CH> BqTOut;
#define BQT_SMEM (BqTIn::WARP_BYTES + BqTOut::WARP_BYTES + 128 + 1024)

template <bool TDII, bool EMIT>
__device__ __forceinline__ void bqt_tile(const unsigned char *t, unsigned char *ot, int lane, BqState<float> &st,
                                         float b0, float b1, float b2, float na1, float na2, float wet, float dry)
{
#pragma unroll
    for (int line = 0; line < BQT_CH; line++) {
        float4 v[8];
#pragma unroll
        for (int ch = 0; ch < 8; ch++) v[ch] = *(const float4 *)(t + jt_tile_off(line, lane, ch));
#pragma unroll
        for (int ch = 0; ch < 8; ch++) {
            v[ch].x = bq_step<float, TDII>(st, v[ch].x, b0, b1, b2, na1, na2, wet, dry);
            v[ch].y = bq_step<float, TDII>(st, v[ch].y, b0, b1, b2, na1, na2, wet, dry);
            v[ch].z = bq_step<float, TDII>(st, v[ch].z, b0, b1, b2, na1, na2, wet, dry);
            v[ch].w = bq_step<float, TDII>(st, v[ch].w, b0, b1, b2, na1, na2, wet, dry);
        }
        if (EMIT) {
#pragma unroll
            for (int ch = 0; ch < 8; ch++) *(float4 *)(ot + jt_tile_off(line, lane, ch)) = v[ch];
        }
    }
}

template <bool TDII>
__global__ void __launch_bounds__(32)
k_biquad_tiles(const __grid_constant__ CUtensorMap in_map, const __grid_constant__ CUtensorMap out_map, const float *__restrict__ x,
               float *__restrict__ y, int64_t n, int64_t seg, int64_t warm, int64_t rows_full,
               float b0, float b1, float b2, float na1, float na2, float wet, float dry)
{
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = (unsigned char *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int lane = threadIdx.x & 31;
    const int64_t row0 = (int64_t)blockIdx.x * 32;
    BqTIn in; BqTOut out;
    in.init(smem, (uint64_t *)(smem + BqTIn::WARP_BYTES + BqTOut::WARP_BYTES), &in_map, x, n, seg, warm, row0, rows_full);
    out.init(smem + BqTIn::WARP_BYTES, &out_map, y, n, seg, row0, rows_full);
    BqState<float> st;
    in.prime();
    int tile = 0;
    for (; tile < in.own_tile0; tile++) {
        in.prefetch();
        const unsigned char *t = in.wait(tile);
        bqt_tile<TDII, false>(t, nullptr, lane, st, b0, b1, b2, na1, na2, wet, dry);
        in.release();
    }
    for (; tile < in.ntiles; tile++) {
        in.prefetch();
        const unsigned char *t = in.wait(tile);
        bqt_tile<TDII, true>(t, out.tile(), lane, st, b0, b1, b2, na1, na2, wet, dry);
        in.release();
        out.commit();
    }
    out.finish();
}

static bool launch_biquad_tiles(jt_ctx *c, const Sig &in, Sig &o, const BiquadCoef &k, bool tdii, double mix)
{
    static const char *off = getenv("JT_NO_TILES");
    if ((off && *off == '1') || in.fmt != JT_FMT_FLT) return false;
    const int R = BqTIn::R;
    const int64_t seg = 8192;
    int64_t warm = biquad_warmup(k);
    warm = (warm + R - 1) / R * R;
    if (in.n < 2 * seg) return false;
    CUtensorMap mi, mo;
    if (!jt_lane_tensor_map(&mi, in.d, 4, in.n, seg, BQT_CH) || !jt_lane_tensor_map(&mo, o.d, 4, in.n, seg, BQT_CH)) return false;
    const int64_t lanes = (in.n + seg - 1) / seg;
    const float wet = (float)mix, dry = (float)(1. - wet);
    JtLaunch L(c, "biquad");
    const int grid = (int)((lanes + 31) / 32);
    if (tdii) {
        jt_smem_optin((const void *)k_biquad_tiles<true>, BQT_SMEM);
        k_biquad_tiles<true><<<grid, 32, BQT_SMEM, c->stream>>>(mi, mo, (const float *)in.d, (float *)o.d, in.n, seg, warm, in.n / seg,
                                                               (float)k.b0, (float)k.b1, (float)k.b2, (float)-k.a1, (float)-k.a2, wet, dry);
    } else {
        jt_smem_optin((const void *)k_biquad_tiles<false>, BQT_SMEM);
        k_biquad_tiles<false><<<grid, 32, BQT_SMEM, c->stream>>>(mi, mo, (const float *)in.d, (float *)o.d, in.n, seg, warm, in.n / seg,
                                                                (float)k.b0, (float)k.b1, (float)k.b2, (float)-k.a1, (float)-k.a2, wet, dry);
    }
    return true;
}

template <class T, class F>
static void launch_biquad(jt_ctx *c, const Sig &in, Sig &o, const BiquadCoef &k, bool tdii, double mix)
{
    const int seg = 8192, warm = biquad_warmup(k);
    const int64_t lanes = (in.n + seg - 1) / seg;
    const int grid = (int)((lanes + 63) / 64);
    const F wet = (F)mix, dry = (F)(1. - wet);
    JtLaunch L(c, "biquad");
    if (tdii) k_biquad<T, F, true><<<grid, 64, 0, c->stream>>>((const T *)in.d, (T *)o.d, in.n, seg, warm, (F)k.b0, (F)k.b1, (F)k.b2, (F)-k.a1, (F)-k.a2, wet, dry);
    else k_biquad<T, F, false><<<grid, 64, 0, c->stream>>>((const T *)in.d, (T *)o.d, in.n, seg, warm, (F)k.b0, (F)k.b1, (F)k.b2, (F)-k.a1, (F)-k.a2, wet, dry);
}

Sig jt_biquad(jt_ctx *c, const Sig &in, const BiquadCoef &k, bool tdii, double mix)
{
    Sig o = in; o.d = jt_dalloc_bytes(c, (size_t)std::max<int64_t>(in.n, 1) * jt_fmt_bytes(in.fmt));
    if (in.n <= 0) return o;
    if (launch_biquad_tiles(c, in, o, k, tdii, mix)) return o;
    if (in.fmt == JT_FMT_FLT) launch_biquad<float, float>(c, in, o, k, tdii, mix);
    else if (in.fmt == JT_FMT_DBL) launch_biquad<double, double>(c, in, o, k, tdii, mix);
    else if (in.fmt == JT_FMT_S16) launch_biquad<int16_t, float>(c, in, o, k, tdii, mix);
    else if (in.fmt == JT_FMT_S32) launch_biquad<int32_t, double>(c, in, o, k, tdii, mix);
    else JT_THROW(JT_ERR_UNSUPPORTED, "biquad on sample format %d", in.fmt);
    return o;
}

// ---------------------------------------------------------------------------------------
// K20: all bands x all segments in one launch.  lane = (band, segment); highpass(lo) then
// lowpass(hi), both direct form I in the link format, then astats' sum of squares.
// ---------------------------------------------------------------------------------------
template <class T> struct BandNorm { static __device__ __forceinline__ double nd(T v) { return (double)v; } };      // astats' normalised sample
template <> struct BandNorm<int16_t> { static __device__ __forceinline__ double nd(int16_t v) { return (double)v / 32767.0; } };
template <> struct BandNorm<int32_t> { static __device__ __forceinline__ double nd(int32_t v) { return (double)v / 2147483647.0; } };
struct BandCoef { float hb0, hb1, hb2, hna1, hna2, lb0, lb1, lb2, lna1, lna2; double dh[5], dl[5]; };

template <class T, class F>
__global__ void __launch_bounds__(64)
k_band_rms(const T *__restrict__ x, int64_t n, int seg, int warm, int64_t segs_per_band, int n_bands,
           const BandCoef *__restrict__ coef, double *__restrict__ sumsq /* n_bands */, int64_t acc_from)
{
    const int64_t lane = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double acc = 0;
    const int band = (int)(lane / segs_per_band);
    if (band < n_bands) {
        const int64_t s0 = (lane % segs_per_band) * seg, s1 = min(s0 + (int64_t)seg, n);
        const BandCoef k = coef[band];
        F hb0, hb1, hb2, hna1, hna2, lb0, lb1, lb2, lna1, lna2;
        if (sizeof(F) == 4) { hb0 = k.hb0; hb1 = k.hb1; hb2 = k.hb2; hna1 = k.hna1; hna2 = k.hna2; lb0 = k.lb0; lb1 = k.lb1; lb2 = k.lb2; lna1 = k.lna1; lna2 = k.lna2; }
        else { hb0 = (F)k.dh[0]; hb1 = (F)k.dh[1]; hb2 = (F)k.dh[2]; hna1 = (F)k.dh[3]; hna2 = (F)k.dh[4]; lb0 = (F)k.dl[0]; lb1 = (F)k.dl[1]; lb2 = (F)k.dl[2]; lna1 = (F)k.dl[3]; lna2 = (F)k.dl[4]; }
        BqState<F> sh, sl;
        for (int64_t i = max((int64_t)0, s0 - warm); i < s1; i++) {
            const T h = BqIO<T, F>::out(bq_step<F, false>(sh, (F)x[i], hb0, hb1, hb2, hna1, hna2, (F)1, (F)0));
            const T l = BqIO<T, F>::out(bq_step<F, false>(sl, (F)h, lb0, lb1, lb2, lna1, lna2, (F)1, (F)0));
            if (i >= s0 && i >= acc_from) {
                const double nd = BandNorm<T>::nd(l);
                acc = fma(nd, nd, acc);
            }
        }
    }
    // lanes of one warp may straddle two bands only at band boundaries; add per lane
    if (band < n_bands && acc != 0) atomicAdd(&sumsq[band], acc);
}

// sums of squares of the band-filtered signal over samples [acc_from, in.n) of `in`; the samples before acc_from only warm the
// filters up (a rank of a sharded stream measuring its part of a region: jt_process_audio_sharded).  acc_from = 0 with `in`
// starting at the region's first sample is the whole measurement.
void jt_band_sumsq(jt_ctx *c, const Sig &in, int64_t acc_from, const double *lo, const double *hi, int n_bands, double *sumsq_host)
{
    for (int b = 0; b < n_bands; b++) sumsq_host[b] = 0;
    if (in.n <= 0 || acc_from >= in.n) return;
    std::vector<BandCoef> hc(n_bands);
    int warm = 1024;
    for (int b = 0; b < n_bands; b++) {
        BiquadCoef h = jt_biquad_design(true, lo[b], 0.707, in.rate, false), l = jt_biquad_design(false, hi[b], 0.707, in.rate, false);
        BandCoef &k = hc[b];
        k.hb0 = (float)h.b0; k.hb1 = (float)h.b1; k.hb2 = (float)h.b2; k.hna1 = (float)-h.a1; k.hna2 = (float)-h.a2;
        k.lb0 = (float)l.b0; k.lb1 = (float)l.b1; k.lb2 = (float)l.b2; k.lna1 = (float)-l.a1; k.lna2 = (float)-l.a2;
        k.dh[0] = h.b0; k.dh[1] = h.b1; k.dh[2] = h.b2; k.dh[3] = -h.a1; k.dh[4] = -h.a2;
        k.dl[0] = l.b0; k.dl[1] = l.b1; k.dl[2] = l.b2; k.dl[3] = -l.a1; k.dl[4] = -l.a2;
        warm = std::max(warm, std::min(std::max(biquad_warmup(h), biquad_warmup(l)), JT_BAND_WARM_MAX));
    }
    BandCoef *d_coef = jt_dalloc<BandCoef>(c, n_bands);
    double *d_sum = jt_dalloc<double>(c, n_bands);
    BandCoef *h_coef = jt_pinned<BandCoef>(c, (size_t)n_bands);
    memcpy(h_coef, hc.data(), sizeof(BandCoef) * n_bands);
    jt_copy_small(c, d_coef, h_coef, sizeof(BandCoef) * n_bands);
    JT_CUDA(cudaMemsetAsync(d_sum, 0, sizeof(double) * n_bands, c->stream));
    const int seg = 4096;
    const int64_t spb = (in.n + seg - 1) / seg, lanes = spb * n_bands;
    const int grid = (int)((lanes + 63) / 64);
    {
        JtLaunch L(c, "band_rms");
        if (in.fmt == JT_FMT_FLT) k_band_rms<float, float><<<grid, 64, 0, c->stream>>>((const float *)in.d, in.n, seg, warm, spb, n_bands, d_coef, d_sum, acc_from);
        else if (in.fmt == JT_FMT_DBL) k_band_rms<double, double><<<grid, 64, 0, c->stream>>>((const double *)in.d, in.n, seg, warm, spb, n_bands, d_coef, d_sum, acc_from);
        else if (in.fmt == JT_FMT_S16) k_band_rms<int16_t, float><<<grid, 64, 0, c->stream>>>((const int16_t *)in.d, in.n, seg, warm, spb, n_bands, d_coef, d_sum, acc_from);
        else if (in.fmt == JT_FMT_S32) k_band_rms<int32_t, double><<<grid, 64, 0, c->stream>>>((const int32_t *)in.d, in.n, seg, warm, spb, n_bands, d_coef, d_sum, acc_from);
        else JT_THROW(JT_ERR_UNSUPPORTED, "band rms on sample format %d", in.fmt);
    }
    double *h_sum = jt_pinned<double>(c, (size_t)n_bands);
    jt_copy_small(c, h_sum, d_sum, sizeof(double) * n_bands);
    JT_CUDA(cudaStreamSynchronize(c->stream));
    memcpy(sumsq_host, h_sum, sizeof(double) * n_bands);
}

void jt_band_rms_batch(jt_ctx *c, const Sig &in, const double *lo, const double *hi, int n_bands, double *rms_db, int32_t *found)
{
    for (int b = 0; b < n_bands; b++) { rms_db[b] = 0; found[b] = 0; }
    if (in.n <= 0) return;
    std::vector<double> hs(n_bands);
    jt_band_sumsq(c, in, 0, lo, hi, n_bands, hs.data());
    for (int b = 0; b < n_bands; b++) { rms_db[b] = jt_wire("%f", log10(sqrt(hs[b] / (double)in.n)) * 20); found[b] = 1; }
}
