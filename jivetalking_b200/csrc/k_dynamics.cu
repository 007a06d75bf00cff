// agate, acompressor, deesser (libavfilter/af_agate.c, af_sidechaincompress.c, af_deesser.c), f64:
//   "agate=threshold=..:ratio=..:attack=5.00:release=200:range=..:knee=3.0:detection=rms:makeup=1.0"
//   "acompressor=threshold=..:ratio=3.0:attack=10:release=200:makeup=1.00:knee=4.0:detection=rms:mix=1.00"
//   "deesser=i=..:m=0.50:f=0.80"               (reference: filters.go:869-932)
// Gate and compressor are a switching one-pole envelope follower (sequential, ~2 FLOP per
// sample) followed by a memoryless log-domain gain law (parallel, ~2 transcendental calls per
// sample).  The follower env' = env + (d-env)*(d>env ? a : r) is a contraction with factor
// (1 - min(a,r)) per sample, so the stream is cut into segments, one lane each, started
// 37/min(a,r) samples early from zero: after the warm-up the lane's state equals the
// sequential state to the last bit of f64.  The gain law then runs fully parallel.
// The de-esser is a per-sample nonlinear recurrence (Airwindows DeEss) and runs as lanes too.
#include "jt_internal.h"
#include "jt_device.cuh"
#include "jt_lanes.cuh"
#include "jt_tiles.cuh"
#include <cstdlib>

#define FAKE_INFINITY (65536.0 * 65536.0)

// ---- switching envelope follower ----------------------------------------------------------
#define ENV_R 128           // input rows of 1 KB, double-buffered (a 4-deep ring of 512 B rows measured 20 % slower)
#define ENV_RO 32           // output rows of 256 B
#define ENV_THREADS 32
typedef LaneStage<double, ENV_R, 2> EnvIn;
typedef LaneStore<double, ENV_RO> EnvOut;
// env' = env + (d - env) * (d > env ? attack : release), rounded like the scalar C (no contraction).
// Both branches are evaluated and the comparison selects, so the carried chain is
// sub -> mul -> add -> select (~28 cycles at 8.2 cycles per dependent f64 op, profiles/ubench_r1.txt)
// instead of compare -> select -> sub -> mul -> add.
__device__ __forceinline__ double env_step(double e, double d, double attack_coeff, double release_coeff)
{
    const double t = __dsub_rn(d, e);
    const double ea = __dadd_rn(e, __dmul_rn(t, attack_coeff)), er = __dadd_rn(e, __dmul_rn(t, release_coeff));
    return d > e ? ea : er;
}

__global__ void __launch_bounds__(ENV_THREADS)
k_envelope(const double *__restrict__ x, double *__restrict__ env, int64_t n, int seg, int warm,
           double attack_coeff, double release_coeff, int rms)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int warp = threadIdx.x >> 5;
    unsigned char *wsm = smem + (size_t)warp * (EnvIn::WARP_BYTES + EnvOut::WARP_BYTES);
    const int64_t lane = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t s0 = min(lane * seg, n), s1 = min(s0 + (int64_t)seg, n);
    const int64_t begin = max((int64_t)0, s0 - warm);
    EnvIn in; EnvOut out;
    in.init(wsm, x + begin, lane * seg < n ? s1 - begin : 0);
    out.init(wsm + EnvIn::WARP_BYTES, env + s0);
    // warm and seg are multiples of ENV_R, so a tile is either all warm-up or all output: the inner
    // loops are branch-free, which lets ptxas hoist the shared-memory loads and overlap everything
    // except the carried chain
    double e = 0.0;
    in.prime();
    for (int tile = 0; tile < in.ntiles; tile++) {
        in.prefetch();
        const double *row = in.wait(tile);
        const int nv = in.valid(tile);
        const int64_t i0 = begin + (int64_t)tile * ENV_R;
        const bool emit = i0 >= s0;
        for (int k0 = 0; k0 < nv; k0 += ENV_RO) {
            const int nb = min(ENV_RO, nv - k0);
            double *orow = out.row();
            int k = 0;
            // register blocks of 8: all loads and detector values first, then only the carried chain
            for (; k + 8 <= nb; k += 8) {
                double d[8];
#pragma unroll
                for (int j = 0; j < 8; j++) { const double v = row[k0 + k + j]; d[j] = rms ? v * v : fabs(v); }
#pragma unroll
                for (int j = 0; j < 8; j++) { e = env_step(e, d[j], attack_coeff, release_coeff); d[j] = e; }
                if (emit) {
#pragma unroll
                    for (int j = 0; j < 8; j++) orow[k + j] = d[j];
                }
            }
            for (; k < nb; k++) {
                const double v = row[k0 + k], dd = rms ? v * v : fabs(v);
                e = env_step(e, dd, attack_coeff, release_coeff);
                if (emit) orow[k] = e;
            }
            if (emit) out.commit(nb);
        }
        in.release();
    }
    out.finish();
}

// ---- the same follower on tensor-map tiles (jt_tiles.cuh) ------------------------------------------------------------------
// One warp per CTA, 32 samples per lane per tile (two 128-byte lines), a 4-deep ring in, two tiles out: 49 KB per warp, four
// warps per SM -- one per scheduler, which is what the FP64 pipe can feed (5 DP instructions per sample, 2 issue cycles each).
// The update runs in the form  e' = fma(e, 1 - c, c * d)  with both branches evaluated: the products c * d do not depend on
// the state, so the carried chain is ONE fma plus the select, and the branch condition d > e is taken on the bit patterns
// (both are non-negative, so they order like integers) on the integer pipe, in the shadow of the fma.  Against the scalar
// C form  e + (d - e) * c  the result differs in the last place per step; the follower is a contraction, so the differences
// do not accumulate beyond ~1e-16 / min(c) relative (2.6e-13 at 200 ms release), far inside the 1e-12 the gain stages are
// tested to.
#ifndef ENVT_CH
#define ENVT_CH 4
#endif
#ifndef ENVT_NS
#define ENVT_NS 3
#endif
typedef LaneTileIn<double, ENVT_CH, ENVT_NS> EnvTIn;
typedef LaneTileOut<double, ENVT_CH> EnvTOut;
#define ENVT_SMEM (EnvTIn::WARP_BYTES + EnvTOut::WARP_BYTES + 128 + 1024)      // tiles in, tiles out, barriers, alignment slack

template <int RMS, bool EMIT>
__device__ __forceinline__ double envt_tile(const unsigned char *t, unsigned char *ot, int lane, double e, double ca, double cr, double ka, double kr)
{
#pragma unroll
    for (int line = 0; line < ENVT_CH; line++) {
        double2 v[8];
#pragma unroll
        for (int ch = 0; ch < 8; ch++) v[ch] = *(const double2 *)(t + jt_tile_off(line, lane, ch));
#pragma unroll
        for (int ch = 0; ch < 8; ch++) {
            const double d0 = RMS ? v[ch].x * v[ch].x : fabs(v[ch].x), d1 = RMS ? v[ch].y * v[ch].y : fabs(v[ch].y);
            const double ad0 = ca * d0, rd0 = cr * d0, ad1 = ca * d1, rd1 = cr * d1;
            const double ea0 = fma(e, ka, ad0), er0 = fma(e, kr, rd0);
            e = __double_as_longlong(d0) > __double_as_longlong(e) ? ea0 : er0;
            v[ch].x = e;
            const double ea1 = fma(e, ka, ad1), er1 = fma(e, kr, rd1);
            e = __double_as_longlong(d1) > __double_as_longlong(e) ? ea1 : er1;
            v[ch].y = e;
        }
        if (EMIT) {
#pragma unroll
            for (int ch = 0; ch < 8; ch++) *(double2 *)(ot + jt_tile_off(line, lane, ch)) = v[ch];
        }
    }
    return e;
}

template <int RMS>
__global__ void __launch_bounds__(32)
k_envelope_tiles(const __grid_constant__ CUtensorMap in_map, const __grid_constant__ CUtensorMap out_map, const double *__restrict__ x,
                 double *__restrict__ env, int64_t n, int64_t seg, int64_t warm, int64_t rows_full, double attack_coeff, double release_coeff)
{
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = (unsigned char *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int lane = threadIdx.x & 31;
    const int64_t row0 = (int64_t)blockIdx.x * 32;
    EnvTIn in; EnvTOut out;
    in.init(smem, (uint64_t *)(smem + EnvTIn::WARP_BYTES + EnvTOut::WARP_BYTES), &in_map, x, n, seg, warm, row0, rows_full);
    out.init(smem + EnvTIn::WARP_BYTES, &out_map, env, n, seg, row0, rows_full);
    const double ka = 1.0 - attack_coeff, kr = 1.0 - release_coeff;
    double e = 0.0;
    in.prime();
    int tile = 0;
    for (; tile < in.own_tile0; tile++) {                                   // warm-up: nothing is stored
        in.prefetch();
        const unsigned char *t = in.wait(tile);
        e = envt_tile<RMS, false>(t, nullptr, lane, e, attack_coeff, release_coeff, ka, kr);
        in.release();
    }
    for (; tile < in.ntiles; tile++) {
        in.prefetch();
        const unsigned char *t = in.wait(tile);
        e = envt_tile<RMS, true>(t, out.tile(), lane, e, attack_coeff, release_coeff, ka, kr);
        in.release();
        out.commit();
    }
    out.finish();
}

static bool run_envelope_tiles(jt_ctx *c, const Sig &in, double *env, double ac, double rc, int64_t warm, int rms)
{
    static const char *off = getenv("JT_NO_TILES");
    if (off && *off == '1') return false;
    const int R = EnvTIn::R;
    warm = (warm + R - 1) / R * R;                          // tiles are all warm-up or all output
    // segment: long enough that the warm-up re-read (warm / seg) does not make the pass HBM bound, short enough to fill the SMs
    static const char *seg_env = getenv("JT_ENV_SEG");
    const int64_t slots = (int64_t)c->num_sms * (int64_t)((227 * 1024) / (ENVT_SMEM + 1024)) * 32;      // lanes of one wave
    int64_t seg = std::max<int64_t>(seg_env ? atoll(seg_env) : 16384, (in.n + slots - 1) / slots);
    seg = std::min<int64_t>((seg + 127) / 128 * 128, 1 << 22);
    if (in.n < 2 * seg || seg % R) return false;
    CUtensorMap mi, mo;
    if (!jt_lane_tensor_map(&mi, in.d, 8, in.n, seg, ENVT_CH) || !jt_lane_tensor_map(&mo, env, 8, in.n, seg, ENVT_CH)) return false;
    const int64_t lanes = (in.n + seg - 1) / seg;
    JtLaunch L(c, "envelope_follower");
    if (rms) {
        jt_smem_optin((const void *)k_envelope_tiles<1>, ENVT_SMEM);
        k_envelope_tiles<1><<<(int)((lanes + 31) / 32), 32, ENVT_SMEM, c->stream>>>(mi, mo, (const double *)in.d, env, in.n, seg, warm, in.n / seg, ac, rc);
    } else {
        jt_smem_optin((const void *)k_envelope_tiles<0>, ENVT_SMEM);
        k_envelope_tiles<0><<<(int)((lanes + 31) / 32), 32, ENVT_SMEM, c->stream>>>(mi, mo, (const double *)in.d, env, in.n, seg, warm, in.n / seg, ac, rc);
    }
    return true;
}

static double *run_envelope(jt_ctx *c, const Sig &in, double attack_ms, double release_ms, int rms)
{
    const double ac = std::fmin(1., 1. / (attack_ms * in.rate / 4000.)), rc = std::fmin(1., 1. / (release_ms * in.rate / 4000.));
    double *env = jt_dalloc<double>(c, in.n);
    const double cmin = std::fmin(ac, rc);
    int64_t warm = cmin >= 1.0 ? 1 : (int64_t)std::ceil(37.0 / -std::log1p(-cmin)) + 16;
    if (warm > (1 << 22)) warm = 1 << 22;
    if (run_envelope_tiles(c, in, env, ac, rc, warm, rms)) return env;
    warm = (warm + ENV_R - 1) / ENV_R * ENV_R;                 // tile-aligned (see the kernel)
    // Every lane re-reads `warm` samples before its segment (37 release time constants, ~89k samples at
    // 200 ms / 48 kHz), so short segments multiply the HBM traffic while long ones leave the GPU to a handful
    // of lanes.  84 KB of staging per warp lets 2 warps share an SM: the segment is sized so that all lanes
    // run in one wave (20480 samples at one hour of audio: 5.3x re-read), never below 16384.
    const int64_t slots = (int64_t)c->num_sms * 2 * ENV_THREADS;
    int64_t seg64 = std::max<int64_t>(16384, (in.n + slots - 1) / slots);
    seg64 = std::min<int64_t>((seg64 + ENV_R - 1) / ENV_R * ENV_R, 1 << 20);
    const int seg = (int)seg64;
    const int64_t lanes = (in.n + seg - 1) / seg;
    JtLaunch L(c, "envelope_follower");
    const size_t smem = (ENV_THREADS / 32) * (EnvIn::WARP_BYTES + EnvOut::WARP_BYTES);
    jt_smem_optin((const void *)k_envelope, (size_t)(smem));
    k_envelope<<<(int)((lanes + ENV_THREADS - 1) / ENV_THREADS), ENV_THREADS, smem, c->stream>>>((const double *)in.d, env, in.n, seg, (int)warm, ac, rc, rms);
    return env;
}

__device__ __forceinline__ double hermite_interpolation(double x, double x0, double x1, double p0, double p1, double m0, double m1)
{
    const double width = x1 - x0, t = (x - x0) / width;
    m0 *= width; m1 *= width;
    const double t2 = t * t, t3 = t2 * t;
    const double ct0 = p0, ct1 = m0, ct2 = -3 * p0 - 2 * m0 + 3 * p1 - m1, ct3 = 2 * p0 + m0 - 2 * p1 + m1;
    return ct3 * t3 + ct2 * t2 + ct1 * t + ct0;
}

struct GateK { double ratio, thres, knee, knee_start, knee_stop, lin_knee_stop, range, makeup; };

__global__ void __launch_bounds__(256)
k_gate_apply(const double *__restrict__ x, const double *__restrict__ env, double *__restrict__ y, int64_t n, GateK k)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double ls = env[i]; double gain = 1.0;
        if (ls > 0.0 && ls < k.lin_knee_stop) {
            const double slope = log(ls);
            double tratio = k.ratio;
            if (fabs(k.ratio - FAKE_INFINITY) < 1.0) tratio = 1000.;
            double g = (slope - k.thres) * tratio + k.thres;
            if (k.knee > 1. && slope > k.knee_start)
                g = hermite_interpolation(slope, k.knee_start, k.knee_stop, ((k.knee_start - k.thres) * tratio + k.thres), k.knee_stop, tratio, 1.);
            gain = fmax(k.range, exp(g - slope));
        }
        y[i] = x[i] * (1.0 * gain * k.makeup);
    }
}

Sig jt_agate(jt_ctx *c, const Sig &in, const GateParams &p)
{
    if (in.fmt != JT_FMT_DBL) JT_THROW(JT_ERR_INVALID_ARG, "agate expects f64 input");
    Sig o = in; o.d = jt_dalloc<double>(c, in.n);
    if (in.n <= 0) return o;
    double *env = run_envelope(c, in, p.attack, p.release, p.detection_rms);
    GateK k;
    double lin_threshold = p.threshold; const double lin_knee_sqrt = sqrt(p.knee);
    if (p.detection_rms) lin_threshold *= lin_threshold;
    k.lin_knee_stop = lin_threshold * lin_knee_sqrt;
    const double lin_knee_start = lin_threshold / lin_knee_sqrt;
    k.thres = log(lin_threshold); k.knee_start = log(lin_knee_start); k.knee_stop = log(k.lin_knee_stop);
    k.ratio = p.ratio; k.knee = p.knee; k.range = p.range; k.makeup = p.makeup;
    JtLaunch L(c, "agate_gain");
    k_gate_apply<<<jt_grid_for(in.n, 256, c->num_sms, 16), 256, 0, c->stream>>>((const double *)in.d, env, (double *)o.d, in.n, k);
    return o;
}

struct CompK { double ratio, thres, knee, knee_start, knee_stop, compressed_knee_stop, detector, makeup, mix; int rms; };

__global__ void __launch_bounds__(256)
k_comp_apply(const double *__restrict__ x, const double *__restrict__ env, double *__restrict__ y, int64_t n, CompK k)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double ls = env[i]; double gain = 1.0;
        if (ls > 0.0 && ls > k.detector) {
            double slope = log(ls), g, delta;
            if (k.rms) slope *= 0.5;
            if (fabs(k.ratio - FAKE_INFINITY) < 1.0) { g = k.thres; delta = 0.0; }
            else { g = (slope - k.thres) / k.ratio + k.thres; delta = 1.0 / k.ratio; }
            if (k.knee > 1.0 && slope < k.knee_stop)
                g = hermite_interpolation(slope, k.knee_start, k.knee_stop, k.knee_start, k.compressed_knee_stop, 1.0, delta);
            gain = exp(g - slope);
        }
        y[i] = x[i] * 1.0 * (gain * k.makeup * k.mix + (1. - k.mix));
    }
}

Sig jt_acompressor(jt_ctx *c, const Sig &in, const CompParams &p)
{
    if (in.fmt != JT_FMT_DBL) JT_THROW(JT_ERR_INVALID_ARG, "acompressor expects f64 input");
    Sig o = in; o.d = jt_dalloc<double>(c, in.n);
    if (in.n <= 0) return o;
    double *env = run_envelope(c, in, p.attack, p.release, p.detection_rms);
    CompK k;
    k.thres = log(p.threshold);
    const double lin_knee_start = p.threshold / sqrt(p.knee), lin_knee_stop = p.threshold * sqrt(p.knee);
    k.knee_start = log(lin_knee_start); k.knee_stop = log(lin_knee_stop);
    k.compressed_knee_stop = (k.knee_stop - k.thres) / p.ratio + k.thres;
    k.detector = p.detection_rms ? lin_knee_start * lin_knee_start : lin_knee_start;
    k.ratio = p.ratio; k.knee = p.knee; k.makeup = p.makeup; k.mix = p.mix; k.rms = p.detection_rms;
    JtLaunch L(c, "acompressor_gain");
    k_comp_apply<<<jt_grid_for(in.n, 256, c->num_sms, 16), 256, 0, c->stream>>>((const double *)in.d, env, (double *)o.d, in.n, k);
    return o;
}

// ---- de-esser -----------------------------------------------------------------------------
__global__ void __launch_bounds__(64)
k_deesser(const double *__restrict__ x, double *__restrict__ y, int64_t n, int seg, int warm,
          double intensity, double maxdess, double iirAmount)
{
    const int64_t lane = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t s0 = lane * seg; if (s0 >= n) return;
    const int64_t s1 = min(s0 + (int64_t)seg, n);
    double s1v = 0, s2v = 0, s3v = 0, ratioA = 1.0, ratioB = 1.0, iirA = 0, iirB = 0;
    int64_t i = max((int64_t)0, s0 - warm);
    for (; i < s1; i++) {
        double sample = x[i];
        s3v = s2v; s2v = s1v; s1v = sample;
        const double m1 = (s1v - s2v) * ((s1v - s2v) / 1.3);
        const double m2 = (s2v - s3v) * ((s1v - s2v) / 1.3);
        double sense = (m1 - m2) * ((m1 - m2) / 1.3);
        const double attackspeed = 7.0 + sense * 1024;
        sense = 1.0 + intensity * intensity * sense;
        sense = fmin(sense, intensity);
        const double recovery = 1.0 + (0.01 / sense);
        const double offset = 1.0 - fabs(sample);
        if (i & 1) {        // flip starts at 0 -> B on even samples, A on odd samples
            iirA = (iirA * (1.0 - (offset * iirAmount))) + (sample * (offset * iirAmount));
            if (ratioA < sense) ratioA = ((ratioA * attackspeed) + sense) / (attackspeed + 1.0);
            else ratioA = 1.0 + ((ratioA - 1.0) / recovery);
            ratioA = fmin(ratioA, maxdess);
            sample = iirA + ((sample - iirA) / ratioA);
        } else {
            iirB = (iirB * (1.0 - (offset * iirAmount))) + (sample * (offset * iirAmount));
            if (ratioB < sense) ratioB = ((ratioB * attackspeed) + sense) / (attackspeed + 1.0);
            else ratioB = 1.0 + ((ratioB - 1.0) / recovery);
            ratioB = fmin(ratioB, maxdess);
            sample = iirB + ((sample - iirB) / ratioB);
        }
        if (i >= s0) y[i] = sample;
    }
}

Sig jt_deesser(jt_ctx *c, const Sig &in, double intensity_opt, double max_opt, double freq_opt)
{
    if (in.fmt != JT_FMT_DBL) JT_THROW(JT_ERR_INVALID_ARG, "deesser expects f64 input");
    Sig o = in; o.d = jt_dalloc<double>(c, in.n);
    if (in.n <= 0) return o;
    const double overallscale = in.rate < 44100 ? 44100.0 / in.rate : in.rate / 44100.0;
    const double intensity = pow(intensity_opt, 5) * (8192 / overallscale);
    const double maxdess = 1.0 / pow(10.0, ((max_opt - 1.0) * 48.0) / 20);
    const double iirAmount = pow(freq_opt, 2) / overallscale;
    // slowest-forgetting state: ratio release, factor 1/(1 + 0.01/sense) per two samples with sense <= maxdess
    int64_t warm = (int64_t)(37.0 * 2.0 * std::fmax(maxdess, 1.0) / 0.01) + 64;
    if (warm & 1) warm++;                                     // keep the A/B alternation phase
    if (warm > (1 << 22)) warm = 1 << 22;
    const int seg = 32768;
    const int64_t lanes = (in.n + seg - 1) / seg;
    JtLaunch L(c, "deesser");
    k_deesser<<<(int)((lanes + 63) / 64), 64, 0, c->stream>>>((const double *)in.d, (double *)o.d, in.n, seg, (int)warm, intensity, maxdess, iirAmount);
    return o;
}
