// loudnorm DYNAMIC mode (libavfilter/af_loudnorm.c) at its forced 192 kHz / f64 links -- the path the reference lands on
// whenever the linear-mode preconditions of its Pass-4 spec fail (measured_LRA printing as 0.00 on steady material,
// measured_LRA above the target, a projected peak above TP: internal/processor/normalise.go:683-693) and that it
// survives by resampling back with aresample=<rate> (normalise.go:1294-1304).
//
// The filter is a chain of three sequential state machines.  Cut along their data dependences instead of along time:
//   1. GAIN TARGETS: one value per 100 ms frame (delta[]), a function of the input meter's short-term / integrated /
//      relative-threshold values after that frame -- i.e. of the per-100 ms K-weighted energies the meter kernel
//      (k_r128.cu) already produces.  36 000 values per hour: replayed on the host (LnMeterState), including the
//      21-tap Gaussian that turns them into the per-frame gain pair.  The one feedback from the OUTPUT -- while the
//      stream has not yet reached the measured threshold the filter watches its own output's short-term loudness --
//      is resolved by speculation: run a prefix assuming the switch never happens, read the output meter, place the
//      switch, then run the stream once with the final targets (outputs before the switch do not depend on it).
//   2. GAIN APPLICATION: out-of-place, embarrassingly parallel (piecewise-linear gain over the 3 s delayed input).
//   3. PEAK LIMITER: a four-state machine over a 210 ms ring.  The ring is a sliding window over the gained stream,
//      so the limiter runs IN PLACE on the linear array; it is idle (state OUT, nothing but scanning) wherever no
//      sample exceeds the ceiling, so the host cuts the stream at idle stretches (per-frame maxima from a parallel
//      reduction) and each active region is walked by one warp: the state machine runs warp-uniform, its two data-
//      parallel parts -- scanning for the next peak, multiplying a run of samples by the envelope -- use the 32 lanes.
//      The ring's corner cases are kept: the FINAL frame refills the ring with one constant gain (two arrays: the
//      ramped stream A up to the refill point + 110 ms, the refilled stream B after it), positions past the ring's
//      written end read the slot's previous occupant, the first call's prev-sample rule, negative sample counters.
// No FMA contraction anywhere on the sample path: the reference is x86-64 code without FMA.
#include "jt_internal.h"
#include "jt_device.cuh"
#include <algorithm>
#include <cstdio>
#include <climits>

namespace {
constexpr int F100 = 19200, F3000 = 576000, LBUF = 40320, LOOK = 1920;      // frame_size(192000, 100 / 3000 / 210 / 10)
enum { ST_OUT = 0, ST_ATTACK, ST_SUSTAIN, ST_RELEASE };

struct LnGeom {
    int64_t N;        // input == output samples
    int64_t Q;        // samples emitted before the FINAL frame (== N - 556800)
    int K;            // INNER frames
    int nbK;          // samples of the last INNER frame (19200 unless the stream ends inside it)
    int n_calls;      // limiter calls: 1 (first) + K (inner) + 29 (final)
};

// geometry of limiter call c: output window start P, samples nb, end of the ring's written data, array stage
struct LnCall { int64_t P, E; int nb; bool stageB, first; };
__host__ __device__ inline LnCall ln_call(const LnGeom &G, int c)
{
    LnCall k;
    if (c <= G.K) {
        k.P = (int64_t)F100 * c; k.nb = (c == G.K && c > 0) ? G.nbK : F100; k.stageB = false; k.first = c == 0;
        k.E = c == 0 ? LBUF : k.P + (LBUF - F100) + k.nb;
    } else {
        k.P = G.Q + (int64_t)F100 * (c - G.K - 1); k.nb = F100; k.stageB = true; k.first = false; k.E = k.P + LBUF;
    }
    return k;
}

struct LnArrays { double *lim; double *atail; };        // lim[0 .. N + LBUF), atail[0 .. LBUF - F100)
__device__ __forceinline__ double *ln_at(const LnArrays &A, const LnGeom &G, const LnCall &k, int64_t p)
{
    if (p >= k.E) p -= LBUF;                             // past the written end: the slot still holds its previous occupant
    if (k.stageB || p < G.Q) return A.lim + p;
    return A.atail + (p - G.Q);
}

// ---- 2. gain application ----------------------------------------------------------------------
template <class T>
__global__ void k_ln_gain(const T *__restrict__ x, LnGeom G, const double *__restrict__ g, const double *__restrict__ gn,
                          double d0, double offset, int64_t a_end, LnArrays A)
{
    int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; j < a_end; j += stride) {
        const double v = jt_as_f64(x[j]);
        double o;
        if (j < LBUF) o = __dmul_rn(__dmul_rn(v, d0), offset);
        else {
            const int64_t t = j - LBUF; const int k = (int)(t / F100) + 1; const int n = (int)(t - (int64_t)(k - 1) * F100);
            const int nb = k == G.K ? G.nbK : F100;
            const double gk = g[k], gd = __dsub_rn(gn[k], gk);
            const double gg = __dadd_rn(gk, __dmul_rn(__ddiv_rn((double)n, (double)nb), gd));
            o = __dmul_rn(__dmul_rn(v, gg), offset);
        }
        if (j < G.Q) A.lim[j] = o; else A.atail[j - G.Q] = o;
    }
}
template <class T>
__global__ void k_ln_gain_final(const T *__restrict__ x, LnGeom G, double gfin, double offset, double *__restrict__ lim)
{
    int64_t j = G.Q + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; j < G.N + LBUF; j += stride) lim[j] = j < G.N ? __dmul_rn(__dmul_rn(jt_as_f64(x[j]), gfin), offset) : 0.0;
}

// ---- per-call maxima of the range an idle limiter scans (plus the first call's 10 ms head) ------
__global__ void __launch_bounds__(256)
k_ln_call_max(LnGeom G, LnArrays A, int c_end, double *__restrict__ call_max)
{
    __shared__ double red[8];
    for (int c = blockIdx.x; c < c_end; c += gridDim.x) {
        const LnCall k = ln_call(G, c);
        const int64_t a = k.first ? 0 : k.P + LOOK, b = k.P + LOOK + k.nb + 12;     // +12: the "next" / 10-sample hold checks
        double m = 0.0;
        for (int64_t p = a + threadIdx.x; p < b; p += 256) m = fmax(m, fabs(*ln_at(A, G, k, p)));
        m = jt_warp_max(m);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
        __syncthreads();
        if (threadIdx.x == 0) { for (int i = 1; i < 8; i++) m = fmax(m, red[i]); call_max[c] = m; }
        __syncthreads();
    }
}

// ---- 3. the limiter, one warp per active region ------------------------------------------------
struct LnRegion { int c0, c1; };        // calls [c0, c1); the limiter is idle (OUT) entering c0

struct LnLim {
    const LnGeom &G; const LnArrays &A; LnCall k; double ceiling;
    int state = ST_OUT, env_cnt = 0, attack_length = LOOK; int64_t env_pos = 0; double gr0 = 1.0, gr1 = 1.0, prev = 0.0;
    int lane;
    __device__ LnLim(const LnGeom &g, const LnArrays &a, double ceil_) : G(g), A(a), ceiling(ceil_) { lane = threadIdx.x & 31; }
    __device__ __forceinline__ double rd(int64_t p) const { return fabs(*ln_at(A, G, k, p)); }

    // detect_peak(): first n in [0, count) at start + n that is a local maximum above the ceiling not exceeded within the
    // next 11 samples; prev-sample semantics of the scalar loop (a rejected candidate does not update prev)
    __device__ int detect(int64_t start, int count, double *value, int64_t *pos)
    {
        if (k.first) prev = rd(start - 1);
        for (int n = 0; n < count; n += 32) {
            const int chunk = min(32, count - n);
            const double mine = lane < chunk ? rd(start + n + lane) : 0.0;
            const unsigned hot = __ballot_sync(0xffffffffu, lane < chunk && mine > ceiling);
            if (!hot) { prev = __shfl_sync(0xffffffffu, mine, chunk - 1); continue; }
            const int j0 = __ffs(hot) - 1;
            if (j0 > 0) prev = __shfl_sync(0xffffffffu, mine, j0 - 1);
            for (int j = j0; j < chunk; j++) {
                const int64_t p = start + n + j;
                const double cur = __shfl_sync(0xffffffffu, mine, j);
                const double next = rd(p + 1);
                if (prev <= cur && next <= cur && cur > ceiling && (n + j) > 0) {
                    bool detected = true;
                    for (int i = 2; i < 12; i++) if (rd(p + i) > cur) { detected = false; break; }
                    if (detected) { prev = cur; *value = cur; *pos = p; return n + j; }
                } else prev = cur;
            }
        }
        return -1;
    }
    // buf[env_pos + t] *= env(env_cnt + t), t in [0, m)
    template <int KIND> __device__ void apply(int m, int length)
    {
        for (int t = lane; t < m; t += 32) {
            double env;
            if (KIND == ST_SUSTAIN) env = gr1;
            else {
                const double fr = __ddiv_rn((double)(env_cnt + t), (double)(length - 1));
                env = KIND == ST_ATTACK ? __dsub_rn(gr0, __dmul_rn(fr, __dsub_rn(gr0, gr1))) : __dadd_rn(gr0, __dmul_rn(fr, __dsub_rn(gr1, gr0)));
            }
            double *q = ln_at(A, G, k, env_pos + t);
            *q = __dmul_rn(*q, env);
        }
        __syncwarp();
    }

    __device__ void call()      // true_peak_limiter() without its copy-out / clamp tail
    {
        const int nb = k.nb; int smp_cnt = 0;
        if (k.first) {
            double mx = 0.0;
            for (int n = lane; n < LOOK; n += 32) mx = fmax(mx, rd(n));
            mx = jt_warp_max(mx);
            if (mx > ceiling) {
                gr1 = __ddiv_rn(ceiling, mx); state = ST_SUSTAIN;
                for (int n = lane; n < LOOK; n += 32) { double *q = ln_at(A, G, k, n); *q = __dmul_rn(*q, gr1); }
                __syncwarp();
            }
        }
        do {
            double value = 0.0; int64_t pos = 0;
            switch (state) {
            case ST_OUT: {
                const int d = detect(k.P + smp_cnt + LOOK, nb - smp_cnt, &value, &pos);
                if (d != -1) {
                    env_cnt = 0; smp_cnt += d - attack_length;
                    gr0 = 1.0; gr1 = __ddiv_rn(ceiling, value); state = ST_ATTACK;
                    env_pos = pos - attack_length;
                } else smp_cnt = nb;
            } break;
            case ST_ATTACK: {
                const int m = max(0, min(attack_length - env_cnt, nb - smp_cnt));
                apply<ST_ATTACK>(m, attack_length);
                env_cnt += m; env_pos += m; smp_cnt += m;
                if (smp_cnt < nb) { env_cnt = 0; attack_length = LOOK; state = ST_SUSTAIN; }
            } break;
            case ST_SUSTAIN: {
                const int d = detect(k.P + smp_cnt + LOOK, nb, &value, &pos);
                if (d == -1) { state = ST_RELEASE; gr0 = gr1; gr1 = 1.0; env_cnt = 0; }
                else {
                    const double gr = __ddiv_rn(ceiling, value);
                    if (gr < gr1) { state = ST_ATTACK; attack_length = d <= 1 ? 2 : d; gr0 = gr1; gr1 = gr; env_cnt = 0; }
                    else {
                        const int m = min(d, nb - smp_cnt);
                        apply<ST_SUSTAIN>(m, 0);
                        env_cnt = m; env_pos += m; smp_cnt += m;
                    }
                }
            } break;
            case ST_RELEASE: {
                const int m = max(0, min(F100 - env_cnt, nb - smp_cnt));
                apply<ST_RELEASE>(m, F100);
                env_cnt += m; env_pos += m; smp_cnt += m;
                if (smp_cnt < nb) { env_cnt = 0; state = ST_OUT; }
            } break;
            }
        } while (smp_cnt < nb);
    }
};

__global__ void __launch_bounds__(32)
k_ln_limiter(LnGeom G, LnArrays A, double ceiling, const LnRegion *__restrict__ regions, int c_end, int *__restrict__ not_idle)
{
    const LnRegion R = regions[blockIdx.x];
    LnLim L(G, A, ceiling);
    if (R.c0 > 0) { L.k = ln_call(G, R.c0 - 1); L.prev = L.rd(L.k.P + LOOK + L.k.nb - 1); }     // last sample the idle scan saw
    for (int c = R.c0; c < R.c1; c++) {
        L.k = ln_call(G, c);
        if (c == G.K + 1 && L.state != ST_OUT) L.env_pos = G.Q + (L.env_pos % LBUF);   // the FINAL frame re-bases the ring at slot 0
        L.call();
    }
    if (L.state != ST_OUT && R.c1 < c_end && L.lane == 0) atomicExch(not_idle, 1);
}

__global__ void k_ln_clamp(double *__restrict__ lim, int64_t n, double ceiling)
{
    int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; j < n; j += stride) { const double v = lim[j]; if (fabs(v) > ceiling) lim[j] = ceiling * (v < 0 ? -1 : 1); }
}

// ---- 1. gain targets -----------------------------------------------------------------------------
struct GainPlan { std::vector<double> g, gn; double d0 = 1, gfin = 1; bool above0 = true; };

struct Gauss {
    double w[21];
    Gauss() {
        double total = 0.0; const double sigma = 3.5; const int offset = 21 / 2;
        const double c1 = 1.0 / (sigma * sqrt(2.0 * M_PI)), c2 = 2.0 * pow(sigma, 2.0);
        for (int i = 0; i < 21; i++) { const int x = i - offset; w[i] = c1 * exp(-(pow(x, 2.0) / c2)); total += w[i]; }
        const double adjust = 1.0 / total;
        for (int i = 0; i < 21; i++) w[i] *= adjust;
    }
    double filter(const double *delta, int index) const {
        double result = 0.;
        index = index - 10 > 0 ? index - 10 : index + 20;
        for (int i = 0; i < 21; i++) result += delta[((index + i) < 30) ? (index + i) : (index + i - 30)] * w[i];
        return result;
    }
};

// Replays filter_frame()'s FIRST_FRAME / INNER_FRAME bookkeeping over the input meter's per-100 ms energies hp[].
// flip: the INNER frame at which above_threshold becomes 1 (0 = from the first frame on, INT_MAX = never); the frames
// whose short-term OUTPUT loudness would be consulted are exactly those before it.
static GainPlan gain_plan(const LnGeom &G, const double *hp, const jt_loudnorm_opts &o, int flip)
{
    static const Gauss gauss;
    GainPlan P; P.g.assign(G.K + 2, 1.0); P.gn.assign(G.K + 2, 1.0);
    LnMeterState st(F100, o.dual_mono != 0);
    for (int k = 0; k < 30; k++) st.add_tick(hp, k);
    double delta[30], prev_delta; int index = 1; int above;
    {
        const double shortterm = st.shortterm(hp, 29);
        double env;
        if (shortterm < o.measured_thresh) { above = 0; env = shortterm <= -70. ? 0. : o.I - o.measured_I; }
        else { above = 1; env = shortterm <= -70. ? 0. : o.I - shortterm; }
        for (int n = 0; n < 30; n++) delta[n] = pow(10., env / 20.);
        prev_delta = delta[index];
        P.d0 = delta[index]; P.above0 = above != 0;
    }
    for (int k = 1; k <= G.K; k++) {
        P.g[k] = gauss.filter(delta, index + 10 < 30 ? index + 10 : index + 10 - 30);
        P.gn[k] = gauss.filter(delta, index + 11 < 30 ? index + 11 : index + 11 - 30);
        const bool full = k < G.K || G.nbK == F100;
        if (!full) { index++; if (index >= 30) index -= 30; break; }    // the short last frame's target is never read (targets act 1 s later)
        const int64_t tick = 29 + k;
        st.add_tick(hp, tick);
        const double global = st.global(), shortterm = st.shortterm(hp, tick), relative_threshold = st.relative_threshold();
        if (above == 0) {
            if (shortterm > o.measured_thresh) prev_delta *= 1.0058;
            if (k >= flip) above = 1;          // shortterm_out >= target_i, as established on the output meter
        }
        if (shortterm < relative_threshold || shortterm <= -70. || above == 0) delta[index] = prev_delta;
        else {
            const double d = shortterm - global;
            const double env_global = fabs(d) < (o.LRA / 2.) ? d : (o.LRA / 2.) * (d < 0 ? -1 : 1);
            const double env_shortterm = o.I - shortterm;
            delta[index] = pow(10., (env_global + env_shortterm) / 20.);
        }
        prev_delta = delta[index];
        index++; if (index >= 30) index -= 30;
    }
    P.gfin = gauss.filter(delta, index + 10 < 30 ? index + 10 : index + 10 - 30);
    return P;
}

template <class T>
static void launch_gain(jt_ctx *c, const Sig &x, const LnGeom &G, const GainPlan &P, double offset, int64_t a_end, bool with_final, const LnArrays &A)
{
    const double *d_g = jt_dalloc<double>(c, P.g.size()), *d_gn = jt_dalloc<double>(c, P.gn.size());
    double *h = jt_pinned<double>(c, 2 * P.g.size());
    std::copy(P.g.begin(), P.g.end(), h); std::copy(P.gn.begin(), P.gn.end(), h + P.g.size());
    JT_CUDA(cudaMemcpyAsync((void *)d_g, h, sizeof(double) * P.g.size(), cudaMemcpyHostToDevice, c->stream));
    JT_CUDA(cudaMemcpyAsync((void *)d_gn, h + P.g.size(), sizeof(double) * P.g.size(), cudaMemcpyHostToDevice, c->stream));
    JtLaunch L(c, "loudnorm_dynamic:gain", with_final ? 2 : 1);
    k_ln_gain<T><<<jt_grid_for(a_end, 256, c->num_sms, 16), 256, 0, c->stream>>>((const T *)x.d, G, d_g, d_gn, P.d0, offset, a_end, A);
    if (with_final)
        k_ln_gain_final<T><<<jt_grid_for(G.N + LBUF - G.Q, 256, c->num_sms, 16), 256, 0, c->stream>>>((const T *)x.d, G, P.gfin, offset, A.lim);
}

// limiter over calls [0, c_end): per-call maxima -> idle stretches -> one warp per active region
static void run_limiter(jt_ctx *c, const LnGeom &G, const LnArrays &A, double ceiling, int c_end)
{
    double *d_max = jt_dalloc<double>(c, c_end), *h_max = jt_pinned<double>(c, c_end);
    { JtLaunch L(c, "loudnorm_dynamic:call_max"); k_ln_call_max<<<std::min(c_end, c->num_sms * 8), 256, 0, c->stream>>>(G, A, c_end, d_max); }
    JT_CUDA(cudaMemcpyAsync(h_max, d_max, sizeof(double) * c_end, cudaMemcpyDeviceToHost, c->stream));
    JT_CUDA(cudaStreamSynchronize(c->stream));
    int *d_flag = jt_dalloc<int>(c, 1), *h_flag = jt_pinned<int>(c, 1);
    // SUSTAIN -> RELEASE (100 ms) -> OUT takes at most three idle frames' worth of samples; a region whose warp still
    // ends active (never observed) fails the call rather than emit a stream limited from the wrong state
    const int64_t idle_needed = 3 * (int64_t)F100;
    std::vector<LnRegion> regions;
    int c0 = -1; int64_t idle = 0;
    for (int cc = 0; cc < c_end; cc++) {
        const bool hot = h_max[cc] > ceiling;
        if (hot) { if (c0 < 0) c0 = cc; idle = 0; }
        else if (c0 >= 0) { idle += ln_call(G, cc).nb; if (idle >= idle_needed) { regions.push_back({c0, cc + 1}); c0 = -1; } }
    }
    if (c0 >= 0) regions.push_back({c0, c_end});
    if (regions.empty()) return;
    LnRegion *d_reg = jt_dalloc<LnRegion>(c, regions.size()), *h_reg = jt_pinned<LnRegion>(c, regions.size());
    std::copy(regions.begin(), regions.end(), h_reg);
    JT_CUDA(cudaMemcpyAsync(d_reg, h_reg, sizeof(LnRegion) * regions.size(), cudaMemcpyHostToDevice, c->stream));
    JT_CUDA(cudaMemsetAsync(d_flag, 0, sizeof(int), c->stream));
    { JtLaunch L(c, "loudnorm_dynamic:limiter"); k_ln_limiter<<<(int)regions.size(), 32, 0, c->stream>>>(G, A, ceiling, d_reg, c_end, d_flag); }
    JT_CUDA(cudaMemcpyAsync(h_flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    JT_CUDA(cudaStreamSynchronize(c->stream));
    if (*h_flag) JT_THROW(JT_ERR_CUDA, "internal: a loudnorm limiter region did not return to idle");
}
}   // namespace

// The filter over its 192 kHz input link.  Returns the output link (f64, 192 kHz, same length).  The input meter must have
// been launched by the caller (pd_in, on x) -- its per-100 ms energies drive the gain targets.  *normalization_type
// follows uninit()'s "linear"/"dynamic" print: a stream shorter than 3 s falls back to one gain and prints "linear".
Sig jt_loudnorm_dynamic(jt_ctx *c, const Sig &x, const jt_loudnorm_opts &o, LoudnormPending &pd_in, int *normalization_type)
{
    if (x.rate != 192000) JT_THROW(JT_ERR_INVALID_ARG, "internal: loudnorm dynamic mode runs at 192 kHz");
    const int64_t N = x.n;
    if (N <= 0) { *normalization_type = 1; Sig e = x; e.fmt = JT_FMT_DBL; e.d = jt_dalloc<double>(c, 1); return e; }
    JT_CUDA(cudaEventSynchronize(pd_in.ev));
    const double *hp = pd_in.hp, *hk = pd_in.hk;
    const double ceiling = pow(10., o.TP / 20.), offset = pow(10., o.offset / 20.);
    if (N < F3000) {
        // FIRST_FRAME shorter than 3 s: one gain from the measured integrated loudness, capped by the peak
        LnMeterState st(F100, o.dual_mono != 0);
        for (int64_t k = 0; k < pd_in.nfull; k++) st.add_tick(hp, k);
        double true_peak = 0; for (int64_t k = 0; k < pd_in.nt; k++) true_peak = std::max(true_peak, hk[k]);
        const double global = st.global();
        const double off = pow(10., (o.I - global) / 20.), offset_tp = true_peak * off;
        const double gain = offset_tp < ceiling ? off : ceiling / true_peak;
        *normalization_type = 0;
        return jt_gain_f64(c, x, gain);
    }
    *normalization_type = 1;
    LnGeom G; G.N = N; G.Q = N - (F3000 - F100);
    G.K = (int)((N - F3000 + F100 - 1) / F100); G.nbK = G.K ? (int)((N - F3000) - (int64_t)(G.K - 1) * F100) : F100;
    G.n_calls = 1 + G.K + (F3000 - F100) / F100;
    LnArrays A; A.lim = jt_dalloc<double>(c, (size_t)N + LBUF); A.atail = jt_dalloc<double>(c, LBUF - F100);
    auto gain = [&](const GainPlan &P, int64_t a_end, bool with_final) {
        if (x.fmt == JT_FMT_FLT) launch_gain<float>(c, x, G, P, offset, a_end, with_final, A);
        else if (x.fmt == JT_FMT_DBL) launch_gain<double>(c, x, G, P, offset, a_end, with_final, A);
        else launch_gain<int16_t>(c, x, G, P, offset, a_end, with_final, A);
    };
    const int64_t a_all = G.Q + (LBUF - F100);
    int flip = 0;
    {
        GainPlan P0 = gain_plan(G, hp, o, INT_MAX);
        if (!P0.above0 && G.K > 0) {
            // the stream starts below the measured threshold: find the frame at which the OUTPUT's short-term loudness
            // first reaches the target, on prefixes processed as if it never did
            flip = INT_MAX;
            for (int H = std::min(G.K, 1200); ; H = std::min(G.K, H * 8)) {
                const int c_end = 1 + H;                                   // first + H inner calls
                const LnCall last = ln_call(G, c_end - 1);
                const int64_t a_end = std::min(a_all, last.P + LBUF);
                gain(P0, a_end, false);
                run_limiter(c, G, A, ceiling, c_end);
                const int64_t n_done = last.P + last.nb;
                { JtLaunch L(c, "loudnorm_dynamic:clamp"); k_ln_clamp<<<jt_grid_for(n_done, 256, c->num_sms, 16), 256, 0, c->stream>>>(A.lim, std::min(n_done, G.Q), ceiling); }
                Sig pre; pre.fmt = JT_FMT_DBL; pre.rate = 192000; pre.n = std::min(n_done, G.Q); pre.d = A.lim;
                LoudnormPending po; jt_loudnorm_meter_launch(c, pre, o.dual_mono != 0, po);
                JT_CUDA(cudaEventSynchronize(po.ev));
                LnMeterState so(F100, o.dual_mono != 0);
                for (int k = 1; k <= H && k < po.nfull; k++)
                    if (so.shortterm(po.hp, k) >= o.I) { flip = k; break; }
                if (flip != INT_MAX || H >= G.K) break;
            }
        }
    }
    GainPlan P = gain_plan(G, hp, o, flip);
    gain(P, a_all, true);
    run_limiter(c, G, A, ceiling, G.n_calls);
    { JtLaunch L(c, "loudnorm_dynamic:clamp"); k_ln_clamp<<<jt_grid_for(N, 256, c->num_sms, 16), 256, 0, c->stream>>>(A.lim, N, ceiling); }
    Sig out; out.fmt = JT_FMT_DBL; out.rate = 192000; out.n = N; out.d = A.lim;
    return out;
}

// Dynamic mode meters the flush frame -- the last 2.9 s of the stream, rebuilt from the delay line -- a second time
// (filter_frame() feeds every frame to r128_in): the JSON's input_* values are those of the stream followed by its
// tail.  `tail` receives the per-100 ms values of a short signal [context + tail] laid on the stream's tick grid;
// *first_tick is the global index of its first tick.  The caller takes ticks >= N / s100 from it.
void jt_loudnorm_tail_launch(jt_ctx *c, const Sig &x_end /* ends at the stream's end */, int64_t N_stream, bool dual_mono,
                             LoudnormPending &tail, int64_t *first_tick)
{
    const int s100 = (x_end.rate + 5) / 10;
    const int64_t T = 30 * (int64_t)s100 - s100;                      // 556800 at 192 kHz
    const int64_t W = N_stream % s100 + 10 * (int64_t)s100;
    if (x_end.n < T || x_end.n < W || N_stream < W) JT_THROW(JT_ERR_INVALID_ARG, "internal: loudnorm tail window");
    const size_t b = jt_fmt_bytes(x_end.fmt);
    Sig t = x_end; t.n = W + T; t.d = jt_dalloc_bytes(c, (size_t)(W + T) * b);
    JT_CUDA(cudaMemcpyAsync(t.d, (const char *)x_end.d + (size_t)(x_end.n - W) * b, (size_t)W * b, cudaMemcpyDeviceToDevice, c->stream));
    JT_CUDA(cudaMemcpyAsync((char *)t.d + (size_t)W * b, (const char *)x_end.d + (size_t)(x_end.n - T) * b, (size_t)T * b, cudaMemcpyDeviceToDevice, c->stream));
    jt_loudnorm_meter_launch(c, t, dual_mono, tail);
    *first_tick = (N_stream - W) / s100;
}
