// anlmdn -- non-local means denoiser (libavfilter/af_anlmdn.c), float32:
// "anlmdn=s=0.00001:p=0.0060:r=0.0020:m=3" (reference: filters.go:95-100, 811-816;
// scripts/anlmdn-matrix-spike.sh:292-320).  K = p*fs patch radius, S = r*fs search radius,
// hops of H = 2K+1 samples, 2S lags per output sample.  The dominant Pass-2 cost.
//
// One CTA per hop, one thread per PAIR of lags (j = i-S+v and j = i+1+v share the two patch-edge
// samples of i).  The hop's window (H + 2(K+S) floats) is staged in shared memory; every thread
// keeps its lags' running patch distances in registers and updates them with the same unfused
// float operations, in the same order, as the scalar C (full SSD at the hop's first sample,
// then add-new/subtract-old), so distances are bit-identical to a sequential run.
// The kernel is instruction-issue bound (ncu: 91 % issue-active), so the steady state is written
// for instruction count: eight samples per loop trip, every load a base register + immediate
// (with win[] indices: old/new patch edge of i = S+t-1 / S+t+2K, of lag v = t+v-1 / t+v+2K, of lag
// v+S+1 = t+v+S / t+v+S+2K+1), one warp vote per eight samples for the rare samples where any lag
// falls inside the smoothing cut-off.  Only then are weights computed and reduced per warp with
// shuffles; partial sums of the warps are combined once per 32 samples (double-buffered: one
// barrier per 32 samples).
#include "jt_internal.h"
#include "jt_device.cuh"

#define NLM_CHUNK 32
#define NLM_U 8
#define NLM_MAXWARPS 16

struct NlmK { float sw, smooth, lut_scale, inv_lut_scale, maybe; };   // maybe: distances below it MAY pass d*sw < smooth (superset test)

// weights of one sample for this thread's two lags; returns true when either lag is inside the cut-off
__device__ __forceinline__ bool nlm_weights(float c1, float c2, float x1, float x2, const NlmK &k, float &pw, float &qw)
{
    pw = 0.f; qw = 0.f; bool in = false;
    const float w1 = __fmul_rn(c1, k.sw), w2 = __fmul_rn(c2, k.sw);
    if (!(w1 >= k.smooth)) {
        const unsigned idx = (unsigned)__fmul_rn(w1, k.lut_scale);
        const float w = __expf(-(float)idx * k.inv_lut_scale);
        pw = __fmul_rn(w, x1); qw = w; in = true;
    }
    if (!(w2 >= k.smooth)) {
        const unsigned idx = (unsigned)__fmul_rn(w2, k.lut_scale);
        const float w = __expf(-(float)idx * k.inv_lut_scale);
        pw += __fmul_rn(w, x2); qw += w; in = true;
    }
    return in;
}

__global__ void __launch_bounds__(512)
k_anlmdn(const float *__restrict__ x, float *__restrict__ y, int64_t n, int K, int S, NlmK P, int64_t n_hops)
{
    extern __shared__ float win[];                       // N floats (+ slack read by clamped idle lanes)
    __shared__ float part[2][NLM_CHUNK][NLM_MAXWARPS][2];
    const int H = 2 * K + 1, N = H + 2 * (K + S), offset = N - H;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const bool live = (int)threadIdx.x < S;
    const int v = live ? threadIdx.x : S - 1;            // idle lanes of the last warp shadow lag S-1 and never contribute
    const int K2 = 2 * K;
    for (int64_t h = blockIdx.x; h < n_hops; h += gridDim.x) {
        const int64_t pos = h * (int64_t)H;
        const int nb = (int)min((int64_t)H, n - pos);
        __syncthreads();
        for (int j = threadIdx.x; j < N; j += blockDim.x) {
            const int64_t s = pos - offset + j;
            win[j] = (s >= 0 && s < pos + nb) ? x[s] : 0.f;
        }
        __syncthreads();
        // full patch distances at the hop's first sample (t = 0): sum over k = -K..K in order
        float c1 = 0.f, c2 = 0.f;
        {
            const float *a = win + S, *p1 = win + v, *p2 = win + v + S + 1;
            int k = 0;
            for (; k + 8 <= H; k += 8) {
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    const float av = a[k + u];
                    const float t1 = __fsub_rn(av, p1[k + u]), t2 = __fsub_rn(av, p2[k + u]);
                    c1 = __fadd_rn(c1, __fmul_rn(t1, t1)); c2 = __fadd_rn(c2, __fmul_rn(t2, t2));
                }
            }
            for (; k < H; k++) {
                const float av = a[k];
                const float t1 = __fsub_rn(av, p1[k]), t2 = __fsub_rn(av, p2[k]);
                c1 = __fadd_rn(c1, __fmul_rn(t1, t1)); c2 = __fadd_rn(c2, __fmul_rn(t2, t2));
            }
        }
        // sample t = 0 is a chunk of its own, then chunks of 32 samples starting at t = 1
        int pb = 0;
        for (int c0 = 0; c0 < H; c0 = (c0 == 0) ? 1 : c0 + NLM_CHUNK, pb ^= 1) {
            const int cn = c0 == 0 ? 1 : min(NLM_CHUNK, H - c0);
            part[pb][lane][warp][0] = 0.f; part[pb][lane][warp][1] = 0.f;
            __syncwarp();
            if (c0 == 0) {
                if (c1 < 0.f) c1 = 0.f;
                if (c2 < 0.f) c2 = 0.f;
                float pw, qw;
                const bool in = nlm_weights(c1, c2, win[K + v], win[K + v + S + 1], P, pw, qw) && live;
                if (__any_sync(0xffffffffu, in)) {
                    if (!in) { pw = 0.f; qw = 0.f; }
                    for (int o = 16; o; o >>= 1) { pw += __shfl_xor_sync(0xffffffffu, pw, o); qw += __shfl_xor_sync(0xffffffffu, qw, o); }
                    if (lane == 0) { part[pb][0][warp][0] = pw; part[pb][0][warp][1] = qw; }
                }
            } else {
                const float *q0 = win + c0 + v - 1, *q1 = q0 + K2 + 1, *q2 = q0 + S + 1, *q3 = q2 + K2 + 1;
                const float *u0 = win + S + c0 - 1, *u1 = u0 + K2 + 1;
                const float *x1 = win + K + c0 + v, *x2 = x1 + S + 1;
                int ci = 0;
                for (; ci + NLM_U <= cn; ci += NLM_U) {
                    float d1[NLM_U], d2[NLM_U];
                    bool any = false;
#pragma unroll
                    for (int u = 0; u < NLM_U; u++) {
                        const float ao = u0[ci + u], an = u1[ci + u];
                        const float a1 = __fsub_rn(ao, q0[ci + u]), b1 = __fsub_rn(an, q1[ci + u]);
                        const float a2 = __fsub_rn(ao, q2[ci + u]), b2 = __fsub_rn(an, q3[ci + u]);
                        c1 = __fadd_rn(c1, __fadd_rn(-__fmul_rn(a1, a1), __fmul_rn(b1, b1)));
                        c2 = __fadd_rn(c2, __fadd_rn(-__fmul_rn(a2, a2), __fmul_rn(b2, b2)));
                        c1 = fmaxf(c1, 0.f); c2 = fmaxf(c2, 0.f);        // "if (d < 0) d = 0"
                        d1[u] = c1; d2[u] = c2;
                        any |= !(fminf(c1, c2) >= P.maybe);            // cheap superset of (d * sw < smooth); the exact test follows
                    }
                    if (__any_sync(0xffffffffu, any && live)) {
#pragma unroll
                        for (int u = 0; u < NLM_U; u++) {
                            float pw, qw;
                            const bool in = nlm_weights(d1[u], d2[u], x1[ci + u], x2[ci + u], P, pw, qw) && live;
                            if (!__any_sync(0xffffffffu, in)) continue;
                            if (!in) { pw = 0.f; qw = 0.f; }
                            for (int o = 16; o; o >>= 1) { pw += __shfl_xor_sync(0xffffffffu, pw, o); qw += __shfl_xor_sync(0xffffffffu, qw, o); }
                            if (lane == 0) { part[pb][ci + u][warp][0] = pw; part[pb][ci + u][warp][1] = qw; }
                        }
                    }
                }
                for (; ci < cn; ci++) {
                    const float ao = u0[ci], an = u1[ci];
                    const float a1 = __fsub_rn(ao, q0[ci]), b1 = __fsub_rn(an, q1[ci]);
                    const float a2 = __fsub_rn(ao, q2[ci]), b2 = __fsub_rn(an, q3[ci]);
                    c1 = __fadd_rn(c1, __fadd_rn(-__fmul_rn(a1, a1), __fmul_rn(b1, b1)));
                    c2 = __fadd_rn(c2, __fadd_rn(-__fmul_rn(a2, a2), __fmul_rn(b2, b2)));
                    if (c1 < 0.f) c1 = 0.f;
                    if (c2 < 0.f) c2 = 0.f;
                    float pw, qw;
                    const bool in = nlm_weights(c1, c2, x1[ci], x2[ci], P, pw, qw) && live;
                    if (!__any_sync(0xffffffffu, in)) continue;
                    if (!in) { pw = 0.f; qw = 0.f; }
                    for (int o = 16; o; o >>= 1) { pw += __shfl_xor_sync(0xffffffffu, pw, o); qw += __shfl_xor_sync(0xffffffffu, qw, o); }
                    if (lane == 0) { part[pb][ci][warp][0] = pw; part[pb][ci][warp][1] = qw; }
                }
            }
            __syncthreads();
            // the first cn threads combine the warps' partial sums of this chunk while the others move on:
            // the next chunk writes the other half of part[], and is itself followed by a barrier
            if ((int)threadIdx.x < cn) {
                const int ci = threadIdx.x, t = c0 + ci;
                float Ps = 0.f, Qs = 0.f;
                for (int w = 0; w < nwarp; w++) { Ps += part[pb][ci][w][0]; Qs += part[pb][ci][w][1]; }
                Ps = __fadd_rn(Ps, win[K + S + t]); Qs = __fadd_rn(Qs, 1.f);
                if (t < nb) y[pos + t] = __fdiv_rn(Ps, Qs);
            }
        }
    }
}

static int64_t rescale_near(int64_t a, int64_t b, int64_t c) { return (a * b + c / 2) / c; }

Sig jt_anlmdn(jt_ctx *c, const Sig &in, double strength, double patch_s, double research_s, double smooth_m)
{
    if (in.fmt != JT_FMT_FLT) JT_THROW(JT_ERR_INVALID_ARG, "anlmdn expects float input");
    const int K = (int)rescale_near(llround(patch_s * 1e6), in.rate, 1000000);
    const int S = (int)rescale_near(llround(research_s * 1e6), in.rate, 1000000);
    if (K < 1 || S < 1) JT_THROW(JT_ERR_INVALID_ARG, "anlmdn patch/research too small for %d Hz", in.rate);
    const int H = 2 * K + 1, N = H + 2 * (K + S);
    const int threads = std::max(32, ((S + 31) / 32) * 32);
    if (threads > 512) JT_THROW(JT_ERR_UNSUPPORTED, "anlmdn research radius %d samples (max 512)", S);
    Sig o = in; o.d = jt_dalloc<float>(c, in.n);
    if (in.n <= 0) return o;
    const float m = (float)smooth_m, a = (float)strength;
    const float lut_scale = 1.f / m * (float)(1 << 20);
    const float sw = (65536.f / (4 * K + 2)) / sqrtf(a);
    const float smooth = fminf(m, (float)(1 << 20) / lut_scale);
    const int64_t n_hops = (in.n + H - 1) / H;
    const size_t smem = sizeof(float) * (N + 64);
    JT_CUDA(cudaFuncSetAttribute(k_anlmdn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem, 1024)));
    const int grid = jt_grid_for(n_hops, 1, c->num_sms, 64);
    JtLaunch L(c, "anlmdn");
    NlmK P; P.sw = sw; P.smooth = smooth; P.lut_scale = lut_scale; P.inv_lut_scale = 1.f / lut_scale;
    P.maybe = (smooth / sw) * 1.0001f;
    k_anlmdn<<<grid, threads, smem, c->stream>>>((const float *)in.d, (float *)o.d, in.n, K, S, P, n_hops);
    return o;
}
