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

struct NlmK { float sw, smooth, lut_scale, inv_lut_scale, maybe; };   // maybe: distances below it MAY pass d*sw < smooth (superset test)

// weight of one lag at one sample (LUT index as in the C code); false when the lag is outside the cut-off
__device__ __forceinline__ bool nlm_weight(float c, float xs, const NlmK &k, float &pw, float &qw)
{
    const float w1 = __fmul_rn(c, k.sw);
    if (w1 >= k.smooth) return false;
    const unsigned idx = (unsigned)__fmul_rn(w1, k.lut_scale);
    const float w = __expf(-(float)idx * k.inv_lut_scale);
    pw += __fmul_rn(w, xs); qw += w;
    return true;
}

#define NLM_V 3             // lags v = 3u, 3u+1, 3u+2 (and their mirror lags v+S+1) per thread
#define NLM_G 4             // samples per loop trip

// One CTA per hop; thread u owns the lag pairs of v = 3u .. 3u+2, i.e. six running patch distances.  The kernel
// was bound by shared-memory bandwidth (ncu: LSU 77 %, 32 wavefronts per sample) as much as by issue slots: with
// three adjacent lags per thread the patch-edge samples of lag v+1 at sample t are those of lag v at t+1, so a trip
// over four samples loads 6 values per edge stream instead of 12, and a stride of three floats between lanes is
// conflict-free.  With S = 96 one warp covers all 192 lags of a hop.
__global__ void __launch_bounds__(128)
k_anlmdn(const float *__restrict__ x, float *__restrict__ y, int64_t n, int K, int S, NlmK P, int64_t n_hops)
{
    extern __shared__ float win[];                       // N floats (+ slack)
    __shared__ float part[2][NLM_CHUNK][4][2];
    const int H = 2 * K + 1, N = H + 2 * (K + S), offset = N - H;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const int K2 = 2 * K;
    const int v0 = 3 * threadIdx.x;
    bool live[NLM_V];                                    // lags past S-1 read the zeroed slack behind the window and never contribute
#pragma unroll
    for (int m = 0; m < NLM_V; m++) live[m] = v0 + m < S;
    for (int64_t h = blockIdx.x; h < n_hops; h += gridDim.x) {
        const int64_t pos = h * (int64_t)H;
        const int nb = (int)min((int64_t)H, n - pos);
        __syncthreads();
        for (int j = threadIdx.x; j < N + 120; j += blockDim.x) {
            const int64_t s = pos - offset + j;
            win[j] = (j < N && s >= 0 && s < pos + nb) ? x[s] : 0.f;
        }
        __syncthreads();
        // full patch distances at the hop's first sample (t = 0): sum over k = 0 .. 2K in order, per lag
        float c1[NLM_V], c2[NLM_V];
#pragma unroll
        for (int m = 0; m < NLM_V; m++) { c1[m] = 0.f; c2[m] = 0.f; }
        {
            const float *a = win + S, *q1 = win + v0, *q2 = win + v0 + S + 1;
            float r1a = q1[0], r1b = q1[1], r2a = q2[0], r2b = q2[1];
            for (int k = 0; k < H; k++) {
                const float av = a[k], r1c = q1[k + 2], r2c = q2[k + 2];
                float t;
                t = __fsub_rn(av, r1a); c1[0] = __fadd_rn(c1[0], __fmul_rn(t, t));
                t = __fsub_rn(av, r1b); c1[1] = __fadd_rn(c1[1], __fmul_rn(t, t));
                t = __fsub_rn(av, r1c); c1[2] = __fadd_rn(c1[2], __fmul_rn(t, t));
                t = __fsub_rn(av, r2a); c2[0] = __fadd_rn(c2[0], __fmul_rn(t, t));
                t = __fsub_rn(av, r2b); c2[1] = __fadd_rn(c2[1], __fmul_rn(t, t));
                t = __fsub_rn(av, r2c); c2[2] = __fadd_rn(c2[2], __fmul_rn(t, t));
                r1a = r1b; r1b = r1c; r2a = r2b; r2b = r2c;
            }
        }
        // weights + warp reduction of one sample (rare path): d1/d2 = this thread's six distances at that sample
        auto emit = [&](const float *d1, const float *d2, int t, int slot, int pbuf) {
            float pw = 0.f, qw = 0.f; bool in = false;
#pragma unroll
            for (int m = 0; m < NLM_V; m++) {
                if (live[m]) {
                    in |= nlm_weight(d1[m], win[K + t + v0 + m], P, pw, qw);
                    in |= nlm_weight(d2[m], win[K + t + v0 + m + S + 1], P, pw, qw);
                }
            }
            if (!__any_sync(0xffffffffu, in)) return;
            for (int o = 16; o; o >>= 1) { pw += __shfl_xor_sync(0xffffffffu, pw, o); qw += __shfl_xor_sync(0xffffffffu, qw, o); }
            if (lane == 0) { part[pbuf][slot][warp][0] = pw; part[pbuf][slot][warp][1] = qw; }
        };
        // sample t = 0 is a chunk of its own, then chunks of 32 samples starting at t = 1
        int pb = 0;
        for (int c0 = 0; c0 < H; c0 = (c0 == 0) ? 1 : c0 + NLM_CHUNK, pb ^= 1) {
            const int cn = c0 == 0 ? 1 : min(NLM_CHUNK, H - c0);
            part[pb][lane][warp][0] = 0.f; part[pb][lane][warp][1] = 0.f;
            __syncwarp();
            if (c0 == 0) {
#pragma unroll
                for (int m = 0; m < NLM_V; m++) { c1[m] = fmaxf(c1[m], 0.f); c2[m] = fmaxf(c2[m], 0.f); }
                emit(c1, c2, 0, 0, pb);
            } else {
                // old / new patch edge of lag v0+m at sample t: p0[t+m] / p1[t+m]; of lag v0+m+S+1: p2[t+m] / p3[t+m]
                const float *p0 = win + v0 - 1, *p1 = p0 + K2 + 1, *p2 = p0 + S + 1, *p3 = p2 + K2 + 1;
                const float *u0 = win + S - 1, *u1 = u0 + K2 + 1;
                int ci = 0;
                for (; ci + NLM_G <= cn; ci += NLM_G) {
                    const int t0 = c0 + ci;
                    float e0[NLM_G + 2], e1[NLM_G + 2], e2[NLM_G + 2], e3[NLM_G + 2];
#pragma unroll
                    for (int j = 0; j < NLM_G + 2; j++) { e0[j] = p0[t0 + j]; e1[j] = p1[t0 + j]; e2[j] = p2[t0 + j]; e3[j] = p3[t0 + j]; }
                    float d1[NLM_G][NLM_V], d2[NLM_G][NLM_V];
                    float lo = 3.0e38f;
#pragma unroll
                    for (int g = 0; g < NLM_G; g++) {
                        const float ao = u0[t0 + g], an = u1[t0 + g];
#pragma unroll
                        for (int m = 0; m < NLM_V; m++) {
                            const float a1 = __fsub_rn(ao, e0[g + m]), b1 = __fsub_rn(an, e1[g + m]);
                            const float a2 = __fsub_rn(ao, e2[g + m]), b2 = __fsub_rn(an, e3[g + m]);
                            c1[m] = fmaxf(__fadd_rn(c1[m], __fadd_rn(-__fmul_rn(a1, a1), __fmul_rn(b1, b1))), 0.f);
                            c2[m] = fmaxf(__fadd_rn(c2[m], __fadd_rn(-__fmul_rn(a2, a2), __fmul_rn(b2, b2))), 0.f);
                            d1[g][m] = c1[m]; d2[g][m] = c2[m];
                            if (live[m]) lo = fminf(lo, fminf(c1[m], c2[m]));
                        }
                    }
                    // cheap superset of (d * sw < smooth) over the trip; the exact test follows in emit()
                    if (__any_sync(0xffffffffu, !(lo >= P.maybe))) {
#pragma unroll
                        for (int g = 0; g < NLM_G; g++) emit(d1[g], d2[g], t0 + g, ci + g, pb);
                    }
                }
                for (; ci < cn; ci++) {
                    const int t = c0 + ci;
                    const float ao = u0[t], an = u1[t];
#pragma unroll
                    for (int m = 0; m < NLM_V; m++) {
                        const float a1 = __fsub_rn(ao, p0[t + m]), b1 = __fsub_rn(an, p1[t + m]);
                        const float a2 = __fsub_rn(ao, p2[t + m]), b2 = __fsub_rn(an, p3[t + m]);
                        c1[m] = fmaxf(__fadd_rn(c1[m], __fadd_rn(-__fmul_rn(a1, a1), __fmul_rn(b1, b1))), 0.f);
                        c2[m] = fmaxf(__fadd_rn(c2[m], __fadd_rn(-__fmul_rn(a2, a2), __fmul_rn(b2, b2))), 0.f);
                    }
                    emit(c1, c2, t, ci, pb);
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
    const int threads = std::max(32, (((S + NLM_V - 1) / NLM_V + 31) / 32) * 32);
    if (threads > 128) JT_THROW(JT_ERR_UNSUPPORTED, "anlmdn research radius %d samples (max 384)", S);
    Sig o = in; o.d = jt_dalloc<float>(c, in.n);
    if (in.n <= 0) return o;
    const float m = (float)smooth_m, a = (float)strength;
    const float lut_scale = 1.f / m * (float)(1 << 20);
    const float sw = (65536.f / (4 * K + 2)) / sqrtf(a);
    const float smooth = fminf(m, (float)(1 << 20) / lut_scale);
    const int64_t n_hops = (in.n + H - 1) / H;
    const size_t smem = sizeof(float) * (N + 128);
    jt_smem_optin((const void *)k_anlmdn, (size_t)(std::max<size_t>(smem, 1024)));
    const int grid = jt_grid_for(n_hops, 1, c->num_sms, 64);
    JtLaunch L(c, "anlmdn");
    NlmK P; P.sw = sw; P.smooth = smooth; P.lut_scale = lut_scale; P.inv_lut_scale = 1.f / lut_scale;
    P.maybe = (smooth / sw) * 1.0001f;
    k_anlmdn<<<grid, threads, smem, c->stream>>>((const float *)in.d, (float *)o.d, in.n, K, S, P, n_hops);
    return o;
}
