// anlmdn -- non-local means denoiser (libavfilter/af_anlmdn.c), float32:
// "anlmdn=s=0.00001:p=0.0060:r=0.0020:m=3" (reference: filters.go:95-100, 811-816;
// scripts/anlmdn-matrix-spike.sh:292-320).  K = p*fs patch radius, S = r*fs search radius,
// hops of H = 2K+1 samples, 2S lags per output sample.  The dominant Pass-2 cost.
//
// One CTA per hop, one thread per PAIR of lags (j = i-S+v and j = i+1+v share the two patch-edge
// samples of i).  The hop's window (H + 2(K+S) floats) is staged in shared memory; every thread
// keeps its lags' running patch distances in registers and updates them with the same unfused
// float operations, in the same order, as the scalar C (full SSD at the hop's first sample,
// then add-new/subtract-old), so distances are bit-identical to a sequential run.  Weighted
// sums are reduced per warp with shuffles (skipped when no lane of the warp is inside the
// smoothing cut-off) and combined across warps once per 32 samples.
#include "jt_internal.h"
#include "jt_device.cuh"

#define NLM_CHUNK 32
#define NLM_MAXWARPS 16

__global__ void __launch_bounds__(512)
k_anlmdn(const float *__restrict__ x, float *__restrict__ y, int64_t n, int K, int S, float sw, float smooth,
         float lut_scale, float inv_lut_scale, int64_t n_hops)
{
    extern __shared__ float win[];                       // N floats
    __shared__ float part[NLM_CHUNK][NLM_MAXWARPS][2];
    const int H = 2 * K + 1, N = H + 2 * (K + S), offset = N - H;
    const int v = threadIdx.x, lane = v & 31, warp = v >> 5, nwarp = blockDim.x >> 5;
    const bool active = v < S;
    const float *f = win + K;
    const int dj1 = -S + v, dj2 = 1 + v;                 // lag v and lag v+S of the scalar loop
    for (int64_t h = blockIdx.x; h < n_hops; h += gridDim.x) {
        const int64_t pos = h * (int64_t)H;
        const int nb = (int)min((int64_t)H, n - pos);
        __syncthreads();
        for (int j = threadIdx.x; j < N; j += blockDim.x) {
            const int64_t s = pos - offset + j;
            win[j] = (s >= 0 && s < pos + nb) ? x[s] : 0.f;
        }
        __syncthreads();
        float c1 = 0.f, c2 = 0.f;
        if (active) {
            const float *fi = f + S, *f1 = fi + dj1, *f2 = fi + dj2;
            float d1 = 0.f, d2 = 0.f;
            for (int k = -K; k <= K; k++) {
                const float a = fi[k];
                const float t1 = __fsub_rn(a, f1[k]), t2 = __fsub_rn(a, f2[k]);
                d1 = __fadd_rn(d1, __fmul_rn(t1, t1)); d2 = __fadd_rn(d2, __fmul_rn(t2, t2));
            }
            c1 = d1; c2 = d2;
        }
        for (int c0 = 0; c0 < H; c0 += NLM_CHUNK) {
            const int cn = min(NLM_CHUNK, H - c0);
            const float *fi = f + S + c0;
            for (int ci = 0; ci < cn; ci++, fi++) {
                float pw = 0.f, qw = 0.f; bool in = false;
                if (active) {
                    const float *f1 = fi + dj1, *f2 = fi + dj2;
                    if (c0 + ci != 0) {
                        const float ao = fi[-K - 1], an = fi[K];
                        const float a1 = __fsub_rn(ao, f1[-K - 1]), b1 = __fsub_rn(an, f1[K]);
                        const float a2 = __fsub_rn(ao, f2[-K - 1]), b2 = __fsub_rn(an, f2[K]);
                        c1 = __fadd_rn(c1, __fadd_rn(-__fmul_rn(a1, a1), __fmul_rn(b1, b1)));
                        c2 = __fadd_rn(c2, __fadd_rn(-__fmul_rn(a2, a2), __fmul_rn(b2, b2)));
                    }
                    if (c1 < 0.f) c1 = 0.f;
                    if (c2 < 0.f) c2 = 0.f;
                    const float w1 = __fmul_rn(c1, sw), w2 = __fmul_rn(c2, sw);
                    if (!(w1 >= smooth)) {
                        const unsigned idx = (unsigned)__fmul_rn(w1, lut_scale);
                        const float w = __expf(-(float)idx * inv_lut_scale);
                        pw = __fmul_rn(w, f1[0]); qw = w; in = true;
                    }
                    if (!(w2 >= smooth)) {
                        const unsigned idx = (unsigned)__fmul_rn(w2, lut_scale);
                        const float w = __expf(-(float)idx * inv_lut_scale);
                        pw += __fmul_rn(w, f2[0]); qw += w; in = true;
                    }
                }
                if (__any_sync(0xffffffffu, in)) {
                    for (int o = 16; o; o >>= 1) { pw += __shfl_xor_sync(0xffffffffu, pw, o); qw += __shfl_xor_sync(0xffffffffu, qw, o); }
                } else { pw = 0.f; qw = 0.f; }
                if (lane == 0) { part[ci][warp][0] = pw; part[ci][warp][1] = qw; }
            }
            __syncthreads();
            if (threadIdx.x < cn) {
                const int ci = threadIdx.x, t = c0 + ci;
                float P = 0.f, Q = 0.f;
                for (int w = 0; w < nwarp; w++) { P += part[ci][w][0]; Q += part[ci][w][1]; }
                P = __fadd_rn(P, f[S + t]); Q = __fadd_rn(Q, 1.f);
                if (t < nb) y[pos + t] = __fdiv_rn(P, Q);
            }
            __syncthreads();
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
    const size_t smem = sizeof(float) * N;
    JT_CUDA(cudaFuncSetAttribute(k_anlmdn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem, 1024)));
    const int grid = jt_grid_for(n_hops, 1, c->num_sms, 64);
    JtLaunch L(c, "anlmdn");
    k_anlmdn<<<grid, threads, smem, c->stream>>>((const float *)in.d, (float *)o.d, in.n, K, S, sw, smooth, lut_scale, 1.f / lut_scale, n_hops);
    return o;
}
