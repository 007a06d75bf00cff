// anlmdn -- non-local means denoiser (libavfilter/af_anlmdn.c), float32:
// "anlmdn=s=0.00001:p=0.0060:r=0.0020:m=3" (reference: filters.go:95-100, 811-816;
// scripts/anlmdn-matrix-spike.sh:292-320).  K = p*fs patch radius, S = r*fs search radius,
// hops of H = 2K+1 samples, 2S lags per output sample.  The dominant Pass-2 cost.
//
// One CTA per hop, one thread per lag.  The hop's window (H + 2(K+S) floats) is staged in
// shared memory; every thread keeps its lag's running patch distance in a register and
// updates it with the same unfused float operations, in the same order, as the scalar C
// (full SSD at the hop's first sample, then add-new/subtract-old), so distances are
// bit-identical to a sequential run.  Weighted sums are reduced per warp with shuffles
// (skipped when no lane of the warp is inside the smoothing cut-off) and combined across
// warps once per 32 samples.
#include "jt_internal.h"
#include "jt_device.cuh"

#define NLM_CHUNK 32
#define NLM_MAXWARPS 32

__global__ void __launch_bounds__(1024)
k_anlmdn(const float *__restrict__ x, float *__restrict__ y, int64_t n, int K, int S, float sw, float smooth,
         float lut_scale, int64_t n_hops)
{
    extern __shared__ float win[];                       // N floats
    __shared__ float part[NLM_CHUNK][NLM_MAXWARPS][2];
    const int H = 2 * K + 1, N = H + 2 * (K + S), offset = N - H;
    const int v = threadIdx.x, lane = v & 31, warp = v >> 5, nwarp = blockDim.x >> 5;
    const bool active = v < 2 * S;
    const float *f = win + K;
    for (int64_t h = blockIdx.x; h < n_hops; h += gridDim.x) {
        const int64_t pos = h * (int64_t)H;
        const int nb = (int)min((int64_t)H, n - pos);
        __syncthreads();
        for (int j = threadIdx.x; j < N; j += blockDim.x) {
            const int64_t s = pos - offset + j;
            win[j] = (s >= 0 && s < pos + nb) ? x[s] : 0.f;
        }
        __syncthreads();
        // lag v: j = i - S + v + (v >= S)
        const int dj = active ? (-S + v + (v >= S ? 1 : 0)) : 0;
        float cache = 0.f;
        if (active) {
            const int i = S, j = i + dj;
            float d = 0.f;
            for (int k = -K; k <= K; k++) { const float t = __fsub_rn(f[i + k], f[j + k]); d = __fadd_rn(d, __fmul_rn(t, t)); }
            cache = d;
        }
        for (int c0 = 0; c0 < H; c0 += NLM_CHUNK) {
            const int cn = min(NLM_CHUNK, H - c0);
            for (int ci = 0; ci < cn; ci++) {
                const int i = S + c0 + ci, j = i + dj;
                float pw = 0.f, qw = 0.f; bool in = false;
                if (active) {
                    if (i != S) {
                        const float a = __fsub_rn(f[i - K - 1], f[j - K - 1]), b = __fsub_rn(f[i + K], f[j + K]);
                        cache = __fadd_rn(cache, __fadd_rn(-__fmul_rn(a, a), __fmul_rn(b, b)));
                    }
                    float distance = cache;
                    if (distance < 0.f) cache = distance = 0.f;
                    float w = __fmul_rn(distance, sw);
                    if (!(w >= smooth)) {
                        const unsigned idx = (unsigned)__fmul_rn(w, lut_scale);
                        w = expf(__fdiv_rn(-(float)idx, lut_scale));
                        pw = __fmul_rn(w, f[j]); qw = w; in = true;
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
    const int threads = ((2 * S + 31) / 32) * 32;
    if (threads > 1024) JT_THROW(JT_ERR_UNSUPPORTED, "anlmdn research radius %d samples (max 512)", S);
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
    k_anlmdn<<<grid, threads, smem, c->stream>>>((const float *)in.d, (float *)o.d, in.n, K, S, sw, smooth, lut_scale, n_hops);
    return o;
}
