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
k_anlmdn(const float *__restrict__ x, float *__restrict__ y, int64_t n, int K, int S, NlmK P, int64_t n_hops,
         const int *__restrict__ hop_list, const int *__restrict__ hop_count)
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
    // with a hop list (k_nlm_screen) only the listed hops are walked; every other hop is already the delayed input
    const int64_t n_work = hop_list ? (int64_t)*hop_count : n_hops;
    for (int64_t hi = blockIdx.x; hi < n_work; hi += gridDim.x) {
        const int64_t h = hop_list ? (int64_t)hop_list[hi] : hi;
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


// ---- screening pass ------------------------------------------------------------------------------------------------
// A lag only enters a sample's average when its patch distance d = sum_{k=-K..K} (f[i+k] - f[j+k])^2 falls below
// smooth / sw (with the reference's s = 0.00001 that is a mean squared difference at -65 dBFS); when no lag does, the
// output is the (delayed) input, bit for bit: P / Q = (0 + f[i]) / (0 + 1).  On programme material almost every hop is of
// that kind, and the exact kernel above spends its time proving it with 192 running distances per sample.  The screen proves
// it for a whole group of hops with an eighth of the instructions, from a LOWER BOUND of every distance the exact kernel
// can see there:
//   * g_v[n] = (f[n] - f[n+v])^2, v = 1..S, covers both signs of the lag (the pair i, i-v is the pair i-v, (i-v)+v);
//   * every patch (H = 2K+1 consecutive n) contains R = floor((H - BS + 1) / BS) whole blocks of BS samples of a grid anchored
//     at the group's first patch start, so d >= the sum of R consecutive block sums G_v[b]; and it lies inside R + 3 blocks,
//     which bounds d -- and with it the rounding the exact kernel's float accumulator can have picked up since the hop's
//     seed (3 H ulp of its largest value; 5 H is allowed for) -- from above;
//   * a group is passed through when, for every lag, min_b LB_v[b] >= 1.01 smooth / sw + crel * max_b UB_v[b], crel covering
//     that rounding and the rounding of the block prefix sums the two bounds are read from.
// Anything else (digital silence, room tone under the cut-off, a decay of more than ~33 dB inside a group) goes to the exact
// kernel through the hop list.  Outputs are identical to running the exact kernel on every hop.
struct NlmScreen { int K, S, H, BS, R, HC, SP, wlen; float thr, crel; };

template <int LPT>
__global__ void __launch_bounds__(256)
k_nlm_screen(const float *__restrict__ x, float *__restrict__ y, int64_t n, NlmScreen Q, int64_t n_hops,
             int *__restrict__ hop_list, int *__restrict__ hop_count)
{
    extern __shared__ __align__(16) float nlm_sm[];
    float *xs = nlm_sm;                 // the group's samples from its first patch start on, zero outside the stream
    float *Gs = nlm_sm + Q.wlen;        // [block][lag] block sums, then their prefix sums in place
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const int H = Q.H, BS = Q.BS, SP = Q.SP, R = Q.R;
    const int64_t n_groups = (n_hops + Q.HC - 1) / Q.HC;
    for (int64_t hg = blockIdx.x; hg < n_groups; hg += gridDim.x) {
        const int64_t h0 = hg * Q.HC;
        const int hn = (int)min((int64_t)Q.HC, n_hops - h0);
        const int64_t pos0 = h0 * (int64_t)H, ps_min = pos0 - 2 * (Q.K + Q.S);
        const int span = (hn - 1) * H + Q.S + H - 1;              // patch starts of the group: ps_min .. ps_min + span
        const int nrs = (span + BS - 1) / BS + 1, nblk = nrs + R + 2;
        const int wl = nblk * BS + SP + 8;
        __syncthreads();
        for (int j = threadIdx.x; j < wl; j += blockDim.x) {
            const int64_t s = ps_min + j;
            xs[j] = (s >= 0 && s < n) ? x[s] : 0.f;
        }
        __syncthreads();
        // block sums: a warp per block, lane l owns the lags 1 + LPT l .. LPT (l + 1)
        for (int b = warp; b < nblk; b += nwarp) {
            const float *pa = xs + b * BS, *pw = pa + 1 + LPT * lane;
            float acc[LPT];
#pragma unroll
            for (int m = 0; m < LPT; m++) acc[m] = 0.f;
            for (int i = 0; i < BS; i += 8) {
                const float4 a0 = *(const float4 *)(pa + i), a1 = *(const float4 *)(pa + i + 4);
                const float a[8] = { a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w };
                float w[8 + LPT - 1];
#pragma unroll
                for (int j = 0; j < 8 + LPT - 1; j++) w[j] = pw[i + j];
#pragma unroll
                for (int j = 0; j < 8; j++)
#pragma unroll
                    for (int m = 0; m < LPT; m++) { const float t = a[j] - w[j + m]; acc[m] = fmaf(t, t, acc[m]); }
            }
#pragma unroll
            for (int m = 0; m < LPT; m++) Gs[b * SP + LPT * lane + m] = acc[m];
        }
        __syncthreads();
        // per lag: prefix sums over the blocks, then the two bounds of every run start
        int flag = 0;
        for (int v = threadIdx.x; v < Q.S; v += blockDim.x) {
            float *col = Gs + v;
            float p = 0.f;
            for (int b = 0; b < nblk; b++) { const float g = col[b * SP]; col[b * SP] = p; p += g; }
            col[nblk * SP] = p;
            float lo = 3.0e38f, hi = 0.f;
            for (int b = 0; b < nrs; b++) {
                const float p0 = col[b * SP];
                lo = fminf(lo, col[(b + R) * SP] - p0);
                hi = fmaxf(hi, col[(b + R + 3) * SP] - p0);
            }
            flag |= !(lo >= 1.01f * Q.thr + Q.crel * hi);
        }
        flag = __syncthreads_or(flag);
        // the pass-through value of every sample of the group (the exact kernel overwrites the listed hops afterwards)
        const int cnt = (int)min((int64_t)hn * H, n - pos0);
        for (int t = threadIdx.x; t < cnt; t += blockDim.x) y[pos0 + t] = __fadd_rn(0.f, xs[t + Q.K + Q.S]);
        if (flag && threadIdx.x == 0) {
            const int base = atomicAdd(hop_count, hn);
            for (int i = 0; i < hn; i++) hop_list[base + i] = (int)(h0 + i);
        }
    }
}

template <int LPT>
static void nlm_screen_launch(jt_ctx *c, const float *x, float *y, int64_t n, const NlmScreen &Q, int64_t n_hops, size_t smem, int *list, int *count)
{
    jt_smem_optin((const void *)k_nlm_screen<LPT>, smem);
    const int64_t n_groups = (n_hops + Q.HC - 1) / Q.HC;
    k_nlm_screen<LPT><<<jt_grid_for(n_groups, 1, c->num_sms, 64), 256, smem, c->stream>>>(x, y, n, Q, n_hops, list, count);
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
    NlmK P; P.sw = sw; P.smooth = smooth; P.lut_scale = lut_scale; P.inv_lut_scale = 1.f / lut_scale;
    P.maybe = (smooth / sw) * 1.0001f;
    // screening pass (see k_nlm_screen): block size ~ H / 18 so that a patch holds R ~ 17 whole blocks, hops grouped so that a
    // group's block sums fit ~70 KB of shared memory (three groups per SM)
    int *d_list = nullptr, *d_count = nullptr;
    const bool no_screen = getenv("JT_ANLMDN_NO_SCREEN") != nullptr;     // tests: every hop through the exact kernel
    NlmScreen Q{};
    Q.K = K; Q.S = S; Q.H = H; Q.BS = 32 * std::max(1, (H + 288) / 577); Q.R = (H - Q.BS + 1) / Q.BS;
    const int lpt_need = (S + 31) / 32;
    const int lpt = lpt_need <= 4 ? lpt_need : lpt_need <= 6 ? 6 : lpt_need <= 8 ? 8 : 12;
    Q.SP = 32 * lpt;
    size_t smem_screen = 0;
    if (!no_screen && Q.R >= 4 && n_hops < (int64_t)1 << 31) {
        for (Q.HC = 8; Q.HC >= 1; Q.HC--) {
            const int span = (Q.HC - 1) * H + S + H - 1;
            const int nrs = (span + Q.BS - 1) / Q.BS + 1, nblk = nrs + Q.R + 2;
            Q.wlen = ((nblk * Q.BS + Q.SP + 8) + 3) & ~3;
            smem_screen = sizeof(float) * ((size_t)Q.wlen + (size_t)(nblk + 1) * Q.SP);
            const double u = 1.0 / 16777216.0;
            Q.crel = (float)(1.5 * (5.0 * H * u + 2.0 * nblk * u * ((double)nblk / (Q.R + 3) + 1.0)));
            if (smem_screen <= 72 * 1024 || Q.HC == 1) break;
        }
        if (smem_screen > 200 * 1024) smem_screen = 0;
    }
    if (smem_screen) {
        Q.thr = smooth / sw;
        d_list = jt_dalloc<int>(c, (size_t)n_hops + 1);
        d_count = d_list + n_hops;
        JT_CUDA(cudaMemsetAsync(d_count, 0, sizeof(int), c->stream));
        JtLaunch L(c, "anlmdn:screen");
        switch (lpt) {
        case 1: nlm_screen_launch<1>(c, (const float *)in.d, (float *)o.d, in.n, Q, n_hops, smem_screen, d_list, d_count); break;
        case 2: nlm_screen_launch<2>(c, (const float *)in.d, (float *)o.d, in.n, Q, n_hops, smem_screen, d_list, d_count); break;
        case 3: nlm_screen_launch<3>(c, (const float *)in.d, (float *)o.d, in.n, Q, n_hops, smem_screen, d_list, d_count); break;
        case 4: nlm_screen_launch<4>(c, (const float *)in.d, (float *)o.d, in.n, Q, n_hops, smem_screen, d_list, d_count); break;
        case 6: nlm_screen_launch<6>(c, (const float *)in.d, (float *)o.d, in.n, Q, n_hops, smem_screen, d_list, d_count); break;
        case 8: nlm_screen_launch<8>(c, (const float *)in.d, (float *)o.d, in.n, Q, n_hops, smem_screen, d_list, d_count); break;
        default: nlm_screen_launch<12>(c, (const float *)in.d, (float *)o.d, in.n, Q, n_hops, smem_screen, d_list, d_count); break;
        }
    }
    JtLaunch L(c, "anlmdn");
    k_anlmdn<<<grid, threads, smem, c->stream>>>((const float *)in.d, (float *)o.d, in.n, K, S, P, n_hops, d_list, d_count);
    return o;
}
