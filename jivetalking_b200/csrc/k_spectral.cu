// aspectralstats (libavfilter/af_aspectralstats.c): win 2048, hann, hop 1024, 13 statistics per
// hop in float32 ("aspectralstats=win_size=2048:win_func=hann:measure=all", reference:
// filters.go:625; formulas docs/Spectral-Metrics-Reference.md:9-33).
// One CTA per hop: the 2048-point window is staged in shared memory, transformed by an
// in-place radix-2 FFT (twiddles from a read-only table), the 1024 magnitudes are reduced
// with warp shuffles.  Flux needs the previous hop's magnitudes, so magnitudes go to HBM
// once and a second tiny kernel forms the differences.
#include "jt_internal.h"
#include "jt_device.cuh"
#include "jt_fft.cuh"
#include <cfloat>
#include <algorithm>
#include <cstring>

#define SP_THREADS 256

template <int K> __device__ __forceinline__ void block_sum(float (&v)[K], float (*red)[K])
{
#pragma unroll
    for (int k = 0; k < K; k++) v[k] = jt_warp_sum(v[k]);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < K; k++) red[w][k] = v[k];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < K; k++) {
        float s = 0;
        for (int i = 0; i < SP_THREADS / 32; i++) s += red[i][k];
        v[k] = s;
    }
}

// Two items (hops) per CTA trip: hop A is the real part and hop B the imaginary part of ONE complex Stockham
// radix-4 transform (jt_fft.cuh); the two real spectra are separated by symmetry and reduced one after the other.
__global__ void __launch_bounds__(SP_THREADS)
k_spectral(const float *__restrict__ x, int64_t n, int win, int rate, int64_t n_items, const int64_t *__restrict__ list,
           const float2 *__restrict__ tw_g, const float *__restrict__ lut,
           float *__restrict__ mags, float *__restrict__ rows)
{
    extern __shared__ float2 sbuf[];                 // bufA[win], bufB[win], tw[win], smag[win/2 floats]
    __shared__ float red[SP_THREADS / 32][8];
    __shared__ float s_scan[SP_THREADS / 32];
    __shared__ int s_idx;
    float2 *bufA = sbuf, *bufB = sbuf + win, *tw = sbuf + 2 * win;
    float *smag = (float *)(sbuf + 3 * win);
    const int hop = win / 2, size = win / 2;
    for (int i = threadIdx.x; i < win; i += SP_THREADS) tw[i] = tw_g[i];
    const int64_t n_pairs = (n_items + 1) / 2;
    for (int64_t pr = blockIdx.x; pr < n_pairs; pr += gridDim.x) {
        const int64_t itA = 2 * pr, itB = itA + 1;
        const bool hasB = itB < n_items;
        const int64_t hA = list ? list[itA] : itA, hB = hasB ? (list ? list[itB] : itB) : 0;
        const int64_t wA = (hA - 1) * (int64_t)hop, wB = (hB - 1) * (int64_t)hop;      // window start samples
        __syncthreads();
        for (int i = threadIdx.x; i < win; i += SP_THREADS) {
            const float l = lut[i];
            const int64_t sa = wA + i, sb = wB + i;
            const float va = (sa >= 0 && sa < n) ? __fmul_rn(x[sa], l) : 0.f;
            const float vb = (hasB && sb >= 0 && sb < n) ? __fmul_rn(x[sb], l) : 0.f;
            bufA[i] = make_float2(va, vb);
        }
        __syncthreads();
        const float2 *Z = af_fft<false>(bufA, bufB, tw, win);
        const float wscale = 1.f / win;
        const int per = size / SP_THREADS;           // 4 for win 2048
        float mA[4], mB[4];
        for (int j = 0; j < per; j++) {
            const int k = threadIdx.x * per + j;
            const float2 zk = Z[k], zn = Z[(win - k) & (win - 1)];
            const float2 fa = make_float2(0.5f * (zk.x + zn.x), 0.5f * (zk.y - zn.y));
            const float2 fb = make_float2(0.5f * (zk.y + zn.y), 0.5f * (zn.x - zk.x));
            mA[j] = hypotf(__fmul_rn(fa.x, wscale), __fmul_rn(fa.y, wscale));
            mB[j] = hypotf(__fmul_rn(fb.x, wscale), __fmul_rn(fb.y, wscale));
        }
        for (int u = 0; u < 2; u++) {
            if (u == 1 && !hasB) break;
            const int64_t it = u ? itB : itA;
            float m[4];
            for (int j = 0; j < per; j++) m[j] = u ? mB[j] : mA[j];
            __syncthreads();
            for (int j = 0; j < per; j++) {
                smag[threadIdx.x * per + j] = m[j];
                mags[it * (int64_t)size + threadIdx.x * per + j] = m[j];
            }
            __syncthreads();
            const float mag0 = smag[0];
            const float scale = (rate / 2) / (float)size;
            const float mean_freq = size * 0.5f;
            // pass 1
            float r1[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            float mx = 0.f;
            for (int j = 0; j < per; j++) {
                const int nn = threadIdx.x * per + j; const float v = m[j];
                r1[0] += v;                                   // sum
                r1[1] += v * nn * scale;                      // centroid num
                r1[2] += v * logf(v + FLT_EPSILON);           // entropy num
                const float ve = FLT_EPSILON + v;
                r1[3] += logf(ve);                            // flatness log-sum
                r1[4] += ve;                                  // flatness den
                if (nn >= 1) { r1[5] += (v - mag0) / nn; r1[6] += v; }   // decrease
                const float q = (nn - mean_freq) / mean_freq; r1[7] += q * q;   // slope den
                mx = fmaxf(mx, v);
            }
            block_sum<8>(r1, red);
            mx = jt_warp_max(mx);
            __syncthreads();
            if ((threadIdx.x & 31) == 0) s_scan[threadIdx.x >> 5] = mx;
            __syncthreads();
            mx = 0.f; for (int i = 0; i < SP_THREADS / 32; i++) mx = fmaxf(mx, s_scan[i]);
            const float sum = r1[0], mean = sum / size;
            const float centroid = sum <= FLT_EPSILON ? 1.f : r1[1] / sum;
            // pass 2
            float r2[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            for (int j = 0; j < per; j++) {
                const int nn = threadIdx.x * per + j; const float v = m[j];
                const float dm = v - mean; r2[0] += dm * dm;                 // variance
                const float df = nn * scale - centroid, df2 = df * df;
                r2[1] += v * df2; r2[2] += v * (df2 * df); r2[3] += v * (df2 * df2);
                r2[4] += ((nn - mean_freq) / mean_freq) * dm;               // slope num
            }
            block_sum<8>(r2, red);
            // rolloff: first bin where the running sum reaches 85 % of the total
            float run = m[0]; float loc[4]; loc[0] = run;
            for (int j = 1; j < per; j++) { run += m[j]; loc[j] = run; }
            float incl = run;                                  // warp inclusive scan of per-thread totals
            const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
            for (int o = 1; o < 32; o <<= 1) { float t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
            __syncthreads();
            if (lane == 31) s_scan[wid] = incl;
            if (threadIdx.x == 0) s_idx = 0x7fffffff;
            __syncthreads();
            float woff = 0; for (int i = 0; i < wid; i++) woff += s_scan[i];
            const float excl = woff + incl - run;
            const float norm = sum * 0.85f;
            int first = 0x7fffffff;
            for (int j = per - 1; j >= 0; j--) if (excl + loc[j] >= norm) first = threadIdx.x * per + j;
            if (first != 0x7fffffff) atomicMin(&s_idx, first);
            __syncthreads();
            if (threadIdx.x == 0) {
                float *r = rows + it * JT_SP_COUNT;
                const float spread = sum <= FLT_EPSILON ? 1.f : sqrtf(r2[1] / sum);
                r[JT_SP_mean] = mean;
                r[JT_SP_variance] = r2[0] / size;
                r[JT_SP_centroid] = centroid;
                r[JT_SP_spread] = spread;
                float den = sum * (spread * spread * spread);
                r[JT_SP_skewness] = den <= FLT_EPSILON ? 1.f : r2[2] / den;
                den = sum * ((spread * spread) * (spread * spread));
                r[JT_SP_kurtosis] = den <= FLT_EPSILON ? 1.f : r2[3] / den;
                den = logf((float)size);
                r[JT_SP_entropy] = den <= FLT_EPSILON ? 1.f : -r1[2] / den;
                { float num = expf(r1[3] / size), d2 = r1[4] / size; r[JT_SP_flatness] = d2 <= FLT_EPSILON ? 0.f : num / d2; }
                r[JT_SP_crest] = mean <= FLT_EPSILON ? 0.f : mx / mean;
                r[JT_SP_flux] = 0.f;                          // second kernel
                r[JT_SP_slope] = fabsf(r1[7]) <= FLT_EPSILON ? 0.f : r2[4] / r1[7];
                r[JT_SP_decrease] = r1[6] <= FLT_EPSILON ? 0.f : r1[5] / r1[6];
                r[JT_SP_rolloff] = scale * (s_idx == 0x7fffffff ? 0 : s_idx);
            }
        }
    }
}

__global__ void __launch_bounds__(256)
k_spectral_flux(const float *__restrict__ mags, int size, int64_t n_items, const int64_t *__restrict__ prev, float *__restrict__ rows)
{
    __shared__ float sw[8];
    for (int64_t h = blockIdx.x; h < n_items; h += gridDim.x) {
        const int64_t pb = prev ? prev[h] : h - 1;      // item holding the previous hop's magnitudes (-1: none, zeros)
        float s = 0;
        for (int i = threadIdx.x; i < size; i += 256) {
            const float a = mags[h * size + i], b = pb >= 0 ? mags[pb * size + i] : 0.f;
            const float d = a - b; s += d * d;
        }
        s = jt_warp_sum(s);
        __syncthreads();
        if ((threadIdx.x & 31) == 0) sw[threadIdx.x >> 5] = s;
        __syncthreads();
        if (threadIdx.x == 0) { float t = 0; for (int i = 0; i < 8; i++) t += sw[i]; rows[h * JT_SP_COUNT + JT_SP_flux] = sqrtf(t); }
    }
}

void jt_aspectralstats_launch(jt_ctx *c, const Sig &in0, int win, const std::vector<int64_t> *wanted, SpectralPending &pd)
{
    pd = SpectralPending();
    Sig in = jt_convert(c, in0, JT_FMT_FLT);
    const int hop = win / 2;
    pd.n_hops = (in.n + hop - 1) / hop;
    pd.sparse = wanted != nullptr;
    if (pd.n_hops <= 0) return;
    if (win < 1024 || (win & (win - 1)) || (win / 2) % SP_THREADS || win / 2 / SP_THREADS > 4)
        JT_THROW(JT_ERR_UNSUPPORTED, "aspectralstats win_size %d", win);
    // Only hops a sink frame will ever show are computed (the reference sees one hop per 100 ms frame,
    // SURVEY 7.3), each with its predecessor for the flux term.
    std::vector<int64_t> prev;
    if (wanted) {
        std::vector<int64_t> w;
        for (int64_t h : *wanted) if (h >= 0 && h < pd.n_hops) { w.push_back(h); if (h > 0) w.push_back(h - 1); }
        std::sort(w.begin(), w.end());
        w.erase(std::unique(w.begin(), w.end()), w.end());
        pd.items.swap(w);
        prev.resize(pd.items.size());
        for (size_t i = 0; i < pd.items.size(); i++) prev[i] = (i > 0 && pd.items[i - 1] == pd.items[i] - 1) ? (int64_t)i - 1 : -1;
        if (pd.items.empty()) return;
    }
    const int64_t n_items = wanted ? (int64_t)pd.items.size() : pd.n_hops;
    pd.n_items = n_items;
    std::vector<float2> tw(win); std::vector<float> lut(win);
    for (int k = 0; k < win; k++) { double a = -2.0 * M_PI * k / win; tw[k] = make_float2((float)cos(a), (float)sin(a)); }
    for (int i = 0; i < win; i++) lut[i] = (float)(.5 * (1 - cos(2 * M_PI * i / (win - 1))));
    const float2 *d_tw = jt_dev_table(c, "spectral_tw", tw); const float *d_lut = jt_dev_table(c, "spectral_hann", lut);
    int64_t *d_items = nullptr, *d_prev = nullptr;
    if (wanted) {
        // the two index lists go through pinned staging so that the copy is truly asynchronous
        int64_t *h_idx = jt_pinned<int64_t>(c, 2 * (size_t)n_items);
        memcpy(h_idx, pd.items.data(), sizeof(int64_t) * n_items);
        memcpy(h_idx + n_items, prev.data(), sizeof(int64_t) * n_items);
        d_items = jt_dalloc<int64_t>(c, 2 * (size_t)n_items); d_prev = d_items + n_items;
        jt_copy_small(c, d_items, h_idx, sizeof(int64_t) * 2 * n_items);
    }
    float *d_mags = jt_dalloc<float>(c, (size_t)n_items * (win / 2));
    float *d_rows = jt_dalloc<float>(c, (size_t)n_items * JT_SP_COUNT);
    const int grid = jt_grid_for((n_items + 1) / 2, 1, c->num_sms, 8);
    const size_t smem_sp = sizeof(float2) * 3 * (size_t)win + sizeof(float) * (win / 2);
    jt_smem_optin((const void *)k_spectral, (size_t)(smem_sp));
    {
        JtLaunch L(c, "aspectralstats", 2);
        k_spectral<<<grid, SP_THREADS, smem_sp, c->stream>>>((const float *)in.d, in.n, win, in.rate, n_items, d_items, d_tw, d_lut, d_mags, d_rows);
        k_spectral_flux<<<grid, 256, 0, c->stream>>>(d_mags, win / 2, n_items, d_prev, d_rows);
    }
    pd.h_rows = jt_pinned<float>(c, (size_t)n_items * JT_SP_COUNT);
    jt_copy_small(c, pd.h_rows, d_rows, sizeof(float) * (size_t)n_items * JT_SP_COUNT);
    pd.ev = jt_record_event(c);
}

void jt_aspectralstats_finish(jt_ctx *c, SpectralPending &pd, std::vector<float> &rows, int64_t &n_hops)
{
    n_hops = pd.n_hops;
    rows.assign((size_t)std::max<int64_t>(n_hops, 0) * JT_SP_COUNT, 0.f);
    if (n_hops <= 0 || pd.n_items <= 0 || !pd.h_rows) return;
    JT_CUDA(cudaEventSynchronize(pd.ev));
    if (!pd.sparse) memcpy(rows.data(), pd.h_rows, sizeof(float) * rows.size());
    else for (int64_t i = 0; i < pd.n_items; i++) memcpy(&rows[(size_t)pd.items[i] * JT_SP_COUNT], &pd.h_rows[(size_t)i * JT_SP_COUNT], sizeof(float) * JT_SP_COUNT);
}

void jt_aspectralstats(jt_ctx *c, const Sig &in0, int win, std::vector<float> &rows, int64_t &n_hops, const std::vector<int64_t> *wanted)
{
    SpectralPending pd; jt_aspectralstats_launch(c, in0, win, wanted, pd); jt_aspectralstats_finish(c, pd, rows, n_hops);
}
