// adeclick (libavfilter/af_adeclick.c), f64, overlap-save method:
// "adeclick=t=1.7:w=55:o=50:m=s" (reference: filters.go:947-962, 513-521).
// Per window: AR model (autocorrelation + Levinson-Durbin), prediction-error click detector
// with burst fusion, least-squares interpolation of the flagged samples (LDL^T solve).
// Every window depends on the input only -> one WARP per window, windows fully parallel.
// All sums run in the scalar code's order with unfused multiply/add, so the result is
// bit-identical to a sequential run.  The interpolation matrix is banded in click order
// (entries vanish when two clicks are more than ar_order samples apart, and the LDL^T
// fill-in provably stays inside the band), so the factorisation touches only the band:
// same arithmetic, O(n*order^2) instead of O(n^3).
#include "jt_internal.h"
#include "jt_device.cuh"

#define DC_WARPS 4

struct DcConst { int W, hop, skip, order, burst, bw; double threshold; };

__device__ __forceinline__ double jdmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double jdadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double jdsub(double a, double b) { return __dsub_rn(a, b); }

__global__ void __launch_bounds__(DC_WARPS * 32)
k_adeclick(const double *__restrict__ x, double *__restrict__ y, int64_t n, int64_t n_windows, DcConst K,
           double *__restrict__ scratch /* per warp: matrix n*bw + vector + yv + out */, size_t scratch_per_warp)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int W = K.W, order = K.order, bw = K.bw;
    const int nwords = (W + 31) / 32;
    const size_t per_warp = (size_t)W * 8 + (size_t)nwords * 4 + (size_t)W * 2 + (size_t)4 * (order + 1) * 8 + 16;
    unsigned char *base = smem_raw + (size_t)warp * ((per_warp + 15) & ~(size_t)15);
    double *in = (double *)base;
    double *r = in + W, *a = r + (order + 1), *kc = a + (order + 1), *aux = kc + (order + 1);
    unsigned *bits = (unsigned *)(aux + (order + 1));
    unsigned short *index = (unsigned short *)(bits + nwords);
    const int64_t gw = (int64_t)blockIdx.x * DC_WARPS + warp, gstride = (int64_t)gridDim.x * DC_WARPS;
    double *M = scratch + (size_t)gw * scratch_per_warp;       // persistent warps: gw < total warps launched
    const int64_t fifo_len = (int64_t)K.skip + n;

    for (int64_t w = gw; w < n_windows; w += gstride) {
        const int64_t fpos = w * (int64_t)K.hop;
        __syncwarp();
        // window content; past the fifo's end av_audio_fifo_peek leaves the previous peek's samples
        for (int j = lane; j < W; j += 32) {
            int64_t ws = w;
            if (fpos + j >= fifo_len) ws = (fifo_len - 1 - j) >= 0 ? (fifo_len - 1 - j) / K.hop : -1;
            double v = 0.0;
            if (ws >= 0) { const int64_t s = ws * (int64_t)K.hop + j - K.skip; if (s >= 0 && s < n) v = x[s]; }
            in[j] = v;
        }
        for (int j = lane; j < nwords; j += 32) bits[j] = 0u;
        __syncwarp();
        // autocorrelation: one lane per lag, scalar order
        for (int l0 = 0; l0 <= order; l0 += 32) {
            const int l = l0 + lane;
            if (l <= order) {
                double v = 0.0;
                for (int j = l; j < W; j++) v = jdadd(v, jdmul(in[j], in[j - l]));
                r[l] = jdmul(v, 1. / W);
            }
        }
        __syncwarp();
        double sigmae = 0.0;
        if (lane == 0) {        // Levinson-Durbin (af_adeclick.c autoregression())
            for (int i = 0; i < order; i++) a[i] = 0.0;
            kc[0] = a[0] = -r[1] / r[0];
            double alpha = jdmul(r[0], jdsub(1., jdmul(kc[0], kc[0])));
            for (int i = 1; i < order; i++) {
                double eps = 0.;
                for (int j = 0; j < i; j++) eps = jdadd(eps, jdmul(a[j], r[i - j]));
                eps = jdadd(eps, r[i + 1]);
                kc[i] = -eps / alpha;
                alpha = jdmul(alpha, jdsub(1., jdmul(kc[i], kc[i])));
                for (int j = i - 1; j >= 0; j--) kc[j] = jdadd(a[j], jdmul(kc[i], a[i - j - 1]));
                for (int j = 0; j <= i; j++) a[j] = kc[j];
            }
            kc[0] = 1.;
            for (int i = 1; i <= order; i++) kc[i] = a[i - 1];
            sigmae = sqrt(alpha);
        }
        sigmae = __shfl_sync(0xffffffffu, sigmae, 0);
        __syncwarp();
        bool finite = true;
        for (int i = lane; i <= order; i += 32) finite = finite && isfinite(kc[i]);
        finite = __all_sync(0xffffffffu, finite);
        int nclk = 0;
        if (finite) {
            // detection + threshold
            const double thr = jdmul(sigmae, K.threshold);
            for (int i0 = 0; i0 < W; i0 += 32) {
                const int i = i0 + lane; bool c = false;
                if (i < W && i >= order) {
                    double d = 0.0;
                    for (int j = 0; j <= order; j++) d = jdadd(d, jdmul(kc[j], in[i - j]));
                    c = fabs(d) > thr;
                } else if (i < W) c = fabs(0.0) > thr;
                const unsigned m = __ballot_sync(0xffffffffu, c);
                if (lane == 0) bits[i0 >> 5] = m;
            }
            __syncwarp();
            // burst fusion: fill gaps between consecutive original clicks at most `burst` apart
            unsigned fill[4] = {0, 0, 0, 0};       // this lane's words (W <= 4096)
            for (int wd = lane, q = 0; wd < nwords; wd += 32, q++) {
                unsigned add = 0;
                for (int b = 0; b < 32; b++) {
                    const int i = wd * 32 + b; if (i >= W) break;
                    if ((bits[wd] >> b) & 1u) continue;
                    int p = -1, qn = -1;
                    for (int d = 1; d <= K.burst && p < 0; d++) { const int j = i - d; if (j >= 0 && ((bits[j >> 5] >> (j & 31)) & 1u)) p = j; }
                    if (p < 0) continue;
                    for (int d = 1; d <= K.burst && qn < 0; d++) { const int j = i + d; if (j < W && ((bits[j >> 5] >> (j & 31)) & 1u)) qn = j; }
                    if (qn >= 0 && qn - p <= K.burst) add |= 1u << b;
                }
                if (q < 4) fill[q] = add;
            }
            __syncwarp();
            for (int wd = lane, q = 0; wd < nwords; wd += 32, q++) if (q < 4) bits[wd] |= fill[q];
            __syncwarp();
            // clear the edges, compact the ordered index list
            for (int wd = lane; wd < nwords; wd += 32) {
                unsigned m = bits[wd];
                for (int b = 0; b < 32; b++) { const int i = wd * 32 + b; if (i < order || i >= W - order) m &= ~(1u << b); }
                bits[wd] = m;
            }
            __syncwarp();
            int basec = 0;
            for (int w0 = 0; w0 < nwords; w0 += 32) {
                const int wd = w0 + lane;
                const unsigned m = wd < nwords ? bits[wd] : 0u;
                const int cnt = __popc(m);
                int incl = cnt;
                for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
                int pos = basec + incl - cnt;
                unsigned mm = m;
                while (mm) { const int b = __ffs(mm) - 1; mm &= mm - 1; index[pos++] = (unsigned short)(wd * 32 + b); }
                basec += __shfl_sync(0xffffffffu, incl, 31);
            }
            nclk = basec;
            __syncwarp();
        }
        if (finite && nclk > 0) {
            double *vec = M + (size_t)nclk * bw, *yv = vec + nclk, *outv = yv + nclk;
            // aux = autocorrelation of the AR coefficients
            for (int l0 = 0; l0 <= order; l0 += 32) {
                const int l = l0 + lane;
                if (l <= order) { double v = 0.0; for (int j = l; j <= order; j++) v = jdadd(v, jdmul(kc[j], kc[j - l])); aux[l] = jdmul(v, 1.); }
            }
            __syncwarp();
            // banded matrix rows + right-hand side
            for (int i = lane; i < nclk; i += 32) {
                const int ii = index[i];
                for (int cidx = 0; cidx < bw; cidx++) {
                    const int k = i - (bw - 1) + cidx;
                    double v = 0.0;
                    if (k >= 0) { const int d = ii - (int)index[k]; if (d <= order) v = aux[d]; }
                    M[(size_t)i * bw + cidx] = v;
                }
                double value = 0.;
                for (int j = -order; j <= order; j++) {
                    const int p = ii - j;
                    if (!((bits[p >> 5] >> (p & 31)) & 1u)) value = jdsub(value, jdmul(in[p], aux[j < 0 ? -j : j]));
                }
                vec[i] = value;
            }
            __syncwarp();
            // LDL^T inside the band.  entry (j,k) lives at M[j*bw + (k - j + bw - 1)]
#define MAT(j, k) M[(size_t)(j) * bw + ((k) - (j) + bw - 1)]
            bool ok = true;
            for (int i = 0; i < nclk && ok; i++) {
                double value = 0.0;
                if (lane == 0) {
                    value = MAT(i, i);
                    for (int j = max(0, i - (bw - 1)); j < i; j++) value = jdsub(value, jdmul(jdmul(MAT(j, j), MAT(i, j)), MAT(i, j)));
                    MAT(i, i) = value;
                }
                value = __shfl_sync(0xffffffffu, value, 0);
                if (value == 0.) { ok = false; break; }
                __syncwarp();
                for (int j0 = i + 1; j0 < min(nclk, i + bw); j0 += 32) {
                    const int j = j0 + lane;
                    if (j < min(nclk, i + bw)) {
                        double xv = MAT(j, i);
                        for (int k = max(0, j - (bw - 1)); k < i; k++) xv = jdsub(xv, jdmul(jdmul(MAT(k, k), MAT(i, k)), MAT(j, k)));
                        MAT(j, i) = xv / value;
                    }
                }
                __syncwarp();
            }
            if (ok) {
                if (lane == 0) {
                    for (int i = 0; i < nclk; i++) {
                        double value = vec[i];
                        for (int j = max(0, i - (bw - 1)); j < i; j++) value = jdsub(value, jdmul(MAT(i, j), yv[j]));
                        yv[i] = value;
                    }
                    for (int i = nclk - 1; i >= 0; i--) {
                        double o = yv[i] / MAT(i, i);
                        for (int j = i + 1; j < min(nclk, i + bw); j++) o = jdsub(o, jdmul(MAT(j, i), outv[j]));
                        outv[i] = o;
                    }
                }
                __syncwarp();
                for (int i = lane; i < nclk; i += 32) in[index[i]] = outv[i];
                __syncwarp();
            }
#undef MAT
        }
        // overlap-save: this window contributes hop samples starting at `skip`
        for (int j = lane; j < K.hop; j += 32) {
            const int64_t q = w * (int64_t)K.hop + j;
            if (q < n) y[q] = in[K.skip + j];
        }
    }
}

Sig jt_adeclick(jt_ctx *c, const Sig &in, double w_ms, double overlap_pct, double ar_pct, double threshold, double burst_pct, int method_save)
{
    if (in.fmt != JT_FMT_DBL) JT_THROW(JT_ERR_INVALID_ARG, "adeclick expects f64 input");
    if (!method_save) JT_THROW(JT_ERR_UNSUPPORTED, "adeclick overlap-add method (m=a)");
    DcConst K;
    K.W = (int)(in.rate * w_ms / 1000.);
    if (K.W < 100) JT_THROW(JT_ERR_INVALID_ARG, "adeclick window too small");
    if (K.W > 4096) JT_THROW(JT_ERR_UNSUPPORTED, "adeclick window of %d samples (max 4096)", K.W);
    K.order = std::max((int)(K.W * ar_pct / 100.), 1);
    K.burst = (int)(K.W * burst_pct / 1000.);
    K.hop = (int)(K.W * (1. - (overlap_pct / 100.)));
    if (K.hop < 1) JT_THROW(JT_ERR_INVALID_ARG, "adeclick overlap too large");
    K.skip = (K.W - K.hop) / 2; K.bw = K.order + 1; K.threshold = threshold;
    Sig o = in; o.d = jt_dalloc<double>(c, in.n);
    if (in.n <= 0) return o;
    const int64_t n_windows = (in.n + K.hop - 1) / K.hop;
    const int nwords = (K.W + 31) / 32;
    size_t per_warp = (size_t)K.W * 8 + (size_t)nwords * 4 + (size_t)K.W * 2 + (size_t)4 * (K.order + 1) * 8 + 16;
    per_warp = (per_warp + 15) & ~(size_t)15;
    const size_t smem = per_warp * DC_WARPS;
    if (smem > 220 * 1024) JT_THROW(JT_ERR_UNSUPPORTED, "adeclick window needs %zu bytes of shared memory", smem);
    JT_CUDA(cudaFuncSetAttribute(k_adeclick, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = (int)std::max<size_t>(1, std::min<size_t>(4, (220 * 1024) / smem));
    int grid = (int)std::min<int64_t>((n_windows + DC_WARPS - 1) / DC_WARPS, (int64_t)c->num_sms * per_sm);
    if (grid < 1) grid = 1;
    const size_t nmax = (size_t)K.W;
    const size_t scratch_per_warp = nmax * K.bw + 3 * nmax;
    double *scratch = jt_dalloc<double>(c, scratch_per_warp * (size_t)grid * DC_WARPS);
    JtLaunch L(c, "adeclick");
    k_adeclick<<<grid, DC_WARPS * 32, smem, c->stream>>>((const double *)in.d, (double *)o.d, in.n, n_windows, K, scratch, scratch_per_warp);
    return o;
}
