// adeclick (libavfilter/af_adeclick.c), f64, overlap-save method:
// "adeclick=t=1.7:w=55:o=50:m=s" (reference: filters.go:947-962, 513-521).
// Per window: AR model (autocorrelation + Levinson-Durbin), prediction-error click detector
// with burst fusion, least-squares interpolation of the flagged samples (LDL^T solve).
// Every window depends on the input only, so each phase gets the parallelism that suits it:
//   A  autocorrelation      CTA per window, one thread per lag, window staged in shared memory
//   B  Levinson-Durbin      one thread per window (a 48-step dependent recursion, 131k windows/hour)
//   C  detector + fusion    CTA per window, one thread per sample, click bitmap
//   E  interpolation        warp per window: right-looking banded LDL^T in a 49x49 shared-memory ring
//                           with the forward substitution fused, then the back substitution
// Every sum runs in the scalar code's order with unfused multiply/add, so results are bit-identical
// to a sequential run (tests assert equality).  The interpolation matrix is banded in click order
// (two clicks more than ar_order samples apart do not couple, and LDL^T fill-in provably stays inside
// the band): the right-looking update applies to element (j,i) exactly the subtractions
// x -= (D_k*L_ik)*L_jk, k ascending, that the scalar left-looking loop applies.
#include "jt_internal.h"
#include "jt_device.cuh"

struct DcConst { int W, hop, skip, order, burst, bw, nwords; double threshold; };

__device__ __forceinline__ double jdmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double jdadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double jdsub(double a, double b) { return __dsub_rn(a, b); }

// sample j of window w: the fifo is `skip` zeros followed by the stream; past the fifo's end
// av_audio_fifo_peek leaves the previous peek's samples in place
__device__ __forceinline__ double dc_sample(const double *__restrict__ x, int64_t n, int64_t w, int j, const DcConst &K)
{
    const int64_t fifo_len = (int64_t)K.skip + n;
    int64_t ws = w;
    if (w * (int64_t)K.hop + j >= fifo_len) ws = (fifo_len - 1 - j) >= 0 ? (fifo_len - 1 - j) / K.hop : -1;
    if (ws < 0) return 0.0;
    const int64_t s = ws * (int64_t)K.hop + j - K.skip;
    return (s >= 0 && s < n) ? x[s] : 0.0;
}

// ---- A: autocorrelation -------------------------------------------------------------------
__global__ void __launch_bounds__(64)
k_dc_autocorr(const double *__restrict__ x, int64_t n, int64_t n_windows, DcConst K, double *__restrict__ r_out)
{
    extern __shared__ double s_in[];
    for (int64_t w = blockIdx.x; w < n_windows; w += gridDim.x) {
        __syncthreads();
        for (int j = threadIdx.x; j < K.W; j += blockDim.x) s_in[j] = dc_sample(x, n, w, j, K);
        __syncthreads();
        const int l = threadIdx.x;
        if (l <= K.order) {
            double v = 0.0;
            for (int j = l; j < K.W; j++) v = jdadd(v, jdmul(s_in[j], s_in[j - l]));
            r_out[w * K.bw + l] = jdmul(v, 1. / K.W);
        }
    }
}

// ---- B: Levinson-Durbin (af_adeclick.c autoregression()) -----------------------------------
#define DC_MAXORDER 128
__global__ void __launch_bounds__(128)
k_dc_levinson(const double *__restrict__ r_in, int64_t n_windows, DcConst K, double *__restrict__ acoef, double *__restrict__ sigmae)
{
    const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_windows) return;
    double r[DC_MAXORDER + 1], a[DC_MAXORDER + 1], k[DC_MAXORDER + 1];
    const int order = K.order;
    for (int i = 0; i <= order; i++) { r[i] = r_in[w * K.bw + i]; a[i] = 0.0; }
    k[0] = a[0] = -r[1] / r[0];
    double alpha = jdmul(r[0], jdsub(1., jdmul(k[0], k[0])));
    for (int i = 1; i < order; i++) {
        double eps = 0.;
        for (int j = 0; j < i; j++) eps = jdadd(eps, jdmul(a[j], r[i - j]));
        eps = jdadd(eps, r[i + 1]);
        k[i] = -eps / alpha;
        alpha = jdmul(alpha, jdsub(1., jdmul(k[i], k[i])));
        for (int j = i - 1; j >= 0; j--) k[j] = jdadd(a[j], jdmul(k[i], a[i - j - 1]));
        for (int j = 0; j <= i; j++) a[j] = k[j];
    }
    k[0] = 1.;
    for (int i = 1; i <= order; i++) k[i] = a[i - 1];
    bool finite = true;
    for (int i = 0; i <= order; i++) { acoef[w * K.bw + i] = k[i]; finite = finite && isfinite(k[i]); }
    sigmae[2 * w] = sqrt(alpha);
    sigmae[2 * w + 1] = finite ? 1.0 : 0.0;
}

// ---- C: detector, burst fusion, edge clearing -> click bitmap + count ------------------------
__global__ void __launch_bounds__(256)
k_dc_detect(const double *__restrict__ x, int64_t n, int64_t n_windows, DcConst K, const double *__restrict__ acoef,
            const double *__restrict__ sigmae, unsigned *__restrict__ bits_out, int *__restrict__ count_out)
{
    extern __shared__ double s_in[];                 // W samples, then order+1 coefficients
    __shared__ unsigned s_bits[160], s_fill[160];
    __shared__ int s_cnt;
    double *kc = s_in + K.W;
    for (int64_t w = blockIdx.x; w < n_windows; w += gridDim.x) {
        __syncthreads();
        const bool finite = sigmae[2 * w + 1] != 0.0;
        if (!finite) {          // acoefficients not finite: the window passes through untouched
            for (int j = threadIdx.x; j < K.nwords; j += blockDim.x) bits_out[w * K.nwords + j] = 0u;
            if (threadIdx.x == 0) count_out[w] = 0;
            continue;
        }
        for (int j = threadIdx.x; j < K.W; j += blockDim.x) s_in[j] = dc_sample(x, n, w, j, K);
        for (int j = threadIdx.x; j <= K.order; j += blockDim.x) kc[j] = acoef[w * K.bw + j];
        if (threadIdx.x == 0) s_cnt = 0;
        __syncthreads();
        const double thr = jdmul(sigmae[2 * w], K.threshold);
        for (int i0 = 0; i0 < K.nwords * 32; i0 += blockDim.x) {
            const int i = i0 + threadIdx.x; bool cflag = false;
            if (i < K.W) {
                double d = 0.0;
                if (i >= K.order) for (int j = 0; j <= K.order; j++) d = jdadd(d, jdmul(kc[j], s_in[i - j]));
                cflag = fabs(d) > thr;
            }
            const unsigned m = __ballot_sync(0xffffffffu, cflag);
            if ((threadIdx.x & 31) == 0 && (i >> 5) < K.nwords) s_bits[i >> 5] = m;
        }
        __syncthreads();
        // burst fusion: a gap between two consecutive original clicks at most `burst` apart is filled
        for (int wd = threadIdx.x; wd < K.nwords; wd += blockDim.x) {
            unsigned add = 0;
            for (int b = 0; b < 32; b++) {
                const int i = wd * 32 + b; if (i >= K.W) break;
                if ((s_bits[wd] >> b) & 1u) continue;
                int p = -1, q = -1;
                for (int d = 1; d <= K.burst && p < 0; d++) { const int j = i - d; if (j >= 0 && ((s_bits[j >> 5] >> (j & 31)) & 1u)) p = j; }
                if (p < 0) continue;
                for (int d = 1; d <= K.burst && q < 0; d++) { const int j = i + d; if (j < K.W && ((s_bits[j >> 5] >> (j & 31)) & 1u)) q = j; }
                if (q >= 0 && q - p <= K.burst) add |= 1u << b;
            }
            s_fill[wd] = add;
        }
        __syncthreads();
        for (int wd = threadIdx.x; wd < K.nwords; wd += blockDim.x) {
            unsigned m = s_bits[wd] | s_fill[wd];
            for (int b = 0; b < 32; b++) { const int i = wd * 32 + b; if (i < K.order || i >= K.W - K.order) m &= ~(1u << b); }
            bits_out[w * K.nwords + wd] = m;
            if (m) atomicAdd(&s_cnt, __popc(m));
        }
        __syncthreads();
        if (threadIdx.x == 0) count_out[w] = s_cnt;
    }
}

// ---- E: interpolation ----------------------------------------------------------------------
#define DC_WARPS 5          // 20.7 KB of ring per warp: two 5-warp CTAs fill an SM's 227 KB
__global__ void __launch_bounds__(DC_WARPS * 32)
k_dc_interp(const double *__restrict__ x, double *__restrict__ y, int64_t n, int64_t n_windows, DcConst K,
            const double *__restrict__ acoef, const unsigned *__restrict__ bits_in, const int *__restrict__ count_in,
            double *__restrict__ scratch, size_t scratch_per_warp, int *__restrict__ next_window)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int order = K.order, bw = K.bw;                    // ring is bw x bw
    __shared__ unsigned char s_pa[2048], s_pb[2048];   // pair table (a, b), b <= a <= 63, ordered by a then b
    const int npairs_full = order * (order + 1) / 2;
    for (int p = threadIdx.x; p < npairs_full; p += blockDim.x) {
        int a = 1; while (a * (a + 1) / 2 <= p) a++;
        s_pa[p] = (unsigned char)a; s_pb[p] = (unsigned char)(p - a * (a - 1) / 2 + 1);
    }
    __syncthreads();
    const size_t per_warp = ((size_t)bw * bw + 3 * (size_t)bw + (size_t)((K.nwords + 1) / 2)) * sizeof(double);
    double *ring = (double *)(smem_raw + (size_t)warp * per_warp);
    double *kc = ring + (size_t)bw * bw, *aux = kc + bw, *vring = aux + bw;     // vring: right-hand side / forward-substituted values of rows k .. k+order
    unsigned *sbits = (unsigned *)(vring + bw);                                  // the window's click bitmap
    const int64_t gw = (int64_t)blockIdx.x * DC_WARPS + warp;
    double *S = scratch + (size_t)gw * scratch_per_warp;
    // scratch layout: Lg[nmax][order], Dg[nmax], vec[nmax], outv[nmax], idx (as int)[nmax]
    const size_t nmax = (size_t)K.W;
    double *Lg = S, *yd = Lg + nmax * order, *vec = yd + nmax, *outv = vec + nmax;      // yd[i] = y_i / D_i
    int *idx = (int *)(outv + nmax), *am = idx + 2 * ((nmax + 1) / 2);
    // ring slot of matrix index k + a given rk = k % bw (a <= order < bw): no integer division in the inner loops
#define WRAP(v) ((v) >= bw ? (v) - bw : (v))
#define RINGS(r, cidx) ring[(size_t)(r) * bw + (cidx)]

    for (;;) {
        int64_t w = 0;
        if (lane == 0) w = atomicAdd(next_window, 1);
        w = __shfl_sync(0xffffffffu, w, 0);
        if (w >= n_windows) break;
        const int nclk = count_in[w];
        if (nclk <= 0) continue;
        __syncwarp();
        // ordered index list from the bitmap
        int basec = 0;
        for (int w0 = 0; w0 < K.nwords; w0 += 32) {
            const int wd = w0 + lane;
            const unsigned m = wd < K.nwords ? bits_in[w * K.nwords + wd] : 0u;
            if (wd < K.nwords) sbits[wd] = m;
            const int cnt = __popc(m);
            int incl = cnt;
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
            int pos = basec + incl - cnt;
            unsigned mm = m;
            while (mm) { const int b = __ffs(mm) - 1; mm &= mm - 1; idx[pos++] = wd * 32 + b; }
            basec += __shfl_sync(0xffffffffu, incl, 31);
        }
        for (int l = lane; l <= order; l += 32) kc[l] = acoef[w * bw + l];
        __syncwarp();
        // aux = autocorrelation of the AR coefficients (scale 1)
        for (int l = lane; l <= order; l += 32) {
            double v = 0.0;
            for (int j = l; j <= order; j++) v = jdadd(v, jdmul(kc[j], kc[j - l]));
            aux[l] = jdmul(v, 1.);
        }
        __syncwarp();
        // right-hand side: known neighbours of every click, j = -order .. order (sample p = ii - j descending).
        // Interior windows read the stream directly; only the first / last windows need the fifo edge rules.
        {
            const int64_t f0 = w * (int64_t)K.hop, s_base = f0 - K.skip;
            const bool interior = f0 + K.W <= (int64_t)K.skip + n && s_base >= 0 && s_base + K.W <= n;
            const double *xw = x + s_base;
            for (int i = lane; i < nclk; i += 32) {
                const int ii = idx[i];
                double value = 0.;
                // a flagged neighbour contributes a +0.0 product (value - 0.0 == value bit for bit), so the products
                // of four taps are formed ahead of the subtraction chain instead of one multiply per dependent step
#pragma unroll 4
                for (int j = -order; j <= order; j++) {
                    const int p = ii - j;
                    const bool known = !((sbits[p >> 5] >> (p & 31)) & 1u);
                    const double xs = interior ? xw[p] : dc_sample(x, n, w, p, K);
                    const double pr = known ? jdmul(xs, aux[j < 0 ? -j : j]) : 0.0;
                    value = jdsub(value, pr);
                }
                vec[i] = value;
            }
        }
        __syncwarp();
        // row j enters the ring (slot rj): A(j,i) = aux[idx_j - idx_i] for the earlier clicks i it couples to
        // (at most `order` samples away); slots outside that profile are never read.  Its right-hand side enters vring.
        auto load_row = [&](int j, int rj) {
            if (j >= nclk) return;
            const int ij = idx[j];
            for (int t0 = 0; t0 < bw; t0 += 32) {
                const int t = t0 + lane, i = j - t;
                const int d = (t < bw && i >= 0) ? ij - idx[i] : order + 1;
                if (d <= order) RINGS(rj, WRAP(rj + bw - t)) = aux[d];
                if (!__ballot_sync(0xffffffffu, d <= order && lane == 31)) break;      // sorted: nothing further couples
            }
        };
        for (int j = 0; j <= order; j++) load_row(j, j);
        for (int j = lane; j <= order && j < nclk; j += 32) vring[j] = vec[j];
        __syncwarp();
        bool ok = true;
        int rk = 0;
        for (int k = 0; k < nclk; k++) {
            const double Dk = RINGS(rk, rk);
            if (Dk == 0.) { ok = false; break; }
            // rows coupled to click k: later clicks at most `order` samples away.  Everything below them in
            // column k is an exact zero of the factor (two clicks further apart never couple, fill-in stays
            // inside this profile), and x - 0*y == x bit for bit, so those updates are skipped.
            int amax;
            {
                const int ik = idx[k], j0 = k + 1 + lane, j1 = j0 + 32;
                const bool c0 = j0 < nclk && idx[j0] - ik <= order, c1 = j1 < nclk && j1 - k <= order && idx[j1] - ik <= order;
                amax = __popc(__ballot_sync(0xffffffffu, c0)) + __popc(__ballot_sync(0xffffffffu, c1));
            }
            const double yk = vring[rk];
            const double v_in = (lane == 0 && k + bw < nclk) ? vec[k + bw] : 0.0;      // right-hand side of the row entering below: fetched early
            // column k: L(j,k) = A'(j,k) / D_k ; forward substitution v_j -= L(j,k) * y_k
            for (int a = lane + 1; a <= amax; a += 32) {
                const int ra = WRAP(rk + a);
                const double L = RINGS(ra, rk) / Dk;
                RINGS(ra, rk) = L;
                Lg[(size_t)k * order + (a - 1)] = L;
                vring[ra] = jdsub(vring[ra], jdmul(L, yk));
            }
            if (lane == 0) { yd[k] = yk; outv[k] = Dk; am[k] = amax; }      // y_k / D_k is formed after the loop, off the critical path
            __syncwarp();
            // trailing update: A'(k+a, k+b) -= (D_k * L(k+b,k)) * L(k+a,k), 1 <= b <= a <= amax
            const int npairs = amax * (amax + 1) / 2;
            for (int p = lane; p < npairs; p += 32) {
                const int ra = WRAP(rk + s_pa[p]), rb = WRAP(rk + s_pb[p]);
                const double Lb = RINGS(rb, rk), La = RINGS(ra, rk);
                RINGS(ra, rb) = jdsub(RINGS(ra, rb), jdmul(jdmul(Dk, Lb), La));
            }
            load_row(k + bw, rk);           // row k+order+1 takes the slot row k leaves (no trailing update touches it)
            if (lane == 0 && k + bw < nclk) vring[rk] = v_in;
            __syncwarp();
            rk = rk + 1 == bw ? 0 : rk + 1;
        }
        if (!ok) continue;                  // factorisation hit a zero pivot: af_adeclick.c leaves the window as is
        __syncwarp();
        for (int i = lane; i < nclk; i += 32) yd[i] = yd[i] / outv[i];
        __syncwarp();
        // back substitution: out_i = y_i / D_i - sum_{j>i} L(j,i) * out_j, j ascending.  Lane a keeps out_{i+1+a}
        // (and out_{i+33+a}) in registers; moving to i-1 shifts them up by one lane.
        {
            double win0 = 0.0, win1 = 0.0;
            int amax = am[nclk - 1];
            double o_y = yd[nclk - 1];
            double L0 = lane < amax ? Lg[(size_t)(nclk - 1) * order + lane] : 0.0;
            double L1 = lane + 32 < amax ? Lg[(size_t)(nclk - 1) * order + lane + 32] : 0.0;
            for (int i = nclk - 1; i >= 0; i--) {
                // next row's operands are fetched before this row's dependent chain starts
                int amax_n = 0; double oy_n = 0.0, L0n = 0.0, L1n = 0.0;
                if (i > 0) {
                    amax_n = am[i - 1]; oy_n = yd[i - 1];
                    if (lane < amax_n) L0n = Lg[(size_t)(i - 1) * order + lane];
                    if (lane + 32 < amax_n) L1n = Lg[(size_t)(i - 1) * order + lane + 32];
                }
                const double prod0 = jdmul(L0, win0), prod1 = jdmul(L1, win1);
                double o = o_y;
                for (int a = 0; a < amax; a++) {
                    const double pr = __shfl_sync(0xffffffffu, a < 32 ? prod0 : prod1, a & 31);
                    o = jdsub(o, pr);
                }
                if (lane == 0) outv[i] = o;
                const double carry = __shfl_sync(0xffffffffu, win0, 31);
                win1 = __shfl_up_sync(0xffffffffu, win1, 1); if (lane == 0) win1 = carry;
                win0 = __shfl_up_sync(0xffffffffu, win0, 1); if (lane == 0) win0 = o;
                amax = amax_n; o_y = oy_n; L0 = L0n; L1 = L1n;
            }
        }
        __syncwarp();
        // only the hop-sized middle of the window is emitted (overlap-save)
        for (int i = lane; i < nclk; i += 32) {
            const int p = idx[i] - K.skip;
            if (p >= 0 && p < K.hop) { const int64_t q = w * (int64_t)K.hop + p; if (q < n) y[q] = outv[i]; }
        }
        __syncwarp();
    }
#undef RINGS
#undef WRAP
}

Sig jt_adeclick(jt_ctx *c, const Sig &in, double w_ms, double overlap_pct, double ar_pct, double threshold, double burst_pct, int method_save)
{
    if (in.fmt != JT_FMT_DBL) JT_THROW(JT_ERR_INVALID_ARG, "adeclick expects f64 input");
    if (!method_save) JT_THROW(JT_ERR_UNSUPPORTED, "adeclick overlap-add method (m=a)");
    DcConst K;
    K.W = (int)(in.rate * w_ms / 1000.);
    if (K.W < 100) JT_THROW(JT_ERR_INVALID_ARG, "adeclick window too small");
    if (K.W > 5120) JT_THROW(JT_ERR_UNSUPPORTED, "adeclick window of %d samples (max 5120)", K.W);
    K.order = std::max((int)(K.W * ar_pct / 100.), 1);
    if (K.order > 63) JT_THROW(JT_ERR_UNSUPPORTED, "adeclick AR order %d (max 63)", K.order);
    K.burst = (int)(K.W * burst_pct / 1000.);
    K.hop = (int)(K.W * (1. - (overlap_pct / 100.)));
    if (K.hop < 1) JT_THROW(JT_ERR_INVALID_ARG, "adeclick overlap too large");
    K.skip = (K.W - K.hop) / 2; K.bw = K.order + 1; K.threshold = threshold; K.nwords = (K.W + 31) / 32;
    Sig o = in; o.d = jt_dalloc<double>(c, in.n);
    if (in.n <= 0) return o;
    const int64_t nw = (in.n + K.hop - 1) / K.hop;
    double *d_r = jt_dalloc<double>(c, (size_t)nw * K.bw), *d_a = jt_dalloc<double>(c, (size_t)nw * K.bw), *d_sig = jt_dalloc<double>(c, (size_t)nw * 2);
    unsigned *d_bits = jt_dalloc<unsigned>(c, (size_t)nw * K.nwords);
    int *d_cnt = jt_dalloc<int>(c, nw + 1);
    int *d_next = d_cnt + nw;
    JT_CUDA(cudaMemsetAsync(d_next, 0, sizeof(int), c->stream));
    // the windows pass through except where clicks are repaired: out = in, then E overwrites
    JT_CUDA(cudaMemcpyAsync(o.d, in.d, sizeof(double) * (size_t)in.n, cudaMemcpyDeviceToDevice, c->stream));
    const size_t smemA = sizeof(double) * K.W, smemC = sizeof(double) * (K.W + K.bw);
    jt_smem_optin((const void *)k_dc_autocorr, (size_t)(smemA));
    jt_smem_optin((const void *)k_dc_detect, (size_t)(smemC));
    const size_t per_warp = ((size_t)K.bw * K.bw + 3 * (size_t)K.bw + (size_t)((K.nwords + 1) / 2)) * sizeof(double);
    const size_t smemE = per_warp * DC_WARPS;
    jt_smem_optin((const void *)k_dc_interp, (size_t)(smemE));
    int per_sm = (int)std::max<size_t>(1, std::min<size_t>(8, (227 * 1024) / (smemE + 4096 + 1024)));
    int gridE = (int)std::min<int64_t>((nw + DC_WARPS - 1) / DC_WARPS, (int64_t)c->num_sms * per_sm);
    if (gridE < 1) gridE = 1;
    const size_t nmax = (size_t)K.W;
    const size_t scratch_per_warp = nmax * K.order + 3 * nmax + 2 * ((nmax + 1) / 2) + 8;
    double *scratch = jt_dalloc<double>(c, scratch_per_warp * (size_t)gridE * DC_WARPS);
    { JtLaunch L(c, "adeclick:autocorr");
    k_dc_autocorr<<<jt_grid_for(nw, 1, c->num_sms, 64), 64, smemA, c->stream>>>((const double *)in.d, in.n, nw, K, d_r); }
    { JtLaunch L(c, "adeclick:levinson");
    k_dc_levinson<<<(int)((nw + 127) / 128), 128, 0, c->stream>>>(d_r, nw, K, d_a, d_sig); }
    { JtLaunch L(c, "adeclick:detect");
    k_dc_detect<<<jt_grid_for(nw, 1, c->num_sms, 32), 256, smemC, c->stream>>>((const double *)in.d, in.n, nw, K, d_a, d_sig, d_bits, d_cnt); }
    JtLaunch L(c, "adeclick:interp");
    k_dc_interp<<<gridE, DC_WARPS * 32, smemE, c->stream>>>((const double *)in.d, (double *)o.d, in.n, nw, K, d_a, d_bits, d_cnt,
                                                            scratch, scratch_per_warp, d_next);
    return o;
}
