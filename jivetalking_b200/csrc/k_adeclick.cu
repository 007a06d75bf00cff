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
// r[l] = (1/W) sum_{j=l}^{W-1} s[j] s[j-l], j ascending (the scalar loop's order), for l = 0 .. order.  A thread owns FOUR
// consecutive lags and walks j once: the operand s[j-l] of lag l at step j is the operand of lag l-1 at step j-1, so it
// moves through a four-deep register window and every step costs two shared-memory loads (s[j], s[j-l0]) for four
// multiply-adds -- the one-lag-per-thread version was bound by shared-memory bandwidth (two loads per multiply-add).
// Leading products with a not-yet-valid operand are +-0.0 added to a sum that is still +0.0: bit-identical.
#define DC_AC_G 4            // windows per CTA
#define DC_AC_LT 4           // lags per thread
__device__ __forceinline__ void dc_ac_lags(const double *__restrict__ sw, int PL, int l0, int W, double *v)
{
    // window sample i at plane i & 3, slot i >> 2 of sw.  j and l0 are multiples of 4 at the top of every trip, so a trip reads
    // slot q of each plane for s[j .. j+3] and slot q - l0/4 for s[j-l0 .. j-l0+3]
    double v0 = 0.0, v1 = 0.0, v2 = 0.0, v3 = 0.0;
    double c1 = 0.0, c2 = 0.0, c3 = 0.0;                 // s[j-l0-1], s[j-l0-2], s[j-l0-3] (zero before the window)
    const double *pb = sw + (l0 >> 2), *pa = sw;
    int j = l0;
    // four steps per trip: the eight loads and sixteen products of a trip are independent of the four running sums,
    // whose additions stay in j order
    for (; j + 4 <= W; j += 4, pb++, pa++) {
        const double b0 = pb[0], b1 = pb[PL], b2 = pb[2 * PL], b3 = pb[3 * PL];
        const double a0 = pa[0], a1 = pa[PL], a2 = pa[2 * PL], a3 = pa[3 * PL];
        const double p00 = jdmul(b0, a0), p01 = jdmul(b0, c1), p02 = jdmul(b0, c2), p03 = jdmul(b0, c3);
        const double p10 = jdmul(b1, a1), p11 = jdmul(b1, a0), p12 = jdmul(b1, c1), p13 = jdmul(b1, c2);
        const double p20 = jdmul(b2, a2), p21 = jdmul(b2, a1), p22 = jdmul(b2, a0), p23 = jdmul(b2, c1);
        const double p30 = jdmul(b3, a3), p31 = jdmul(b3, a2), p32 = jdmul(b3, a1), p33 = jdmul(b3, a0);
        v0 = jdadd(jdadd(jdadd(jdadd(v0, p00), p10), p20), p30);
        v1 = jdadd(jdadd(jdadd(jdadd(v1, p01), p11), p21), p31);
        v2 = jdadd(jdadd(jdadd(jdadd(v2, p02), p12), p22), p32);
        v3 = jdadd(jdadd(jdadd(jdadd(v3, p03), p13), p23), p33);
        c1 = a3; c2 = a2; c3 = a1;
    }
    for (; j < W; j++) {
        const double sj = sw[(j & 3) * PL + (j >> 2)], a0 = sw[((j - l0) & 3) * PL + ((j - l0) >> 2)];
        v0 = jdadd(v0, jdmul(sj, a0)); v1 = jdadd(v1, jdmul(sj, c1)); v2 = jdadd(v2, jdmul(sj, c2)); v3 = jdadd(v3, jdmul(sj, c3));
        c3 = c2; c2 = c1; c1 = a0;
    }
    v[0] = v0; v[1] = v1; v[2] = v2; v[3] = v3;
}

__global__ void __launch_bounds__(64)
k_dc_autocorr(const double *__restrict__ x, int64_t n, int64_t n_windows, DcConst K, int PL, double *__restrict__ r_out)
{
    // The DC_AC_G consecutive windows of a CTA overlap (hop < W): they are staged ONCE as the contiguous span they cover, in four
    // planes (sample 4k+p of the span at plane p, slot k).  Threads of a window read addresses 4 samples apart -- s[j - l0],
    // l0 = 4 * thread -- which in planes are consecutive slots; the s[j] every thread of a window needs is one broadcast.
    extern __shared__ double s_in[];
    const int tpw = (K.order + DC_AC_LT) / DC_AC_LT;      // threads per window: ceil((order + 1) / 4)
    const int64_t groups = (n_windows + DC_AC_G - 1) / DC_AC_G;
    const int g = threadIdx.x / tpw, l0 = (threadIdx.x - g * tpw) * DC_AC_LT;
    const double inv = 1. / K.W;
    for (int64_t grp = blockIdx.x; grp < groups; grp += gridDim.x) {
        const int64_t w0 = grp * DC_AC_G;
        const int nwin = (int)min((int64_t)DC_AC_G, n_windows - w0);
        const int64_t s_base = w0 * (int64_t)K.hop - K.skip;
        const int span = K.W + (nwin - 1) * K.hop;
        const bool shared_span = (K.hop & 3) == 0 && s_base >= 0 && s_base + span <= n;     // no fifo edge rule inside the group
        double v[4];
        if (shared_span) {
            __syncthreads();
#pragma unroll 8
            for (int i = threadIdx.x; i < span; i += 64) s_in[(i & 3) * PL + (i >> 2)] = x[s_base + i];
            __syncthreads();
            if (g < nwin) {
                dc_ac_lags(s_in + g * (K.hop >> 2), PL, l0, K.W, v);
                const int64_t w = w0 + g;
                for (int t = 0; t < 4; t++) if (l0 + t <= K.order) r_out[w * K.bw + l0 + t] = jdmul(v[t], inv);
            }
        } else {
            for (int gg = 0; gg < nwin; gg++) {           // stream edges: one window at a time through the fifo's rules
                __syncthreads();
                for (int j = threadIdx.x; j < K.W; j += 64) s_in[(j & 3) * PL + (j >> 2)] = dc_sample(x, n, w0 + gg, j, K);
                __syncthreads();
                if (g == gg) {
                    dc_ac_lags(s_in, PL, l0, K.W, v);
                    for (int t = 0; t < 4; t++) if (l0 + t <= K.order) r_out[(w0 + gg) * K.bw + l0 + t] = jdmul(v[t], inv);
                }
            }
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
// The interpolation's linear system is banded in click order; its band width is the largest number of later clicks within
// `order` samples of any click.  Windows are binned by that width so that k_dc_interp can run each bin with a ring just
// large enough (DC_CLASS_R): the narrow bins -- almost all windows -- fit three times as many warps on an SM.
#define DC_NCLASS 4
__host__ __device__ inline int dc_class_ring(int cl, int bw) { return cl == 0 ? 16 : cl == 1 ? 24 : cl == 2 ? 32 : bw; }

__global__ void __launch_bounds__(256)
k_dc_detect(const double *__restrict__ x, int64_t n, int64_t n_windows, DcConst K, const double *__restrict__ acoef,
            const double *__restrict__ sigmae, unsigned *__restrict__ bits_out, int *__restrict__ count_out,
            int *__restrict__ class_count /* DC_NCLASS */, int *__restrict__ class_list /* DC_NCLASS x n_windows */)
{
    extern __shared__ double s_raw[];                // the window (zero-padded: OP samples before, 4 after) in four planes, then the coefficients
    __shared__ unsigned s_bits[160], s_fill[160];
    __shared__ int s_cnt, s_maxc;
    // sample i (-OP <= i < W + 4) lives at plane (i + OP) & 3, slot (i + OP) >> 2: a thread's four samples are 4 apart from its
    // neighbour's, which in planes is the next slot -- no bank conflicts -- and the zero padding removes every bounds check
    const int OP = (K.order + 3) & ~3, PL = (OP + K.W + 4 + 3) / 4 + 1;
    double *kc = s_raw + 4 * PL;
#define SI(i) s_raw[(((i) + OP) & 3) * PL + (((i) + OP) >> 2)]
    for (int64_t w = blockIdx.x; w < n_windows; w += gridDim.x) {
        __syncthreads();
        const bool finite = sigmae[2 * w + 1] != 0.0;
        if (!finite) {          // acoefficients not finite: the window passes through untouched
            for (int j = threadIdx.x; j < K.nwords; j += blockDim.x) bits_out[w * K.nwords + j] = 0u;
            if (threadIdx.x == 0) count_out[w] = 0;
            continue;
        }
        for (int j = threadIdx.x; j < K.W; j += blockDim.x) SI(j) = dc_sample(x, n, w, j, K);
        for (int j = threadIdx.x; j < OP; j += blockDim.x) SI(j - OP) = 0.0;
        for (int j = threadIdx.x; j < 4; j += blockDim.x) SI(K.W + j) = 0.0;
        for (int j = threadIdx.x; j <= K.order + 3; j += blockDim.x) kc[j] = j <= K.order ? acoef[w * K.bw + j] : 0.0;
        if (threadIdx.x == 0) { s_cnt = 0; s_maxc = 0; }
        __syncthreads();
        const double thr = jdmul(sigmae[2 * w], K.threshold);
        // prediction error d_i = sum_{j=0}^{order} k[j] s[i-j], j ascending.  A thread owns four consecutive samples: the
        // operand s[i-j] moves through a four-deep register window, two shared-memory loads per four multiply-adds
        for (int i0 = 0; i0 < K.nwords * 32; i0 += blockDim.x * 4) {
            const int ib = i0 + threadIdx.x * 4;             // samples ib .. ib + 3
            double d0 = 0.0, d1 = 0.0, d2 = 0.0, d3 = 0.0;
            if (ib < K.W) {
                // s_in[ib + t - j], t = 0 .. 3: three carried values and four new ones per four steps of j; the products of a
                // trip are formed first, the four running sums take them in j order.  (Past `order` the coefficients are
                // zero-padded and the last trip's extra products are skipped.)
                double a1 = SI(ib + 1), a2 = SI(ib + 2), a3 = SI(ib + 3);
                int j = 0;
                for (; j + 4 <= K.order + 1; j += 4) {
                    const double k0 = kc[j], k1 = kc[j + 1], k2 = kc[j + 2], k3 = kc[j + 3];
                    const double n0 = SI(ib - j), n1 = SI(ib - j - 1), n2 = SI(ib - j - 2), n3 = SI(ib - j - 3);
                    // step j: (n0, a1, a2, a3); j+1: (n1, n0, a1, a2); j+2: (n2, n1, n0, a1); j+3: (n3, n2, n1, n0)
                    d0 = jdadd(jdadd(jdadd(jdadd(d0, jdmul(k0, n0)), jdmul(k1, n1)), jdmul(k2, n2)), jdmul(k3, n3));
                    d1 = jdadd(jdadd(jdadd(jdadd(d1, jdmul(k0, a1)), jdmul(k1, n0)), jdmul(k2, n1)), jdmul(k3, n2));
                    d2 = jdadd(jdadd(jdadd(jdadd(d2, jdmul(k0, a2)), jdmul(k1, a1)), jdmul(k2, n0)), jdmul(k3, n1));
                    d3 = jdadd(jdadd(jdadd(jdadd(d3, jdmul(k0, a3)), jdmul(k1, a2)), jdmul(k2, a1)), jdmul(k3, n0));
                    a1 = n3; a2 = n2; a3 = n1;
                }
                for (; j <= K.order; j++) {
                    const double kj = kc[j], a0 = SI(ib - j);
                    d0 = jdadd(d0, jdmul(kj, a0)); d1 = jdadd(d1, jdmul(kj, a1)); d2 = jdadd(d2, jdmul(kj, a2)); d3 = jdadd(d3, jdmul(kj, a3));
                    a3 = a2; a2 = a1; a1 = a0;
                }
            }
            unsigned nib = 0;
            if (ib + 0 < K.W && ib + 0 >= K.order && fabs(d0) > thr) nib |= 1u;
            if (ib + 1 < K.W && ib + 1 >= K.order && fabs(d1) > thr) nib |= 2u;
            if (ib + 2 < K.W && ib + 2 >= K.order && fabs(d2) > thr) nib |= 4u;
            if (ib + 3 < K.W && ib + 3 >= K.order && fabs(d3) > thr) nib |= 8u;
            // a warp covers 128 samples = 4 words: lane q < 4 assembles word q from the nibbles of lanes 8q .. 8q + 7
            unsigned word = 0;
            const int lane = threadIdx.x & 31;
            for (int m = 0; m < 8; m++) {
                const unsigned nb = __shfl_sync(0xffffffffu, nib, ((lane & 3) << 3) + m);
                word |= nb << (4 * m);
            }
            const int wd = ((i0 + (threadIdx.x & ~31) * 4) >> 5) + lane;
            if (lane < 4 && wd < K.nwords) s_bits[wd] = word;
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
            s_bits[wd] = m;                          // (each thread rewrites only the words it read above)
            if (m) atomicAdd(&s_cnt, __popc(m));
        }
        __syncthreads();
        // band width: for every click, the later clicks at most `order` samples away (order <= 63: three words at most)
        for (int wd = threadIdx.x; wd < K.nwords; wd += blockDim.x) {
            unsigned m = s_bits[wd]; int best = 0;
            while (m) {
                const int b = __ffs(m) - 1; m &= m - 1;
                const int p0 = wd * 32 + b + 1, p1 = min(wd * 32 + b + K.order, K.W - 1);      // bits p0 .. p1
                int cnt = 0;
                for (int q = p0 >> 5; q <= (p1 >> 5) && q < K.nwords; q++) {
                    unsigned v = s_bits[q];
                    if (q == (p0 >> 5)) v &= 0xffffffffu << (p0 & 31);
                    if (q == (p1 >> 5)) v &= 0xffffffffu >> (31 - (p1 & 31));
                    cnt += __popc(v);
                }
                best = max(best, cnt);
            }
            if (best) atomicMax(&s_maxc, best);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            count_out[w] = s_cnt;
            if (s_cnt > 0) {
                int cl = 0;
                while (cl < DC_NCLASS - 1 && s_maxc + 1 > dc_class_ring(cl, K.bw)) cl++;
                class_list[(size_t)cl * n_windows + atomicAdd(&class_count[cl], 1)] = (int)w;
            }
        }
    }
}

// ---- E: interpolation ----------------------------------------------------------------------
#define DC_WARPS 4          // warps per CTA; the ring of a warp is R x R doubles (R = the bin's band width + 1)
__global__ void __launch_bounds__(DC_WARPS * 32)
k_dc_interp(const double *__restrict__ x, double *__restrict__ y, int64_t n, int64_t n_windows, DcConst K, int R,
            const double *__restrict__ acoef, const unsigned *__restrict__ bits_in, const int *__restrict__ count_in,
            const int *__restrict__ list, const int *__restrict__ list_count,
            double *__restrict__ scratch, size_t scratch_per_warp, int *__restrict__ next_window)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int order = K.order, bw = K.bw;                    // ring is R x R, R <= bw: rows k .. k + R - 1 of the band
    const int n_list = *list_count;
    __shared__ unsigned char s_pa[2048], s_pb[2048];   // pair table (a, b), b <= a <= 63, ordered by a then b
    const int npairs_full = order * (order + 1) / 2;
    for (int p = threadIdx.x; p < npairs_full; p += blockDim.x) {
        int a = 1; while (a * (a + 1) / 2 <= p) a++;
        s_pa[p] = (unsigned char)a; s_pb[p] = (unsigned char)(p - a * (a - 1) / 2 + 1);
    }
    __syncthreads();
    const size_t per_warp = ((size_t)R * R + 2 * (size_t)bw + (size_t)R + (size_t)((K.nwords + 1) / 2)) * sizeof(double);
    double *ring = (double *)(smem_raw + (size_t)warp * per_warp);
    double *kc = ring + (size_t)R * R, *aux = kc + bw, *vring = aux + bw;       // vring: right-hand side / forward-substituted values of rows k .. k+R-1
    unsigned *sbits = (unsigned *)(vring + R);                                   // the window's click bitmap
    const int64_t gw = (int64_t)blockIdx.x * DC_WARPS + warp;
    double *S = scratch + (size_t)gw * scratch_per_warp;
    // scratch layout: Lg[nmax][R-1], Dg[nmax], vec[nmax], outv[nmax], idx (as int)[nmax]
    const size_t nmax = (size_t)K.W;
    const int Ls = R - 1;                                                               // a column of the factor has at most R-1 entries
    double *Lg = S, *yd = Lg + nmax * Ls, *vec = yd + nmax, *outv = vec + nmax;         // yd[i] = y_i / D_i
    int *idx = (int *)(outv + nmax), *am = idx + 2 * ((nmax + 1) / 2);
    // ring slot of matrix index k + a given rk = k % bw (a <= order < bw): no integer division in the inner loops
#define WRAP(v) ((v) >= R ? (v) - R : (v))
#define RINGS(r, cidx) ring[(size_t)(r) * R + (cidx)]

    for (;;) {
        int64_t w = 0;
        if (lane == 0) { const int t = atomicAdd(next_window, 1); w = t < n_list ? list[t] : n_windows; }
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
            for (int t0 = 0; t0 < R; t0 += 32) {
                const int t = t0 + lane, i = j - t;
                const int d = (t < R && i >= 0) ? ij - idx[i] : order + 1;
                if (d <= order) RINGS(rj, WRAP(rj + R - t)) = aux[d];
                if (!__ballot_sync(0xffffffffu, d <= order && lane == 31)) break;      // sorted: nothing further couples
            }
        };
        for (int j = 0; j < R; j++) load_row(j, j);
        for (int j = lane; j < R && j < nclk; j += 32) vring[j] = vec[j];
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
            const double v_in = (lane == 0 && k + R < nclk) ? vec[k + R] : 0.0;        // right-hand side of the row entering below: fetched early
            // column k: L(j,k) = A'(j,k) / D_k ; forward substitution v_j -= L(j,k) * y_k
            for (int a = lane + 1; a <= amax; a += 32) {
                const int ra = WRAP(rk + a);
                const double L = RINGS(ra, rk) / Dk;
                RINGS(ra, rk) = L;
                Lg[(size_t)k * Ls + (a - 1)] = L;
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
            load_row(k + R, rk);            // row k+R takes the slot row k leaves (no trailing update touches it)
            if (lane == 0 && k + R < nclk) vring[rk] = v_in;
            __syncwarp();
            rk = rk + 1 == R ? 0 : rk + 1;
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
            double L0 = lane < amax ? Lg[(size_t)(nclk - 1) * Ls + lane] : 0.0;
            double L1 = lane + 32 < amax ? Lg[(size_t)(nclk - 1) * Ls + lane + 32] : 0.0;
            for (int i = nclk - 1; i >= 0; i--) {
                // next row's operands are fetched before this row's dependent chain starts
                int amax_n = 0; double oy_n = 0.0, L0n = 0.0, L1n = 0.0;
                if (i > 0) {
                    amax_n = am[i - 1]; oy_n = yd[i - 1];
                    if (lane < amax_n) L0n = Lg[(size_t)(i - 1) * Ls + lane];
                    if (lane + 32 < amax_n) L1n = Lg[(size_t)(i - 1) * Ls + lane + 32];
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
    int *d_cnt = jt_dalloc<int>(c, nw + 2 * DC_NCLASS);
    int *d_class_count = d_cnt + nw, *d_next = d_class_count + DC_NCLASS;
    int *d_class_list = jt_dalloc<int>(c, (size_t)DC_NCLASS * nw);
    JT_CUDA(cudaMemsetAsync(d_class_count, 0, 2 * DC_NCLASS * sizeof(int), c->stream));
    // the windows pass through except where clicks are repaired: out = in, then E overwrites
    JT_CUDA(cudaMemcpyAsync(o.d, in.d, sizeof(double) * (size_t)in.n, cudaMemcpyDeviceToDevice, c->stream));
    // plane length of the span of DC_AC_G overlapping windows; = 8 mod 16 so that the four planes start 8 "double banks" apart
    int PLa = (K.W + (DC_AC_G - 1) * std::min(K.hop, K.W) + 3) / 4 + 1;
    PLa = (PLa + 7) / 16 * 16 + 8;
    const size_t smemA = sizeof(double) * 4 * (size_t)PLa;
    const size_t smemC = sizeof(double) * (4 * ((((K.order + 3) & ~3) + K.W + 4 + 3) / 4 + 1) + K.bw + 4);
    if (DC_AC_G * ((K.order + DC_AC_LT) / DC_AC_LT) > 64) JT_THROW(JT_ERR_UNSUPPORTED, "adeclick AR order %d", K.order);
    jt_smem_optin((const void *)k_dc_autocorr, (size_t)(smemA));
    jt_smem_optin((const void *)k_dc_detect, (size_t)(smemC));
    { JtLaunch L(c, "adeclick:autocorr");
    k_dc_autocorr<<<jt_grid_for((nw + DC_AC_G - 1) / DC_AC_G, 1, c->num_sms, 16), 64, smemA, c->stream>>>((const double *)in.d, in.n, nw, K, PLa, d_r); }
    { JtLaunch L(c, "adeclick:levinson");
    k_dc_levinson<<<(int)((nw + 127) / 128), 128, 0, c->stream>>>(d_r, nw, K, d_a, d_sig); }
    { JtLaunch L(c, "adeclick:detect");
    k_dc_detect<<<jt_grid_for(nw, 1, c->num_sms, 32), 256, smemC, c->stream>>>((const double *)in.d, in.n, nw, K, d_a, d_sig, d_bits, d_cnt, d_class_count, d_class_list); }
    // one launch per band-width bin, widest first; a bin's grid fills the SMs as far as its ring allows (64 registers per
    // thread: 32 warps per SM at most)
    const size_t nmax = (size_t)K.W;
    JtLaunch L(c, "adeclick:interp", DC_NCLASS);
    for (int cl = DC_NCLASS - 1; cl >= 0; cl--) {
        const int R = std::min(dc_class_ring(cl, K.bw), K.bw);
        if (cl < DC_NCLASS - 1 && dc_class_ring(cl, K.bw) >= K.bw) continue;          // this bin is empty by construction: the last one covers it
        const size_t per_warp = ((size_t)R * R + 2 * (size_t)K.bw + (size_t)R + (size_t)((K.nwords + 1) / 2)) * sizeof(double);
        const size_t smemE = per_warp * DC_WARPS;
        jt_smem_optin((const void *)k_dc_interp, (size_t)(smemE));
        const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(32 / DC_WARPS, (227 * 1024) / (smemE + 4096 + 1024)));
        int gridE = (int)std::min<int64_t>((nw + DC_WARPS - 1) / DC_WARPS, (int64_t)c->num_sms * per_sm);
        if (gridE < 1) gridE = 1;
        const size_t scratch_per_warp = nmax * (size_t)(R - 1) + 3 * nmax + 2 * ((nmax + 1) / 2) + 8;
        double *scratch = jt_dalloc<double>(c, scratch_per_warp * (size_t)gridE * DC_WARPS);
        k_dc_interp<<<gridE, DC_WARPS * 32, smemE, c->stream>>>((const double *)in.d, (double *)o.d, in.n, nw, K, R, d_a, d_bits, d_cnt,
                                                                d_class_list + (size_t)cl * nw, d_class_count + cl,
                                                                scratch, scratch_per_warp, d_next + cl);
    }
    return o;
}
