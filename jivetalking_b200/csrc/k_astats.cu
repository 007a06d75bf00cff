// astats (libavfilter/af_astats.c, one channel, length=0.05, reset=0) over x[0, n_upto):
// "astats=metadata=1:measure_perchannel=all|0" (reference: filters.go:624,
// analyser_bands.go:33, analyser_output.go:18; formulas docs/Spectral-Metrics-Reference.md:35-56).
// All of it is reductions plus two scalar linear/max recurrences:
//   A  one streaming pass: sums, extrema, differences, zero crossings, bit mask, 8192-bin histogram
//   B  counts / run lengths at the global extrema (Peak_count, Flat_factor)
//   C  50 ms exponential mean square (RMS_peak / RMS_trough): blocked linear recurrence with
//      carry propagation (exact, no warm-up)
//   D  Noise_floor = min over time of the 50 ms sliding-window max |x|: van Herk / Gil-Werman
//      block prefix/suffix maxima
#include "jt_internal.h"
#include "jt_device.cuh"
#include "jt_lanes.cuh"
#include <cfloat>
#include <cstdio>

#define AS_HIST 8192

struct AsPartA {
    double sum, sumsq, mn, mx, min_nz, min_diff, max_diff, diff_sum, diff_sumsq;
    unsigned long long zero_runs, mask;
};

template <class T> struct AsTraits;
template <> struct AsTraits<int16_t> {
    static __device__ __forceinline__ double d(int16_t v) { return (double)v; }
    static __device__ __forceinline__ double nd(int16_t v) { return (double)v / 32767.0; }
    static __device__ __forceinline__ unsigned long long absi(int16_t v) { int i = v; return (unsigned long long)(i < 0 ? -i : i); }
};
template <> struct AsTraits<int32_t> {
    static __device__ __forceinline__ double d(int32_t v) { return (double)v; }
    static __device__ __forceinline__ double nd(int32_t v) { return (double)v / 2147483647.0; }
    static __device__ __forceinline__ unsigned long long absi(int32_t v) { long long i = v; return (unsigned long long)(i < 0 ? -i : i); }
};
template <> struct AsTraits<float> {
    static __device__ __forceinline__ double d(float v) { return (double)v; }
    static __device__ __forceinline__ double nd(float v) { return (double)v; }
    static __device__ __forceinline__ unsigned long long absi(float v) { long long i = __double2ll_rn((double)v * 2147483648.0); return (unsigned long long)(i < 0 ? -i : i); }
};
template <> struct AsTraits<double> {
    static __device__ __forceinline__ double d(double v) { return v; }
    static __device__ __forceinline__ double nd(double v) { return v; }
    static __device__ __forceinline__ unsigned long long absi(double v) { long long i = __double2ll_rn(v * 9223372036854775808.0); return i < 0 ? (unsigned long long)(-(i + 1)) + 1ull : (unsigned long long)i; }
};

template <class T>
__global__ void __launch_bounds__(256)
k_astats_a(const T *__restrict__ x, int64_t n, AsPartA *__restrict__ parts, unsigned long long *__restrict__ ghist, int64_t back)
{   // back: samples readable BEFORE x[0] (a chunk of a longer stream); 0 = x[0] starts the stream
    __shared__ unsigned int shist[AS_HIST];
    __shared__ AsPartA swarp[8];
    for (int i = threadIdx.x; i < AS_HIST; i += blockDim.x) shist[i] = 0;
    __syncthreads();
    AsPartA a;
    a.sum = a.sumsq = a.diff_sum = a.diff_sumsq = 0; a.mn = DBL_MAX; a.mx = -DBL_MAX; a.min_nz = DBL_MAX;
    a.min_diff = DBL_MAX; a.max_diff = 0; a.zero_runs = 0; a.mask = 0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const T v = x[i];
        const double d = AsTraits<T>::d(v), nd = AsTraits<T>::nd(v);
        a.sum += nd; a.sumsq = fma(nd, nd, a.sumsq);
        a.mn = fmin(a.mn, d); a.mx = fmax(a.mx, d);
        if (d != 0) {
            a.min_nz = fmin(a.min_nz, fabs(d));
            int64_t j = i - 1;
            while (j >= -back && AsTraits<T>::d(x[j]) == 0) j--;
            const double prev = j >= -back ? AsTraits<T>::d(x[j]) : 0.0;
            a.zero_runs += ((d > 0) != (prev > 0));
        }
        if (i > 0 || back > 0) {
            const double df = d - AsTraits<T>::d(x[i - 1]), ad = fabs(df);
            a.min_diff = fmin(a.min_diff, ad); a.max_diff = fmax(a.max_diff, ad);
            a.diff_sum += ad; a.diff_sumsq = fma(df, df, a.diff_sumsq);
        }
        a.mask |= AsTraits<T>::absi(v);
        int idx = __double2int_rn(fmin(fabs(nd), 1.0) * (AS_HIST - 1));
        idx = max(0, min(AS_HIST - 1, idx));
        atomicAdd(&shist[idx], 1u);
    }
    // warp then block reduce
    a.sum = jt_warp_sum(a.sum); a.sumsq = jt_warp_sum(a.sumsq);
    a.diff_sum = jt_warp_sum(a.diff_sum); a.diff_sumsq = jt_warp_sum(a.diff_sumsq);
    a.mn = jt_warp_min(a.mn); a.mx = jt_warp_max(a.mx); a.min_nz = jt_warp_min(a.min_nz);
    a.min_diff = jt_warp_min(a.min_diff); a.max_diff = jt_warp_max(a.max_diff);
    for (int o = 16; o; o >>= 1) {
        a.zero_runs += __shfl_xor_sync(0xffffffffu, a.zero_runs, o);
        a.mask |= __shfl_xor_sync(0xffffffffu, a.mask, o);
    }
    if ((threadIdx.x & 31) == 0) swarp[threadIdx.x >> 5] = a;
    __syncthreads();
    if (threadIdx.x == 0) {
        AsPartA r = swarp[0];
        for (int w = 1; w < (blockDim.x >> 5); w++) {
            const AsPartA &b = swarp[w];
            r.sum += b.sum; r.sumsq += b.sumsq; r.diff_sum += b.diff_sum; r.diff_sumsq += b.diff_sumsq;
            r.mn = fmin(r.mn, b.mn); r.mx = fmax(r.mx, b.mx); r.min_nz = fmin(r.min_nz, b.min_nz);
            r.min_diff = fmin(r.min_diff, b.min_diff); r.max_diff = fmax(r.max_diff, b.max_diff);
            r.zero_runs += b.zero_runs; r.mask |= b.mask;
        }
        parts[blockIdx.x] = r;
    }
    for (int i = threadIdx.x; i < AS_HIST; i += blockDim.x)
        if (shist[i]) atomicAdd(&ghist[i], (unsigned long long)shist[i]);
}

// counts[0..3] = min_count, max_count, min_runs (sum of run^2), max_runs
template <class T>
__global__ void __launch_bounds__(256)
k_astats_b(const T *__restrict__ x, int64_t n, const AsPartA *__restrict__ total, unsigned long long *__restrict__ counts)
{
    const double gmin = total->mn, gmax = total->mx;
    unsigned long long c0 = 0, c1 = 0, c2 = 0, c3 = 0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double d = AsTraits<T>::d(x[i]);
        if (d == gmin) {
            c0++;
            if (i + 1 < n && AsTraits<T>::d(x[i + 1]) != gmin) {
                unsigned long long len = 1; int64_t j = i - 1;
                while (j >= 0 && AsTraits<T>::d(x[j]) == gmin) { len++; j--; }
                c2 += len * len;
            }
        }
        if (d == gmax) {
            c1++;
            if (i + 1 < n && AsTraits<T>::d(x[i + 1]) != gmax) {
                unsigned long long len = 1; int64_t j = i - 1;
                while (j >= 0 && AsTraits<T>::d(x[j]) == gmax) { len++; j--; }
                c3 += len * len;
            }
        }
    }
    for (int o = 16; o; o >>= 1) {
        c0 += __shfl_xor_sync(0xffffffffu, c0, o); c1 += __shfl_xor_sync(0xffffffffu, c1, o);
        c2 += __shfl_xor_sync(0xffffffffu, c2, o); c3 += __shfl_xor_sync(0xffffffffu, c3, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (c0) atomicAdd(&counts[0], c0); if (c1) atomicAdd(&counts[1], c1);
        if (c2) atomicAdd(&counts[2], c2); if (c3) atomicAdd(&counts[3], c3);
    }
}

// combine the per-CTA partials of pass A and initialise the min / max cells of C and D.  Fixed order, so deterministic: lane l
// folds parts l, l + 32, ... in sequence, then the 32 lane results fold in a fixed butterfly (one thread walking all ~1200
// partials took 0.35 ms per launch, 2.4 ms per step over the seven astats runs of ProcessAudio)
__device__ __forceinline__ void as_fold(AsPartA &r, const AsPartA &b)
{
    r.sum += b.sum; r.sumsq += b.sumsq; r.diff_sum += b.diff_sum; r.diff_sumsq += b.diff_sumsq;
    r.mn = fmin(r.mn, b.mn); r.mx = fmax(r.mx, b.mx); r.min_nz = fmin(r.min_nz, b.min_nz);
    r.min_diff = fmin(r.min_diff, b.min_diff); r.max_diff = fmax(r.max_diff, b.max_diff);
    r.zero_runs += b.zero_runs; r.mask |= b.mask;
}
__global__ void __launch_bounds__(32)
k_astats_reduce(const AsPartA *__restrict__ parts, int nparts, AsPartA *__restrict__ total, double *__restrict__ mm, float *__restrict__ nf)
{
    __shared__ AsPartA sh[32];
    const int lane = threadIdx.x;
    if (lane < nparts) {
        AsPartA r = parts[lane];
        for (int i = lane + 32; i < nparts; i += 32) as_fold(r, parts[i]);
        sh[lane] = r;
    }
    __syncwarp();
    const int live = min(nparts, 32);
    for (int o = 16; o; o >>= 1) {
        if (lane < o && lane + o < live) { AsPartA r = sh[lane]; as_fold(r, sh[lane + o]); sh[lane] = r; }
        __syncwarp();
    }
    if (lane == 0) { *total = sh[0]; mm[0] = DBL_MAX; mm[1] = 0.0; *nf = FLT_MAX; }
}

// C1: per block of BS samples, zero-state end value of avg = avg*mult + (1-mult)*nd^2.
// One lane per block; each lane's samples arrive as TMA-staged 256-byte rows (jt_lanes.cuh).
#define AS_BS 1024
template <class T> struct AsRow { static constexpr int R = 256 / (int)sizeof(T); };
template <class T>
__global__ void __launch_bounds__(64)
k_astats_c1(const T *__restrict__ x, int64_t n, double mult, double *__restrict__ fin)
{
    constexpr int R = AsRow<T>::R;
    extern __shared__ __align__(16) unsigned char smem[];
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t s = min(b * AS_BS, n), e = min(s + AS_BS, n);
    LaneStage<T, R, 2> in;
    in.init(smem + (size_t)(threadIdx.x >> 5) * LaneStage<T, R, 2>::WARP_BYTES, x + s, e - s);
    double avg = 0; const double om = 1.0 - mult;
    in.prime();
    for (int tile = 0; tile < in.ntiles; tile++) {
        in.prefetch();
        const T *row = in.wait(tile);
        const int nv = in.valid(tile);
        int k = 0;
        for (; k + 8 <= nv; k += 8) {
            double q[8];
#pragma unroll
            for (int j = 0; j < 8; j++) { const double nd = AsTraits<T>::nd(row[k + j]); q[j] = om * nd * nd; }
#pragma unroll
            for (int j = 0; j < 8; j++) avg = avg * mult + q[j];
        }
        for (; k < nv; k++) { const double nd = AsTraits<T>::nd(row[k]); avg = avg * mult + om * nd * nd; }
        in.release();
    }
    if (e > s) fin[b] = avg;
}
// carry[b] = state entering block b.  cy' = cy * M + fin[b] is affine in cy: each thread composes its
// slice of blocks, thread 0 chains the 1024 slices, every thread replays its slice from its true entry value.
__global__ void __launch_bounds__(1024)
k_astats_carry(const double *__restrict__ fin, double *__restrict__ carry, int64_t nb, double mult_bs)
{
    __shared__ double sA[1024], sB[1024], sIn[1024];
    const int t = threadIdx.x;
    const int64_t per = (nb + 1023) / 1024, h0 = min((int64_t)t * per, nb), h1 = min(h0 + per, nb);
    double A = 1.0, B = 0.0;
    for (int64_t h = h0; h < h1; h++) { A *= mult_bs; B = B * mult_bs + fin[h]; }
    sA[t] = A; sB[t] = B;
    __syncthreads();
    if (t == 0) { double cy = 0; for (int i = 0; i < 1024; i++) { sIn[i] = cy; cy = sA[i] * cy + sB[i]; } }
    __syncthreads();
    double cy = sIn[t];
    for (int64_t h = h0; h < h1; h++) { carry[h] = cy; cy = cy * mult_bs + fin[h]; }
}
template <class T>
__global__ void __launch_bounds__(64)
k_astats_c2(const T *__restrict__ x, int64_t n, double mult, const double *__restrict__ carry,
            int64_t tc, double *__restrict__ mm /* [0]=min (init DBL_MAX), [1]=max */)
{
    constexpr int R = AsRow<T>::R;
    extern __shared__ __align__(16) unsigned char smem[];
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t s = min(b * AS_BS, n), e = min(s + AS_BS, n);
    LaneStage<T, R, 2> in;
    in.init(smem + (size_t)(threadIdx.x >> 5) * LaneStage<T, R, 2>::WARP_BYTES, x + s, e - s);
    double lo = DBL_MAX, hi = 0;
    double avg = e > s ? carry[b] : 0.0; const double om = 1.0 - mult;
    in.prime();
    for (int tile = 0; tile < in.ntiles; tile++) {
        in.prefetch();
        const T *row = in.wait(tile);
        const int nv = in.valid(tile);
        const int64_t i0 = s + (int64_t)tile * R;
        if (i0 >= tc) {
            // avg >= 0: non-negative doubles order like their bit patterns, so min / max run on the integer
            // pipe (f64 min/max issue at a third of the f64 add rate on sm_100a, profiles/ubench_r1.txt)
            unsigned long long ulo = (unsigned long long)__double_as_longlong(lo), uhi = (unsigned long long)__double_as_longlong(hi);
            int k = 0;
            for (; k + 8 <= nv; k += 8) {
                double q[8];
#pragma unroll
                for (int j = 0; j < 8; j++) { const double nd = AsTraits<T>::nd(row[k + j]); q[j] = om * nd * nd; }
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    avg = avg * mult + q[j];
                    const unsigned long long u = (unsigned long long)__double_as_longlong(avg);
                    ulo = min(ulo, u); uhi = max(uhi, u);
                }
            }
            for (; k < nv; k++) {
                const double nd = AsTraits<T>::nd(row[k]); avg = avg * mult + om * nd * nd;
                const unsigned long long u = (unsigned long long)__double_as_longlong(avg);
                ulo = min(ulo, u); uhi = max(uhi, u);
            }
            lo = __longlong_as_double((long long)ulo); hi = __longlong_as_double((long long)uhi);
        } else {
            for (int k = 0; k < nv; k++) {
                const double nd = AsTraits<T>::nd(row[k]); avg = avg * mult + om * nd * nd;
                if (i0 + k >= tc) { lo = fmin(lo, avg); hi = fmax(hi, avg); }
            }
        }
        in.release();
    }
    lo = jt_warp_min(lo); hi = jt_warp_max(hi);
    if ((threadIdx.x & 31) == 0) { jt_atomic_min_nonneg(&mm[0], lo); jt_atomic_max_nonneg(&mm[1], hi); }
}

// D: Noise_floor = min over window ends i >= tc-1 of max |nd| over [i-tc+1, i], and how often that minimum
// occurs.  van Herk / Gil-Werman with tc-sized blocks, one CTA per block b: the window ending at offset j of
// block b is (suffix of block b-1 from offset j+1) U (prefix of block b up to j); both running maxima are
// block-parallel scans in shared memory, nothing but the per-block (min, count) pair is written.
#define AS_NF_THREADS 256
__device__ __forceinline__ float as_block_excl_scan_max(float v, float *s_w)
{   // exclusive forward scan of max over the CTA's threads (values >= 0)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float inc = v;
    for (int o = 1; o < 32; o <<= 1) { const float t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc = fmaxf(inc, t); }
    if (lane == 31) s_w[warp] = inc;
    __syncthreads();
    float base = 0.f;
    for (int w = 0; w < warp; w++) base = fmaxf(base, s_w[w]);
    float exc = __shfl_up_sync(0xffffffffu, inc, 1); if (lane == 0) exc = 0.f;
    __syncthreads();
    return fmaxf(base, exc);
}
template <class T>
__global__ void __launch_bounds__(AS_NF_THREADS)
k_astats_nf(const T *__restrict__ x, int64_t n, int tc, float *__restrict__ bmin, unsigned *__restrict__ bcnt)
{
    extern __shared__ float sm_nf[];                 // suf[tc + 1], pre[tc]
    __shared__ float s_w[AS_NF_THREADS / 32];
    __shared__ float s_red[AS_NF_THREADS / 32]; __shared__ unsigned s_cnt[AS_NF_THREADS / 32];
    float *suf = sm_nf, *pre = sm_nf + tc + 1;
    const int t = threadIdx.x, per = (tc + AS_NF_THREADS - 1) / AS_NF_THREADS;
    for (int64_t b = blockIdx.x; b * tc < n; b += gridDim.x) {
        const int64_t s = b * tc;
        const int len = (int)min((int64_t)tc, n - s);
        __syncthreads();
        for (int j = t; j < tc; j += AS_NF_THREADS) {
            suf[j] = b > 0 ? (float)fabs(AsTraits<T>::nd(x[s - tc + j])) : 0.f;
            pre[j] = j < len ? (float)fabs(AsTraits<T>::nd(x[s + j])) : 0.f;
        }
        if (t == 0) suf[tc] = 0.f;
        __syncthreads();
        // prefix maxima of the current block: thread t owns [t*per, (t+1)*per)
        {
            const int j0 = min(t * per, tc), j1 = min(j0 + per, tc);
            float m = 0.f;
            for (int j = j0; j < j1; j++) { m = fmaxf(m, pre[j]); pre[j] = m; }
            const float base = as_block_excl_scan_max(m, s_w);
            for (int j = j0; j < j1; j++) pre[j] = fmaxf(pre[j], base);
        }
        // suffix maxima of the previous block: thread t owns the mirrored range, scanned right to left
        {
            const int r0 = min(t * per, tc), r1 = min(r0 + per, tc);        // mirrored offsets
            float m = 0.f;
            for (int r = r0; r < r1; r++) { const int j = tc - 1 - r; m = fmaxf(m, suf[j]); suf[j] = m; }
            const float base = as_block_excl_scan_max(m, s_w);
            for (int r = r0; r < r1; r++) { const int j = tc - 1 - r; suf[j] = fmaxf(suf[j], base); }
        }
        __syncthreads();
        // windows ending in this block: offsets j with s + j >= tc - 1
        const int jlo = b == 0 ? tc - 1 : 0;
        float lo = FLT_MAX;
        for (int j = jlo + t; j < len; j += AS_NF_THREADS) lo = fminf(lo, fmaxf(suf[j + 1], pre[j]));
        for (int o = 16; o; o >>= 1) lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        if ((t & 31) == 0) s_red[t >> 5] = lo;
        __syncthreads();
        float g = FLT_MAX;
        for (int w = 0; w < AS_NF_THREADS / 32; w++) g = fminf(g, s_red[w]);
        unsigned c = 0;
        for (int j = jlo + t; j < len; j += AS_NF_THREADS) c += (fmaxf(suf[j + 1], pre[j]) == g);
        for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
        if ((t & 31) == 0) s_cnt[t >> 5] = c;
        __syncthreads();
        if (t == 0) { unsigned tot = 0; for (int w = 0; w < AS_NF_THREADS / 32; w++) tot += s_cnt[w]; bmin[b] = g; bcnt[b] = tot; }
    }
}
__global__ void __launch_bounds__(1024)
k_astats_nf_reduce(const float *__restrict__ bmin, const unsigned *__restrict__ bcnt, int64_t nblk, float *__restrict__ gmin,
                   unsigned long long *__restrict__ cnt)
{
    __shared__ float s_m[32]; __shared__ unsigned long long s_c[32];
    float lo = FLT_MAX;
    for (int64_t i = threadIdx.x; i < nblk; i += 1024) lo = fminf(lo, bmin[i]);
    for (int o = 16; o; o >>= 1) lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    if ((threadIdx.x & 31) == 0) s_m[threadIdx.x >> 5] = lo;
    __syncthreads();
    float g = FLT_MAX; for (int w = 0; w < 32; w++) g = fminf(g, s_m[w]);
    unsigned long long c = 0;
    for (int64_t i = threadIdx.x; i < nblk; i += 1024) if (bmin[i] == g) c += bcnt[i];
    for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) s_c[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) { unsigned long long tot = 0; for (int w = 0; w < 32; w++) tot += s_c[w]; *gmin = g; *cnt = tot; }
}

struct AstatsHost { AsPartA total; unsigned long long hist[AS_HIST + 8]; double mm[2]; float nf; };

template <class T>
static void astats_launch_t(jt_ctx *c, const Sig &in, int64_t n, AstatsPending &pd, int64_t own0 = 0, int64_t global_first = 0)
{
    // chunk mode (own0 > 0): statistics of samples [own0, own0 + n) of `in`, which are samples
    // [global_first, global_first + n) of a longer stream; the samples before own0 are context only
    const T *x = (const T *)in.d + own0;
    const int64_t back = own0;
    const double time_constant = 0.05;
    const int tc = (int)std::fmax(time_constant * in.rate + .5, 1);
    const double mult = exp((-1 / time_constant / in.rate));
    pd.n = n; pd.fmt = in.fmt; pd.rate = in.rate; pd.tc = tc;

    const int gridA = jt_grid_for(n, 256, c->num_sms, 8);
    AsPartA *d_parts = jt_dalloc<AsPartA>(c, gridA + 1), *d_total = d_parts + gridA;
    unsigned long long *d_hist = jt_dalloc<unsigned long long>(c, AS_HIST + 8);
    unsigned long long *d_counts = d_hist + AS_HIST;     // 4 counts + noise-floor count
    double *d_mm = jt_dalloc<double>(c, 2);
    float *d_nf = jt_dalloc<float>(c, 1);
    JT_CUDA(cudaMemsetAsync(d_hist, 0, sizeof(unsigned long long) * (AS_HIST + 8), c->stream));
    { JtLaunch L(c, "astats:sums_hist", 2);
      k_astats_a<T><<<gridA, 256, 0, c->stream>>>(x, n, d_parts, d_hist, back);
      k_astats_reduce<<<1, 32, 0, c->stream>>>(d_parts, gridA, d_total, d_mm, d_nf); }
    // B: extrema counts
    { JtLaunch L(c, "astats:extrema_runs"); k_astats_b<T><<<gridA, 256, 0, c->stream>>>(x, n, d_total, d_counts); }
    // C: exponential mean square min/max
    // the 50 ms exponential window is warmed up over the context (>= 37 time constants for an exact carry-in)
    const int64_t wu = std::min<int64_t>(back, 40 * (int64_t)tc);
    const T *xc = x - wu; const int64_t nc = n + wu;
    const int64_t track_from = std::max<int64_t>(wu, (int64_t)tc - (global_first - wu));
    const int BS = AS_BS; const int64_t nb = (nc + BS - 1) / BS;
    if (global_first + n > tc) {
        double *d_fin = jt_dalloc<double>(c, nb), *d_carry = jt_dalloc<double>(c, nb);
        JtLaunch L(c, "astats:rms_scan", 3);
        const size_t smemC = 2 * LaneStage<T, AsRow<T>::R, 2>::WARP_BYTES;
        jt_smem_optin((const void *)k_astats_c1<T>, (size_t)(smemC));
        jt_smem_optin((const void *)k_astats_c2<T>, (size_t)(smemC));
        k_astats_c1<T><<<(int)((nb + 63) / 64), 64, smemC, c->stream>>>(xc, nc, mult, d_fin);
        k_astats_carry<<<1, 1024, 0, c->stream>>>(d_fin, d_carry, nb, pow(mult, (double)BS));
        k_astats_c2<T><<<(int)((nb + 63) / 64), 64, smemC, c->stream>>>(xc, nc, mult, d_carry, track_from, d_mm);
    }
    // D: noise floor
    // windows END in the owned range; a chunk in mid-stream takes the tc-1 samples before it as the first windows' body
    const int64_t nfb = std::min<int64_t>(back, tc - 1);
    const T *xn = x - nfb; const int64_t nn = n + nfb;
    if (nn >= tc) {
        const int64_t nbt = (nn + tc - 1) / tc;
        float *d_bmin = jt_dalloc<float>(c, nbt); unsigned *d_bcnt = jt_dalloc<unsigned>(c, nbt);
        const size_t smemD = sizeof(float) * (2 * (size_t)tc + 1);
        if (smemD > 200 * 1024) JT_THROW(JT_ERR_UNSUPPORTED, "astats at %d Hz (50 ms window of %d samples)", in.rate, tc);
        jt_smem_optin((const void *)k_astats_nf<T>, (size_t)(smemD));
        JtLaunch L(c, "astats:noise_floor", 2);
        k_astats_nf<T><<<jt_grid_for(nbt, 1, c->num_sms, 32), AS_NF_THREADS, smemD, c->stream>>>(xn, nn, tc, d_bmin, d_bcnt);
        k_astats_nf_reduce<<<1, 1024, 0, c->stream>>>(d_bmin, d_bcnt, nbt, d_nf, d_counts + 4);
    }
    AstatsHost *h = (AstatsHost *)jt_pinned_bytes(c, sizeof(AstatsHost));
    pd.host = h;
    jt_copy_small(c, &h->total, d_total, sizeof(AsPartA));
    jt_copy_small(c, h->hist, d_hist, sizeof(unsigned long long) * (AS_HIST + 8));
    jt_copy_small(c, h->mm, d_mm, sizeof(double) * 2);
    jt_copy_small(c, &h->nf, d_nf, sizeof(float));
    pd.ev = jt_record_event(c);
}

void jt_astats_finish(jt_ctx *c, AstatsPending &pd, AstatsResult &out)
{
    for (int i = 0; i < JT_AS_COUNT; i++) out.v[i] = NAN;
    out.overall_rms = out.overall_peak = NAN; out.nb_samples = 0;
    if (pd.n <= 0 || !pd.host) return;
    JT_CUDA(cudaEventSynchronize(pd.ev));
    jt_astats_host_finalize(pd.host, pd.n, pd.fmt, pd.tc, out);
}

size_t jt_astats_host_bytes() { return sizeof(AstatsHost); }

// merge the partial statistics of a later chunk into `dst` (chunks in stream order)
void jt_astats_host_merge(void *dst_, const void *src_)
{
    AstatsHost &d = *(AstatsHost *)dst_; const AstatsHost &s = *(const AstatsHost *)src_;
    // Peak_count / Flat_factor runs and the noise-floor count belong to whoever holds the global extreme
    const bool same_mn = s.total.mn == d.total.mn, same_mx = s.total.mx == d.total.mx;
    if (s.total.mn < d.total.mn) { d.hist[AS_HIST + 0] = s.hist[AS_HIST + 0]; d.hist[AS_HIST + 2] = s.hist[AS_HIST + 2]; }
    else if (same_mn) { d.hist[AS_HIST + 0] += s.hist[AS_HIST + 0]; d.hist[AS_HIST + 2] += s.hist[AS_HIST + 2]; }
    if (s.total.mx > d.total.mx) { d.hist[AS_HIST + 1] = s.hist[AS_HIST + 1]; d.hist[AS_HIST + 3] = s.hist[AS_HIST + 3]; }
    else if (same_mx) { d.hist[AS_HIST + 1] += s.hist[AS_HIST + 1]; d.hist[AS_HIST + 3] += s.hist[AS_HIST + 3]; }
    if (s.nf < d.nf) { d.nf = s.nf; d.hist[AS_HIST + 4] = s.hist[AS_HIST + 4]; }
    else if (s.nf == d.nf) d.hist[AS_HIST + 4] += s.hist[AS_HIST + 4];
    AsPartA &r = d.total; const AsPartA &b = s.total;
    r.sum += b.sum; r.sumsq += b.sumsq; r.diff_sum += b.diff_sum; r.diff_sumsq += b.diff_sumsq;
    r.mn = std::fmin(r.mn, b.mn); r.mx = std::fmax(r.mx, b.mx); r.min_nz = std::fmin(r.min_nz, b.min_nz);
    r.min_diff = std::fmin(r.min_diff, b.min_diff); r.max_diff = std::fmax(r.max_diff, b.max_diff);
    r.zero_runs += b.zero_runs; r.mask |= b.mask;
    for (int i = 0; i < AS_HIST; i++) d.hist[i] += s.hist[i];
    d.mm[0] = std::fmin(d.mm[0], s.mm[0]); d.mm[1] = std::fmax(d.mm[1], s.mm[1]);
}

void jt_astats_host_finalize(const void *host, int64_t n, int fmt, int tc, AstatsResult &out)
{
    for (int i = 0; i < JT_AS_COUNT; i++) out.v[i] = NAN;
    out.overall_rms = out.overall_peak = NAN; out.nb_samples = 0;
    if (n <= 0 || !host) return;
    const AstatsHost *h = (const AstatsHost *)host;
    const AsPartA &r = h->total;
    const unsigned long long *hist = h->hist;
    const int maxbits = fmt == JT_FMT_S16 ? 16 : (fmt == JT_FMT_FLT || fmt == JT_FMT_S32) ? 32 : 64;
    const float h_nf = h->nf;
    const double N = (double)n;
    const double scale = fmt == JT_FMT_S16 ? 32767.0 : fmt == JT_FMT_S32 ? 2147483647.0 : 1.0;
    const double nmin = r.mn / scale, nmax = r.mx / scale;
    double min_s2 = h->mm[0], max_s2 = h->mm[1];
    if (n <= tc) min_s2 = max_s2 = r.sumsq / N;      // af_astats.c: fewer samples than the window
    const double min_count = (double)hist[AS_HIST + 0], max_count = (double)hist[AS_HIST + 1];
    const double min_runs = (double)hist[AS_HIST + 2], max_runs = (double)hist[AS_HIST + 3];
#define DB(x) (log10(x) * 20)
    double *v = out.v;
    for (int i = 0; i < JT_AS_COUNT; i++) v[i] = NAN;
    v[JT_AS_DC_offset] = r.sum / N;
    v[JT_AS_Min_level] = r.mn; v[JT_AS_Max_level] = r.mx;
    v[JT_AS_Min_difference] = r.min_diff; v[JT_AS_Max_difference] = r.max_diff;
    v[JT_AS_Mean_difference] = r.diff_sum / (N - 1);
    v[JT_AS_RMS_difference] = sqrt(r.diff_sumsq / (N - 1));
    v[JT_AS_Peak_level] = DB(std::fmax(-nmin, nmax));
    v[JT_AS_RMS_level] = DB(sqrt(r.sumsq / N));
    v[JT_AS_RMS_peak] = DB(sqrt(max_s2));
    v[JT_AS_RMS_trough] = DB(sqrt(min_s2));
    v[JT_AS_Crest_factor] = r.sumsq ? std::fmax(-r.mn, r.mx) / sqrt(r.sumsq / N) : 1;
    v[JT_AS_Flat_factor] = DB((min_runs + max_runs) / (min_count + max_count));
    v[JT_AS_Noise_floor] = n >= tc ? DB((double)h_nf) : NAN;
    v[JT_AS_Noise_floor_count] = (double)hist[AS_HIST + 4];
    { double e = 0; for (int i = 0; i < AS_HIST; i++) { double p = hist[i] / N; if (p > 1e-8) e += p * log2(p); }
      v[JT_AS_Entropy] = -e / log2((double)AS_HIST); }
    { int depth = 0; for (int i = 0; i < maxbits; i++) depth += !!(r.mask & (1ULL << i)); v[JT_AS_Bit_depth] = depth; }
    v[JT_AS_Dynamic_range] = DB(2 * std::fmax(fabs(r.mn), fabs(r.mx)) / r.min_nz);
    v[JT_AS_Zero_crossings] = (double)r.zero_runs;
    v[JT_AS_Zero_crossings_rate] = r.zero_runs / N;
    v[JT_AS_Number_of_samples] = N;
    out.overall_rms = v[JT_AS_RMS_level];
    out.overall_peak = v[JT_AS_Peak_level];
    out.nb_samples = N;
    (void)min_count; (void)max_count;
#undef DB
}

void jt_astats_launch(jt_ctx *c, const Sig &in, int64_t n_upto, AstatsPending &pd)
{
    pd = AstatsPending();
    const int64_t n = std::min(n_upto, in.n);
    if (n <= 0) return;
    if (in.fmt == JT_FMT_S16) astats_launch_t<int16_t>(c, in, n, pd);
    else if (in.fmt == JT_FMT_S32) astats_launch_t<int32_t>(c, in, n, pd);
    else if (in.fmt == JT_FMT_FLT) astats_launch_t<float>(c, in, n, pd);
    else astats_launch_t<double>(c, in, n, pd);
}

void jt_astats(jt_ctx *c, const Sig &in, int64_t n_upto, AstatsResult &out)
{
    AstatsPending pd; jt_astats_launch(c, in, n_upto, pd); jt_astats_finish(c, pd, out);
}

void jt_astats_chunk_launch(jt_ctx *c, const Sig &in, int64_t own0, int64_t own_n, int64_t global_first, AstatsPending &pd)
{
    pd = AstatsPending();
    if (own_n <= 0 || own0 < 0 || own0 + own_n > in.n) return;
    if (in.fmt == JT_FMT_S16) astats_launch_t<int16_t>(c, in, own_n, pd, own0, global_first);
    else if (in.fmt == JT_FMT_S32) astats_launch_t<int32_t>(c, in, own_n, pd, own0, global_first);
    else if (in.fmt == JT_FMT_FLT) astats_launch_t<float>(c, in, own_n, pd, own0, global_first);
    else astats_launch_t<double>(c, in, own_n, pd, own0, global_first);
}
