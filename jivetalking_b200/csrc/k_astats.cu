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
k_astats_a(const T *__restrict__ x, int64_t n, AsPartA *__restrict__ parts, unsigned long long *__restrict__ ghist)
{
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
            while (j >= 0 && AsTraits<T>::d(x[j]) == 0) j--;
            const double prev = j >= 0 ? AsTraits<T>::d(x[j]) : 0.0;
            a.zero_runs += ((d > 0) != (prev > 0));
        }
        if (i > 0) {
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
k_astats_b(const T *__restrict__ x, int64_t n, double gmin, double gmax, unsigned long long *__restrict__ counts)
{
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

// C1: per block of BS samples, zero-state end value of avg = avg*mult + (1-mult)*nd^2
template <class T>
__global__ void k_astats_c1(const T *__restrict__ x, int64_t n, int BS, double mult, double *__restrict__ fin)
{
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t s = b * BS; if (s >= n) return;
    const int64_t e = min(s + BS, n);
    double avg = 0; const double om = 1.0 - mult;
    for (int64_t i = s; i < e; i++) { const double nd = AsTraits<T>::nd(x[i]); avg = avg * mult + om * nd * nd; }
    fin[b] = avg;
}
// carry[b] = state entering block b
__global__ void k_astats_carry(const double *__restrict__ fin, double *__restrict__ carry, int64_t nb, double mult_bs)
{
    if (blockIdx.x || threadIdx.x) return;
    double cy = 0;
    for (int64_t b = 0; b < nb; b++) { carry[b] = cy; cy = cy * mult_bs + fin[b]; }
}
template <class T>
__global__ void k_astats_c2(const T *__restrict__ x, int64_t n, int BS, double mult, const double *__restrict__ carry,
                            int64_t tc, double *__restrict__ mm /* [0]=min (init DBL_MAX), [1]=max */)
{
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t s = b * BS;
    double lo = DBL_MAX, hi = 0;
    if (s < n) {
        const int64_t e = min(s + BS, n);
        double avg = carry[b]; const double om = 1.0 - mult;
        for (int64_t i = s; i < e; i++) {
            const double nd = AsTraits<T>::nd(x[i]); avg = avg * mult + om * nd * nd;
            if (i >= tc) { lo = fmin(lo, avg); hi = fmax(hi, avg); }
        }
    }
    lo = jt_warp_min(lo); hi = jt_warp_max(hi);
    if ((threadIdx.x & 31) == 0) { jt_atomic_min_nonneg(&mm[0], lo); jt_atomic_max_nonneg(&mm[1], hi); }
}

// D1: per tc-sized block prefix / suffix maxima of |nd|
template <class T>
__global__ void k_astats_d1(const T *__restrict__ x, int64_t n, int tc, float *__restrict__ P, float *__restrict__ S)
{
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t s = b * tc; if (s >= n) return;
    const int64_t e = min(s + tc, n);
    float m = 0;
    for (int64_t i = s; i < e; i++) { m = fmaxf(m, (float)fabs(AsTraits<T>::nd(x[i]))); P[i] = m; }
    m = 0;
    for (int64_t i = e - 1; i >= s; i--) { m = fmaxf(m, (float)fabs(AsTraits<T>::nd(x[i]))); S[i] = m; }
}
__global__ void __launch_bounds__(256)
k_astats_d2(const float *__restrict__ P, const float *__restrict__ S, int64_t n, int tc, float *__restrict__ gmin)
{
    float lo = FLT_MAX;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = tc - 1 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        lo = fminf(lo, fmaxf(S[i - tc + 1], P[i]));
    for (int o = 16; o; o >>= 1) lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    if ((threadIdx.x & 31) == 0) atomicMin((unsigned int *)gmin, __float_as_uint(lo));
}
__global__ void __launch_bounds__(256)
k_astats_d3(const float *__restrict__ P, const float *__restrict__ S, int64_t n, int tc, const float *__restrict__ gmin,
            unsigned long long *__restrict__ cnt)
{
    const float g = *gmin; unsigned long long c = 0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = tc - 1 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        c += (fmaxf(S[i - tc + 1], P[i]) == g);
    for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(cnt, c);
}

template <class T>
static void astats_t(jt_ctx *c, const Sig &in, int64_t n, AstatsResult &out)
{
    const T *x = (const T *)in.d;
    const double time_constant = 0.05;
    const int tc = (int)std::fmax(time_constant * in.rate + .5, 1);
    const double mult = exp((-1 / time_constant / in.rate));
    const int maxbits = in.fmt == JT_FMT_S16 ? 16 : in.fmt == JT_FMT_FLT ? 32 : 64;

    const int gridA = jt_grid_for(n, 256, c->num_sms, 8);
    AsPartA *d_parts = jt_dalloc<AsPartA>(c, gridA);
    unsigned long long *d_hist = jt_dalloc<unsigned long long>(c, AS_HIST + 8);
    unsigned long long *d_counts = d_hist + AS_HIST;     // 4 counts + noise-floor count
    JT_CUDA(cudaMemsetAsync(d_hist, 0, sizeof(unsigned long long) * (AS_HIST + 8), c->stream));
    { JtLaunch L(c, "astats"); k_astats_a<T><<<gridA, 256, 0, c->stream>>>(x, n, d_parts, d_hist); }
    std::vector<AsPartA> parts(gridA);
    JT_CUDA(cudaMemcpyAsync(parts.data(), d_parts, sizeof(AsPartA) * gridA, cudaMemcpyDeviceToHost, c->stream));
    JT_CUDA(cudaStreamSynchronize(c->stream));
    AsPartA r = parts[0];
    for (int i = 1; i < gridA; i++) {
        const AsPartA &b = parts[i];
        r.sum += b.sum; r.sumsq += b.sumsq; r.diff_sum += b.diff_sum; r.diff_sumsq += b.diff_sumsq;
        r.mn = std::fmin(r.mn, b.mn); r.mx = std::fmax(r.mx, b.mx); r.min_nz = std::fmin(r.min_nz, b.min_nz);
        r.min_diff = std::fmin(r.min_diff, b.min_diff); r.max_diff = std::fmax(r.max_diff, b.max_diff);
        r.zero_runs += b.zero_runs; r.mask |= b.mask;
    }
    // B: extrema counts
    { JtLaunch L(c, "astats"); k_astats_b<T><<<gridA, 256, 0, c->stream>>>(x, n, r.mn, r.mx, d_counts); }
    // C: exponential mean square min/max
    double *d_mm = jt_dalloc<double>(c, 2);
    const int BS = 4096; const int64_t nb = (n + BS - 1) / BS;
    double h_mm[2] = {DBL_MAX, 0.0};
    JT_CUDA(cudaMemcpyAsync(d_mm, h_mm, sizeof(h_mm), cudaMemcpyHostToDevice, c->stream));
    if (n > tc) {
        double *d_fin = jt_dalloc<double>(c, nb), *d_carry = jt_dalloc<double>(c, nb);
        JtLaunch L(c, "astats", 3);
        k_astats_c1<T><<<(int)((nb + 63) / 64), 64, 0, c->stream>>>(x, n, BS, mult, d_fin);
        k_astats_carry<<<1, 1, 0, c->stream>>>(d_fin, d_carry, nb, pow(mult, (double)BS));
        k_astats_c2<T><<<(int)((nb + 63) / 64), 64, 0, c->stream>>>(x, n, BS, mult, d_carry, tc, d_mm);
    }
    // D: noise floor
    float *d_nf = jt_dalloc<float>(c, 1);
    float h_nf = FLT_MAX;
    JT_CUDA(cudaMemcpyAsync(d_nf, &h_nf, sizeof(float), cudaMemcpyHostToDevice, c->stream));
    if (n >= tc) {
        float *P = jt_dalloc<float>(c, n), *S = jt_dalloc<float>(c, n);
        const int64_t nbt = (n + tc - 1) / tc;
        const int gridD = jt_grid_for(n - tc + 1, 256, c->num_sms, 8);
        JtLaunch L(c, "astats", 3);
        k_astats_d1<T><<<(int)((nbt + 63) / 64), 64, 0, c->stream>>>(x, n, tc, P, S);
        k_astats_d2<<<gridD, 256, 0, c->stream>>>(P, S, n, tc, d_nf);
        k_astats_d3<<<gridD, 256, 0, c->stream>>>(P, S, n, tc, d_nf, d_counts + 4);
    }
    std::vector<unsigned long long> hist(AS_HIST + 8);
    JT_CUDA(cudaMemcpyAsync(hist.data(), d_hist, sizeof(unsigned long long) * (AS_HIST + 8), cudaMemcpyDeviceToHost, c->stream));
    JT_CUDA(cudaMemcpyAsync(h_mm, d_mm, sizeof(h_mm), cudaMemcpyDeviceToHost, c->stream));
    JT_CUDA(cudaMemcpyAsync(&h_nf, d_nf, sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    JT_CUDA(cudaStreamSynchronize(c->stream));

    const double N = (double)n;
    const double scale = in.fmt == JT_FMT_S16 ? 32767.0 : 1.0;
    const double nmin = r.mn / scale, nmax = r.mx / scale;
    double min_s2 = h_mm[0], max_s2 = h_mm[1];
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

void jt_astats(jt_ctx *c, const Sig &in, int64_t n_upto, AstatsResult &out)
{
    int64_t n = std::min(n_upto, in.n);
    for (int i = 0; i < JT_AS_COUNT; i++) out.v[i] = NAN;
    out.overall_rms = out.overall_peak = NAN; out.nb_samples = 0;
    if (n <= 0) return;
    if (in.fmt == JT_FMT_S16) astats_t<int16_t>(c, in, n, out);
    else if (in.fmt == JT_FMT_FLT) astats_t<float>(c, in, n, out);
    else astats_t<double>(c, in, n, out);
}
