// Polyphase resampler with libswresample's default design (kaiser-windowed sinc,
// filter_size 32, phase_shift 10, cutoff 0.97, exact_rational), used for
//   - ebur128's 192 kHz true-peak oversampler          (reference: filters.go:626)
//   - the aformat=sample_rates=44100 output stage       (reference: filters.go:706-710)
//   - loudnorm's forced 192 kHz input (dynamic mode)    (reference: normalise.go:257-264)
// Output m reads taps x[first_tap(m) .. +L) with swr's start mirror x[-k]=x[k] and end
// reflection x[n+j]=x[n-1-j].
//
// Mapping: one output "period" q holds phase_count outputs and consumes `div` inputs, so
// out[q*pc + r] = sum_i x[q*div - c + o_r + i] * F[ph_r][i] with o_r, ph_r depending on r
// only.  Small phase counts (48k->192k: pc = 4) run thread-per-period with the whole
// filter bank passed as a __grid_constant__ kernel parameter, so every DFMA takes its
// coefficient straight from the constant bank: 1 LDS + pc DFMA per tap.  Large phase
// counts (44.1k->192k: 640, 48k->44.1k: 147) run warp-per-r / lane-per-period so the
// coefficient load is warp-uniform and the x reads are a padded conflict-free stride.
#include "jt_internal.h"
#include "jt_device.cuh"
#include <cstdio>

// ---------------------------------------------------------------------------------------
// host: filter design (libswresample/resample.c build_filter, SWR_FILTER_TYPE_KAISER)
// ---------------------------------------------------------------------------------------
static double bessel_i0(double x)
{
    double y = x * x / 4.0, t = 1.0, sum = 1.0;
    for (int k = 1; k < 200; k++) { t *= y / ((double)k * k); sum += t; if (t < sum * 1e-18) break; }
    return sum;
}
static int64_t gcd_i64(int64_t a, int64_t b) { while (b) { int64_t t = a % b; a = b; b = t; } return a; }

static SwrPlan swr_plan_design(int in_rate, int out_rate);

// The design is a pure function of the two rates (~20k Bessel evaluations for 44.1k -> 192k):
// memoised process-wide, immutable once built.
#include <mutex>
SwrPlan jt_swr_plan(int in_rate, int out_rate)
{
    static std::mutex mu;
    static std::map<std::pair<int, int>, SwrPlan> cache;
    std::lock_guard<std::mutex> g(mu);
    auto it = cache.find({in_rate, out_rate});
    if (it == cache.end()) it = cache.emplace(std::make_pair(in_rate, out_rate), swr_plan_design(in_rate, out_rate)).first;
    return it->second;
}

static SwrPlan swr_plan_design(int in_rate, int out_rate)
{
    SwrPlan p; p.in_rate = in_rate; p.out_rate = out_rate;
    if (in_rate == out_rate) { p.identity = true; p.phase_count = 1; p.filter_length = 1; p.div = 1; return p; }
    const int filter_size = 32, phase_shift = 10;
    const double cutoff = 0.97, beta = 9.0;
    double factor = std::fmin((double)out_rate * cutoff / in_rate, 1.0);
    int64_t g = gcd_i64(out_rate, in_rate);
    int64_t pc = out_rate / g;
    // exact_rational: the reduced phase count when it fits 1 << phase_shift; otherwise all 1024 phases, a fractional index
    // advance and (linear_interp is on by default) resample_linear -- e.g. 22.05 kHz or 11.025 kHz sources into ebur128's
    // 192 kHz true-peak oversampler.  Pinned on the real library: tests/test_oracle_swr.py
    if (pc > (1 << phase_shift)) { pc = 1 << phase_shift; p.linear = true; }
    p.phase_count = (int)pc;
    p.filter_length = std::max((int)std::ceil(filter_size / factor), 1);
    {
        const int64_t num = (int64_t)in_rate * pc, den = out_rate, gg = gcd_i64(num, den);
        p.inc_num = num / gg; p.inc_den = den / gg;
    }
    p.div = (int)(p.inc_num / p.inc_den);
    const int L = p.filter_length, center = (L - 1) / 2;
    const int ph_nb = pc % 2 ? (int)pc : (int)pc / 2 + 1;
    p.bank.assign((size_t)(pc + 1) * L, 0.0);
    std::vector<double> tab(L), sin_lut(ph_nb, 0.0);
    double norm = 0;
    if (factor == 1.0)
        for (int ph = 0; ph < ph_nb; ph++) sin_lut[ph] = sin(M_PI * ph / pc) * (center & 1 ? 1 : -1);
    for (int ph = 0; ph < ph_nb; ph++) {
        double s = sin_lut[ph];
        for (int i = 0; i < L; i++) {
            double x = M_PI * ((double)(i - center) - (double)ph / pc) * factor, y;
            if (x == 0) y = 1.0; else if (factor == 1.0) y = s / x; else y = sin(x) / x;
            double w = 2.0 * x / (factor * L * M_PI);
            y *= bessel_i0(beta * sqrt(std::fmax(1 - w * w, 0)));
            tab[i] = y; s = -s;
            if (!ph) norm += y;
        }
        for (int i = 0; i < L; i++) p.bank[(size_t)ph * L + i] = tab[i] / norm;
        if (pc % 2 == 0)
            for (int i = 0; i < L; i++) p.bank[(size_t)(pc - ph) * L + L - 1 - i] = p.bank[(size_t)ph * L + i];
    }
    p.bank.resize((size_t)pc * L);
    if (p.linear) {
        // row phase_count = phase 0 one sample later (swri_resample_init: row 0 shifted by one tap inside its allocation of
        // FFALIGN(L, 8) slots, the first tap taken from row 0's last allocated slot)
        const int alloc = (L + 7) & ~7;
        p.bank.resize((size_t)(pc + 1) * L, 0.0);
        double *row = p.bank.data() + (size_t)pc * L;
        row[0] = alloc > L ? 0.0 : p.bank[L - 1];
        for (int i = 1; i < L; i++) row[i] = p.bank[i - 1];
    }
    return p;
}

static int64_t avail_outputs(const SwrPlan &p, int64_t n_avail)
{
    const int64_t pc = p.phase_count, L = p.filter_length, c = (L - 1) / 2;
    int64_t span = (1 + n_avail - L) * pc + pc * c;
    if (span <= 0) return 0;
    return (span * p.inc_den + p.inc_num - 1) / p.inc_num;          // outputs m with floor(m inc_num / inc_den) < span
}
int64_t SwrPlan::first_tap(int64_t m) const
{
    const int64_t pc = phase_count, c = (filter_length - 1) / 2;
    return jt_floordiv(-pc * c + jt_floordiv(m * inc_num, inc_den), pc);
}
int64_t SwrPlan::out_count(int64_t n_in) const
{
    if (identity) return n_in;
    if (n_in < filter_length + 1) return 0;
    return avail_outputs(*this, n_in);
}
int64_t SwrPlan::out_count_flush(int64_t n_in) const
{
    if (identity) return n_in;
    int64_t m0 = out_count(n_in);
    int64_t cnt = n_in - first_tap(m0);
    if (cnt > filter_length) cnt = filter_length;
    if (cnt < 0) cnt = 0;
    return avail_outputs(*this, n_in + (cnt + 1) / 2);
}

// ---------------------------------------------------------------------------------------
// device
// ---------------------------------------------------------------------------------------
template <class TIN, class TW> __device__ __forceinline__ TW swr_load(const TIN *x, int64_t idx, int64_t n)
{
    if (idx < 0) idx = -idx; else if (idx >= n) idx = 2 * n - 1 - idx;
    idx = idx < 0 ? 0 : (idx >= n ? n - 1 : idx);
    return (TW)jt_as_f64(x[idx]);
}
template <> __device__ __forceinline__ float swr_load<int16_t, float>(const int16_t *x, int64_t idx, int64_t n)
{
    if (idx < 0) idx = -idx; else if (idx >= n) idx = 2 * n - 1 - idx;
    idx = idx < 0 ? 0 : (idx >= n ? n - 1 : idx);
    return jt_conv<int16_t, float>(x[idx]);
}

template <int PC> struct SmallBank { double f[PC * 32]; };

enum { SWR_MODE_STORE = 0, SWR_MODE_TICKMAX = 1 };

// thread <-> TWO periods (2t, 2t+1), all PC phases of both in registers; L == 32, div == 1 (integer
// upsampling).  The two periods share 31 of their 32 taps' samples, so one shared-memory load feeds 2*PC DFMA
// instead of PC (the kernel was bound by 64-bit shared-memory loads as much as by the FP64 pipe); the tile is
// staged as even / odd samples so that the lanes of a load read consecutive addresses.
template <class TIN, int PC, int MODE>
__global__ void __launch_bounds__(256)
k_swr_small(const TIN *__restrict__ x, int64_t n, int64_t n_periods, int64_t n_out,
            const __grid_constant__ SmallBank<PC> bank, double *__restrict__ out,
            double *__restrict__ tick_max, int tick, int64_t n_ticks)
{
    constexpr int L = 32, C = (L - 1) / 2, TP = 512;              // periods per tile
    __shared__ double sev[TP / 2 + L], sod[TP / 2 + L];          // x[base + 2k], x[base + 2k + 1]
    const int64_t tiles = (n_periods + TP - 1) / TP;
    for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int64_t q0 = tile * TP;
        const int64_t base = q0 - C;                 // first tap of period q0
        __syncthreads();
        for (int i = threadIdx.x; i < TP + L; i += 256) {
            const double v = swr_load<TIN, double>(x, base + i, n);
            if (i & 1) sod[i >> 1] = v; else sev[i >> 1] = v;
        }
        __syncthreads();
        int64_t tile_k = 0, tile_lim = 0;
        if (MODE == SWR_MODE_TICKMAX) {
            int64_t nd0 = q0 - C + L; if (nd0 < L + 1) nd0 = L + 1;
            tile_k = (nd0 + tick - 1) / tick - 1; tile_lim = (tile_k + 1) * (int64_t)tick;
        }
        double acc[2][PC];
#pragma unroll
        for (int r = 0; r < PC; r++) { acc[0][r] = 0.0; acc[1][r] = 0.0; }
        // sample x[2t + i], i = 0 .. L: tap i of period 2t and tap i-1 of period 2t+1 (sums stay in tap order)
#pragma unroll
        for (int i = 0; i <= L; i++) {
            const double v = (i & 1) ? sod[threadIdx.x + (i >> 1)] : sev[threadIdx.x + (i >> 1)];
#pragma unroll
            for (int r = 0; r < PC; r++) {
                if (i < L) acc[0][r] = fma(v, bank.f[r * L + i], acc[0][r]);
                if (i >= 1) acc[1][r] = fma(v, bank.f[r * L + i - 1], acc[1][r]);
            }
        }
#pragma unroll
        for (int p = 0; p < 2; p++) {
            const int64_t q = q0 + 2 * threadIdx.x + p;
            if (MODE == SWR_MODE_STORE) {
                if (q < n_periods) {
#pragma unroll
                    for (int r = 0; r < PC; r++) { int64_t m = q * PC + r; if (m < n_out) out[m] = acc[p][r]; }
                }
            } else {
                double mx = 0.0; int64_t k = -1;
                if (q < n_periods) {
#pragma unroll
                    for (int r = 0; r < PC; r++) { int64_t m = q * PC + r; if (m < n_out) mx = fmax(mx, fabs(acc[p][r])); }
                    int64_t need = q - C + L; if (need < L + 1) need = L + 1;     // inputs swr must have seen
                    // one division per tile (warp-uniform operands), a compare per period: a tile spans 512 inputs, less than a tick
                    k = tick >= TP + L + 2 ? (need > tile_lim ? tile_k + 1 : tile_k) : (need + tick - 1) / tick - 1;
                }
                const int64_t k0 = __shfl_sync(0xffffffffu, k, 0);
                const bool uni = __all_sync(0xffffffffu, k == k0);
                if (uni) {
                    mx = jt_warp_max(mx);
                    if ((threadIdx.x & 31) == 0 && k0 >= 0 && k0 < n_ticks) jt_atomic_max_nonneg(&tick_max[k0], mx);
                } else if (k >= 0 && k < n_ticks) jt_atomic_max_nonneg(&tick_max[k], mx);
            }
        }
    }
}

// generic: lane <-> period, warp <-> r; x tile in padded shared memory, bank in global (uniform loads)
template <class TIN, class TW, int MODE>
__global__ void __launch_bounds__(256)
k_swr_generic(const TIN *__restrict__ x, int64_t n, int64_t n_periods, int64_t n_out,
              int pc, int L, int div, const TW *__restrict__ bank, TW *__restrict__ out,
              double *__restrict__ tick_max, int tick, int64_t n_ticks, int span)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    TW *sx = (TW *)smem_raw;
    const int c = (L - 1) / 2;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const int64_t tiles = (n_periods + 31) / 32;
    for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int64_t q0 = tile * 32;
        const int64_t base = q0 * div - c;
        __syncthreads();
        for (int i = threadIdx.x; i < span; i += blockDim.x) sx[i + (i >> 5)] = swr_load<TIN, TW>(x, base + i, n);
        __syncthreads();
        const int64_t q = q0 + lane;
        double cur_max = 0.0; int64_t cur_k = -1;
        for (int r = warp; r < pc; r += nwarp) {
            const int o_r = (int)(((int64_t)r * div) / pc);
            const int ph = (int)(((int64_t)r * div) % pc);
            const TW *f = bank + (size_t)ph * L;
            const int s0 = lane * div + o_r;
            TW a0 = 0, a1 = 0;
            int i = 0;
            for (; i + 1 < L; i += 2) {
                int ia = s0 + i, ib = ia + 1;
                a0 = fma(sx[ia + (ia >> 5)], __ldg(f + i), a0);
                a1 = fma(sx[ib + (ib >> 5)], __ldg(f + i + 1), a1);
            }
            if (i < L) { int ia = s0 + i; a0 = fma(sx[ia + (ia >> 5)], __ldg(f + i), a0); }
            const TW v = a0 + a1;
            const int64_t m = q * pc + r;
            if (q < n_periods && m < n_out) {
                if (MODE == SWR_MODE_STORE) out[m] = v;
                else {
                    int64_t need = q * div - c + o_r + L; if (need < L + 1) need = L + 1;
                    const int64_t k = (need + tick - 1) / tick - 1;
                    if (k != cur_k) {
                        if (cur_k >= 0 && cur_k < n_ticks) jt_atomic_max_nonneg(&tick_max[cur_k], cur_max);
                        cur_k = k; cur_max = 0.0;
                    }
                    cur_max = fmax(cur_max, fabs((double)v));
                }
            }
        }
        if (MODE == SWR_MODE_TICKMAX && cur_k >= 0 && cur_k < n_ticks) jt_atomic_max_nonneg(&tick_max[cur_k], cur_max);
    }
}


// Inexact ratios (SwrPlan::linear; resample_template.c resample_linear): output m reads from phase index
// -pc c + floor(m num / den), interpolating between that phase's row and the next by frac / den.  Thread per output over a
// staged tile; coefficient rows come through the L1 (1025 rows: the bank does not fit registers or a CTA's shared memory
// at f64).  A rare path (22.05 / 11.025 kHz sources, odd rates): written for clarity, not for the last cycle.
template <class TIN, class TW, int MODE>
__global__ void __launch_bounds__(256)
k_swr_linear(const TIN *__restrict__ x, int64_t n, int64_t n_out, int pc, int L, int64_t num, int64_t den,
             const TW *__restrict__ bank, TW *__restrict__ out, double *__restrict__ tick_max, int tick, int64_t n_ticks,
             int span, int tb)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    TW *sx = (TW *)smem_raw;
    const int c = (L - 1) / 2;
    const int64_t tiles = (n_out + tb - 1) / tb;
    for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int64_t m0 = tile * tb, m1 = min(m0 + tb, n_out);
        const int64_t base = jt_floordiv(-(int64_t)pc * c + (m0 * num) / den, (int64_t)pc);
        __syncthreads();
        for (int i = threadIdx.x; i < span; i += blockDim.x) sx[i] = swr_load<TIN, TW>(x, base + i, n);
        __syncthreads();
        double cur_max = 0.0; int64_t cur_k = -1;
        for (int64_t m = m0 + threadIdx.x; m < m1; m += blockDim.x) {
            const int64_t t = m * num, q = t / den, frac = t - q * den;
            const int64_t pos = -(int64_t)pc * c + q, sidx = jt_floordiv(pos, (int64_t)pc);
            const int ph = (int)(pos - sidx * pc);
            const TW *f = bank + (size_t)ph * L, *w = sx + (sidx - base);
            TW val = 0, v2 = 0;
            for (int i = 0; i < L; i++) { const TW s = w[i]; val = fma(s, __ldg(f + i), val); v2 = fma(s, __ldg(f + L + i), v2); }
            val += (v2 - val) * (TW)frac / (TW)den;
            if (MODE == SWR_MODE_STORE) out[m] = val;
            else {
                int64_t need = sidx + L; if (need < L + 1) need = L + 1;
                const int64_t k = (need + tick - 1) / tick - 1;
                if (k != cur_k) {
                    if (cur_k >= 0 && cur_k < n_ticks) jt_atomic_max_nonneg(&tick_max[cur_k], cur_max);
                    cur_k = k; cur_max = 0.0;
                }
                cur_max = fmax(cur_max, fabs((double)val));
            }
        }
        if (MODE == SWR_MODE_TICKMAX && cur_k >= 0 && cur_k < n_ticks) jt_atomic_max_nonneg(&tick_max[cur_k], cur_max);
    }
}

// ---------------------------------------------------------------------------------------
// f32-internal path, coefficients in registers ("slot" kernel).
// Phases that read the same input window form a slot: when upsampling (pc > div) slot s = o holds the
// R = 4..5 phases r with floor(r*div/pc) == o; when downsampling every phase is its own slot (R = 1).
// Thread <-> slot for the whole launch: its R*L coefficients stay in registers, and for every period q
// it reads the window x[q*div - c + o .. +L) from the CTA's staged tile (consecutive slots read
// consecutive addresses: conflict-free) -- 1 LDS per R FFMA.  Outputs of QB consecutive periods are
// one contiguous run out[q*pc .. (q+QB)*pc): they are collected in shared memory and leave as full
// coalesced rows, converted to the link format on the way (s16 for the 44.1 kHz output stage).
// ---------------------------------------------------------------------------------------
template <class TIN, class TOUT, int L, int R, int QB>
__global__ void __launch_bounds__(160, 2)
k_swr_slot_f32(const TIN *__restrict__ x, int64_t n, int64_t n_periods, int64_t n_out, int pc, int div, int nslots, int qc,
               const float *__restrict__ bank, TOUT *__restrict__ out)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int c = (L - 1) / 2;
    const int span = qc * div + L + div;
    float *sx = (float *)smem_raw;                       // span floats
    float *so = sx + ((span + 3) & ~3);                  // QB * pc floats
    const int s = threadIdx.x;
    const bool live = s < nslots;
    // this slot's phases: r in [rlo, rlo + cnt), all with input offset `off`
    int rlo = 0, cnt = 0, off = 0;
    if (live) {
        if (R == 1) { rlo = s; cnt = 1; off = (int)(((int64_t)s * div) / pc); }
        else { rlo = (int)(((int64_t)s * pc + div - 1) / div); const int rhi = (int)(((int64_t)(s + 1) * pc + div - 1) / div); cnt = min(rhi, pc) - rlo; off = s; }
    }
    float cf[R][L];
#pragma unroll
    for (int j = 0; j < R; j++) {
        const int r = rlo + (j < cnt ? j : 0);
        const int ph = (int)(((int64_t)r * div) % pc);
#pragma unroll
        for (int i = 0; i < L; i++) cf[j][i] = (live && j < cnt) ? bank[(size_t)ph * L + i] : 0.f;
    }
    const int64_t chunks = (n_periods + qc - 1) / qc;
    for (int64_t ch = blockIdx.x; ch < chunks; ch += gridDim.x) {
        const int64_t qa = ch * qc;
        const int nq = (int)min((int64_t)qc, n_periods - qa);
        const int64_t base = qa * div - c;
        __syncthreads();
        for (int i = threadIdx.x; i < span; i += blockDim.x) sx[i] = swr_load<TIN, float>(x, base + i, n);
        __syncthreads();
        for (int q0 = 0; q0 < nq; q0 += QB) {
            float acc[QB][R];
#pragma unroll
            for (int b = 0; b < QB; b++)
#pragma unroll
                for (int j = 0; j < R; j++) acc[b][j] = 0.f;
            if (live) {
                const float *w = sx + q0 * div + off;
#pragma unroll
                for (int i = 0; i < L; i++) {
#pragma unroll
                    for (int b = 0; b < QB; b++) {
                        const float v = w[b * div + i];          // periods past nq read staged slack or stale data: never stored
#pragma unroll
                        for (int j = 0; j < R; j++) acc[b][j] = fmaf(v, cf[j][i], acc[b][j]);
                    }
                }
#pragma unroll
                for (int b = 0; b < QB; b++)
#pragma unroll
                    for (int j = 0; j < R; j++) if (j < cnt) so[b * pc + rlo + j] = acc[b][j];
            }
            __syncthreads();
            const int nb = min(QB, nq - q0);
            const int64_t m0 = (qa + q0) * (int64_t)pc;
            const int64_t lim = min((int64_t)nb * pc, n_out - m0);
            for (int i = threadIdx.x; i < lim; i += blockDim.x) out[m0 + i] = jt_conv<float, TOUT>(so[i]);
            __syncthreads();
        }
    }
}

// ---------------------------------------------------------------------------------------
// f64 path ("q-lane" kernel): lane <-> period q (two periods per lane, 32 apart), warp <-> slot.
// The R phases of a slot share each x load, coefficient loads are warp-uniform 16-byte loads.
// ---------------------------------------------------------------------------------------
template <class TIN, int MODE, int RMAX>
__global__ void __launch_bounds__(256)
k_swr_qlane_f64(const TIN *__restrict__ x, int64_t n, int64_t n_periods, int64_t n_out,
                int pc, int L, int div, int nslots, const double *__restrict__ bank, double *__restrict__ out,
                double *__restrict__ tick_max, int tick, int64_t n_ticks, int span)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *sx = (double *)smem_raw;
    const int c = (L - 1) / 2;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const int64_t tiles = (n_periods + 63) / 64;
    for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int64_t q0 = tile * 64;
        const int64_t base = q0 * div - c;
        __syncthreads();
        for (int i = threadIdx.x; i < span; i += blockDim.x) sx[i + (i >> 5)] = swr_load<TIN, double>(x, base + i, n);
        __syncthreads();
        double cur_max[2] = {0.0, 0.0}; int64_t cur_k[2] = {-1, -1};
        for (int s = warp; s < nslots; s += nwarp) {
            int rlo, cnt, off;
            if (RMAX == 1) { rlo = s; cnt = 1; off = (int)(((int64_t)s * div) / pc); }
            else { rlo = (int)(((int64_t)s * pc + div - 1) / div); const int rhi = (int)(((int64_t)(s + 1) * pc + div - 1) / div); cnt = min(rhi, pc) - rlo; off = s; }
            const double *f[RMAX];
#pragma unroll
            for (int j = 0; j < RMAX; j++) { const int r = rlo + (j < cnt ? j : 0); f[j] = bank + (size_t)(((int64_t)r * div) % pc) * L; }
            double acc[2][RMAX];
#pragma unroll
            for (int t = 0; t < 2; t++)
#pragma unroll
                for (int j = 0; j < RMAX; j++) acc[t][j] = 0.0;
            const int sA = lane * div + off, sB = sA + 32 * div;
            // eight taps per trip: the warp-uniform coefficient loads of a trip are issued together, ahead of the FMAs
#pragma unroll 4
            for (int i = 0; i < L; i += 2) {
                const int a0 = sA + i, a1 = a0 + 1, b0 = sB + i, b1 = b0 + 1;
                const double xa0 = sx[a0 + (a0 >> 5)], xa1 = sx[a1 + (a1 >> 5)], xb0 = sx[b0 + (b0 >> 5)], xb1 = sx[b1 + (b1 >> 5)];
#pragma unroll
                for (int j = 0; j < RMAX; j++) {
                    if (j < cnt) {
                        const double2 cc = __ldg((const double2 *)(f[j] + i));
                        acc[0][j] = fma(xa0, cc.x, acc[0][j]); acc[1][j] = fma(xb0, cc.x, acc[1][j]);
                        acc[0][j] = fma(xa1, cc.y, acc[0][j]); acc[1][j] = fma(xb1, cc.y, acc[1][j]);
                    }
                }
            }
#pragma unroll
            for (int t = 0; t < 2; t++) {
                const int64_t q = q0 + lane + 32 * t;
                if (q >= n_periods) continue;
                if (MODE == SWR_MODE_STORE) {
#pragma unroll
                    for (int j = 0; j < RMAX; j++) { const int64_t m = q * pc + rlo + j; if (j < cnt && m < n_out) out[m] = acc[t][j]; }
                } else {
                    double mx = 0.0; bool any = false;
#pragma unroll
                    for (int j = 0; j < RMAX; j++) { const int64_t m = q * pc + rlo + j; if (j < cnt && m < n_out) { mx = fmax(mx, fabs(acc[t][j])); any = true; } }
                    if (any) {
                        int64_t need = q * div - c + off + L; if (need < L + 1) need = L + 1;
                        const int64_t k = (need + tick - 1) / tick - 1;
                        if (k != cur_k[t]) {
                            if (cur_k[t] >= 0 && cur_k[t] < n_ticks) jt_atomic_max_nonneg(&tick_max[cur_k[t]], cur_max[t]);
                            cur_k[t] = k; cur_max[t] = 0.0;
                        }
                        cur_max[t] = fmax(cur_max[t], mx);
                    }
                }
            }
        }
        if (MODE == SWR_MODE_TICKMAX) {
#pragma unroll
            for (int t = 0; t < 2; t++) if (cur_k[t] >= 0 && cur_k[t] < n_ticks) jt_atomic_max_nonneg(&tick_max[cur_k[t]], cur_max[t]);
        }
    }
}

// ---------------------------------------------------------------------------------------
// f64 path, one PHASE per thread with its L coefficients in registers (64-72 registers), looping over the
// periods of a staged tile: 1 shared-memory load per DFMA and no coefficient traffic at all.  The q-lane kernel
// above re-fetches every coefficient row through the L1 for every tile and was bound by that latency (ncu:
// long-scoreboard 5.4 per issue, FP64 pipe 7 %).  Consecutive phases read (almost) the same window, so the
// lanes of a load hit a handful of addresses: conflict-free.  STORE collects QB periods in shared memory and
// writes full rows; TICKMAX keeps a per-thread running maximum per 100 ms tick.
// ---------------------------------------------------------------------------------------
template <class TW> static const TW *device_bank(jt_ctx *c, const SwrPlan &p);

template <class TIN, int L, int MODE, int NT>
__global__ void __launch_bounds__(NT, NT > 320 ? 1 : 2)
k_swr_phase_f64(const TIN *__restrict__ x, int64_t n, int64_t n_periods, int64_t n_out, int pc, int div, int qc,
                const double *__restrict__ bank, double *__restrict__ out,
                double *__restrict__ tick_max, int tick, int64_t n_ticks)
{
    constexpr int QB = 2;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int c = (L - 1) / 2;
    const int span = qc * div + L + div;
    double *sx = (double *)smem_raw;                     // span doubles
    double *so = sx + span;                              // QB * pc doubles (STORE)
    const int r = threadIdx.x;
    const bool live = r < pc;
    const int off = live ? (int)(((int64_t)r * div) / pc) : 0;
    double cf[L];
    {
        const int ph = live ? (int)(((int64_t)r * div) % pc) : 0;
#pragma unroll
        for (int i = 0; i < L; i++) cf[i] = live ? bank[(size_t)ph * L + i] : 0.0;
    }
    const int64_t chunks = (n_periods + qc - 1) / qc;
    for (int64_t ch = blockIdx.x; ch < chunks; ch += gridDim.x) {
        const int64_t qa = ch * qc;
        const int nq = (int)min((int64_t)qc, n_periods - qa);
        const int64_t base = qa * div - c;
        __syncthreads();
        for (int i = threadIdx.x; i < span; i += NT) sx[i] = swr_load<TIN, double>(x, base + i, n);
        __syncthreads();
        // tick of an output = ceil(need / tick) - 1 with need = inputs swr must have seen: it grows by div per period,
        // so it is tracked with a compare per output instead of a 64-bit division (which cost more than the 32 DFMAs)
        double cur_max = 0.0; int64_t cur_k = -1;
        int64_t need = qa * div - c + off + L, lim = 0;
        if (MODE == SWR_MODE_TICKMAX) {
            const int64_t nd = need < L + 1 ? L + 1 : need;
            cur_k = (nd + tick - 1) / tick - 1; lim = (cur_k + 1) * (int64_t)tick;
        }
        for (int q0 = 0; q0 < nq; q0 += QB) {
            double acc[QB];
#pragma unroll
            for (int b = 0; b < QB; b++) acc[b] = 0.0;
            const double *w = sx + q0 * div + off;
#pragma unroll
            for (int i = 0; i < L; i++) {
#pragma unroll
                for (int b = 0; b < QB; b++) acc[b] = fma(w[b * div + i], cf[i], acc[b]);      // periods past nq read slack: never used
                if ((i & 7) == 7) asm volatile("" ::: "memory");     // keep at most 8 taps of loads in flight: the registers hold the coefficients
            }
            if (MODE == SWR_MODE_STORE) {
                if (live) {
#pragma unroll
                    for (int b = 0; b < QB; b++) so[b * pc + r] = acc[b];
                }
                __syncthreads();
                const int nb = min(QB, nq - q0);
                const int64_t m0 = (qa + q0) * (int64_t)pc;
                const int64_t lim = min((int64_t)nb * pc, n_out - m0);
                for (int i = threadIdx.x; i < lim; i += NT) out[m0 + i] = so[i];
                __syncthreads();
            } else if (live) {
#pragma unroll
                for (int b = 0; b < QB; b++) {
                    const int64_t q = qa + q0 + b, m = q * pc + r;
                    if (q0 + b < nq && m < n_out) {
                        const int64_t nd = need + (int64_t)(q0 + b) * div;           // (values below L+1 clamp to L+1: same tick 0)
                        while (nd > lim) {
                            if (cur_k >= 0 && cur_k < n_ticks && cur_max > 0.0) jt_atomic_max_nonneg(&tick_max[cur_k], cur_max);
                            cur_k++; lim += tick; cur_max = 0.0;
                        }
                        cur_max = fmax(cur_max, fabs(acc[b]));
                    }
                }
            }
        }
        if (MODE == SWR_MODE_TICKMAX && cur_k >= 0 && cur_k < n_ticks) jt_atomic_max_nonneg(&tick_max[cur_k], cur_max);
    }
}

static bool phase_path_ok(const SwrPlan &p)
{
    if (p.linear) return false;
    const int L = p.filter_length;
    if (!(L == 32 || L == 36 || L == 72)) return false;
    return p.phase_count >= 32 && p.phase_count <= 640 && p.div >= 2;
}

template <class TIN, int MODE>
static void launch_phase_f64(jt_ctx *c, const Sig &in, const SwrPlan &p, int64_t n_out, double *out,
                             double *tick_max, int tick, int64_t n_ticks)
{
    const double *d_bank = device_bank<double>(c, p);
    const int pc = p.phase_count, div = p.div, L = p.filter_length;
    const int64_t n_periods = (n_out + pc - 1) / pc;
    int qc = std::max(2, std::min(64, (10 * 1024) / div / 2 * 2));        // periods per staged tile (~80 KB of doubles at most)
    const int span = qc * div + L + div;
    const size_t smem = sizeof(double) * ((size_t)span + (MODE == SWR_MODE_STORE ? 2 * (size_t)pc : 0)) + 16;
    const int grid = jt_grid_for((n_periods + qc - 1) / qc, 1, c->num_sms, 8);
#define PHASE_LAUNCH(LV, NTV) do { auto kfn = k_swr_phase_f64<TIN, LV, MODE, NTV>; \
        jt_smem_optin((const void *)kfn, (size_t)(smem)); \
        kfn<<<grid, NTV, smem, c->stream>>>((const TIN *)in.d, in.n, n_periods, n_out, pc, div, qc, d_bank, out, tick_max, tick, n_ticks); } while (0)
    if (pc <= 160) { if (L == 32) PHASE_LAUNCH(32, 160); else if (L == 36) PHASE_LAUNCH(36, 160); else PHASE_LAUNCH(72, 160); }
    else { if (L == 32) PHASE_LAUNCH(32, 640); else if (L == 36) PHASE_LAUNCH(36, 640); else PHASE_LAUNCH(72, 640); }
#undef PHASE_LAUNCH
}

// device copy of the filter bank in the work format, cached per context (the plan is immutable)
template <class TW>
static const TW *device_bank(jt_ctx *c, const SwrPlan &p)
{
    const size_t nb = p.bank.size();                    // phase_count (+ 1 when linear) rows
    std::vector<TW> hb(nb);
    for (size_t i = 0; i < nb; i++) hb[i] = (TW)p.bank[i];
    return jt_dev_table(c, sizeof(TW) == 4 ? "swr_bank_f32" : "swr_bank_f64", hb);
}

static bool slot_path_ok(const SwrPlan &p)
{
    if (p.linear) return false;
    const int L = p.filter_length;
    const bool up = p.phase_count > p.div;
    const int nslots = up ? p.div : p.phase_count;
    if (nslots < 32 || nslots > 160) return false;
    if (up) return L == 32 && (p.phase_count + p.div - 1) / p.div <= 5;
    return L == 36 || L == 72;
}

template <class TIN, class TOUT>
static void launch_slot_f32(jt_ctx *c, const Sig &in, const SwrPlan &p, int64_t n_out, TOUT *out)
{
    const float *d_bank = device_bank<float>(c, p);
    const int pc = p.phase_count, div = p.div, L = p.filter_length;
    const bool up = pc > div;
    const int nslots = up ? div : pc;
    const int64_t n_periods = (n_out + pc - 1) / pc;
    constexpr int QB = 2;
    int qc = std::max(QB, std::min(64, (40 * 1024) / div / QB * QB));      // periods staged per chunk (~160 KB of floats at most)
    const int span = qc * div + L + div;
    const size_t smem = sizeof(float) * (((size_t)span + 3) / 4 * 4 + (size_t)QB * pc) + 16;
    const int grid = jt_grid_for((n_periods + qc - 1) / qc, 1, c->num_sms, 8);
#define SLOT_LAUNCH(LV, RV) do { auto kfn = k_swr_slot_f32<TIN, TOUT, LV, RV, QB>; \
        jt_smem_optin((const void *)kfn, (size_t)(smem)); \
        kfn<<<grid, 160, smem, c->stream>>>((const TIN *)in.d, in.n, n_periods, n_out, pc, div, nslots, qc, d_bank, out); } while (0)
    if (up) SLOT_LAUNCH(32, 5);
    else if (L == 36) SLOT_LAUNCH(36, 1);
    else SLOT_LAUNCH(72, 1);
#undef SLOT_LAUNCH
}

template <class TIN, int MODE>
static void launch_qlane_f64(jt_ctx *c, const Sig &in, const SwrPlan &p, int64_t n_out, double *out,
                             double *tick_max, int tick, int64_t n_ticks)
{
    const double *d_bank = device_bank<double>(c, p);
    const int pc = p.phase_count, div = p.div, L = p.filter_length;
    const bool up = pc > div;
    const int nslots = up ? div : pc;
    const int64_t n_periods = (n_out + pc - 1) / pc;
    const int span = 64 * div + L + div;
    const size_t smem = (size_t)(span + (span >> 5) + 2) * sizeof(double);
    const int grid = jt_grid_for((n_periods + 63) / 64, 1, c->num_sms, 8);
    if (up) {
        auto kfn = k_swr_qlane_f64<TIN, MODE, 5>;
        jt_smem_optin((const void *)kfn, (size_t)(smem));
        kfn<<<grid, 256, smem, c->stream>>>((const TIN *)in.d, in.n, n_periods, n_out, pc, L, div, nslots, d_bank, out, tick_max, tick, n_ticks, span);
    } else {
        auto kfn = k_swr_qlane_f64<TIN, MODE, 1>;
        jt_smem_optin((const void *)kfn, (size_t)(smem));
        kfn<<<grid, 256, smem, c->stream>>>((const TIN *)in.d, in.n, n_periods, n_out, pc, L, div, nslots, d_bank, out, tick_max, tick, n_ticks, span);
    }
}
static bool qlane_path_ok(const SwrPlan &p)
{
    if (p.linear) return false;
    const bool up = p.phase_count > p.div;
    if (p.filter_length % 2) return false;
    if (up && (p.phase_count + p.div - 1) / p.div > 5) return false;
    const size_t span = 64 * (size_t)p.div + p.filter_length + p.div;
    return (span + (span >> 5) + 2) * sizeof(double) <= 200 * 1024;
}

template <class TIN, int MODE>
static void launch_small(jt_ctx *c, const Sig &in, const SwrPlan &p, int64_t n_out, double *out,
                         double *tick_max, int tick, int64_t n_ticks)
{
    const int64_t n_periods = (n_out + p.phase_count - 1) / p.phase_count;
    const int grid = jt_grid_for(n_periods, 512, c->num_sms, 16);
#define SMALL_CASE(PCV) case PCV: { SmallBank<PCV> b; for (int i = 0; i < PCV * 32; i++) b.f[i] = p.bank[i]; \
        k_swr_small<TIN, PCV, MODE><<<grid, 256, 0, c->stream>>>((const TIN *)in.d, in.n, n_periods, n_out, b, out, tick_max, tick, n_ticks); } break;
    switch (p.phase_count) { SMALL_CASE(2) SMALL_CASE(4) SMALL_CASE(6) SMALL_CASE(8) default: JT_THROW(JT_ERR_UNSUPPORTED, "small pc"); }
#undef SMALL_CASE
}

template <class TIN, class TW, int MODE>
static void launch_generic(jt_ctx *c, const Sig &in, const SwrPlan &p, int64_t n_out, TW *out,
                           double *tick_max, int tick, int64_t n_ticks)
{
    const TW *d_bank = device_bank<TW>(c, p);
    const int64_t n_periods = (n_out + p.phase_count - 1) / p.phase_count;
    const int span = 32 * p.div + p.filter_length + p.div;
    const size_t smem = (size_t)(span + (span >> 5) + 2) * sizeof(TW);
    auto kfn = k_swr_generic<TIN, TW, MODE>;
    jt_smem_optin((const void *)kfn, (size_t)(smem));
    const int grid = jt_grid_for((n_periods + 31) / 32, 1, c->num_sms, 8);
    kfn<<<grid, 256, smem, c->stream>>>((const TIN *)in.d, in.n, n_periods, n_out, p.phase_count, p.filter_length, p.div,
                                          d_bank, out, tick_max, tick, n_ticks, span);
}

template <class TIN, class TW, int MODE>
static void launch_linear(jt_ctx *c, const Sig &in, const SwrPlan &p, int64_t n_out, TW *out,
                          double *tick_max, int tick, int64_t n_ticks)
{
    const TW *d_bank = device_bank<TW>(c, p);
    const int tb = 2048;
    const int span = (int)(((int64_t)tb * p.inc_num / p.inc_den) / p.phase_count) + p.filter_length + 4;
    const size_t smem = (size_t)span * sizeof(TW) + 16;
    if (smem > 200 * 1024) JT_THROW(JT_ERR_UNSUPPORTED, "resample %d -> %d: tile of %d inputs", p.in_rate, p.out_rate, span);
    auto kfn = k_swr_linear<TIN, TW, MODE>;
    jt_smem_optin((const void *)kfn, smem);
    const int grid = jt_grid_for((n_out + tb - 1) / tb, 1, c->num_sms, 8);
    kfn<<<grid, 256, smem, c->stream>>>((const TIN *)in.d, in.n, n_out, p.phase_count, p.filter_length, p.inc_num, p.inc_den,
                                          d_bank, out, tick_max, tick, n_ticks, span, tb);
}

static bool small_path(const SwrPlan &p) { return !p.linear && p.div == 1 && p.filter_length == 32 && (p.phase_count == 2 || p.phase_count == 4 || p.phase_count == 6 || p.phase_count == 8); }

Sig jt_swr_resample(jt_ctx *c, const Sig &in0, const SwrPlan &p, int work_fmt, bool flush, int fuse_out_fmt)
{
    if (p.identity) return jt_convert(c, in0, work_fmt);
    // 32-bit integer input: swr converts to its internal format (flt for s32: swr_init's int_sample_fmt rule) before anything else
    const Sig in = in0.fmt == JT_FMT_S32 ? jt_convert(c, in0, work_fmt) : in0;
    const int64_t n_out = flush ? p.out_count_flush(in.n) : p.out_count(in.n);
    // the f32 slot kernel converts on store: the 44.1 kHz / s16 output stage never materialises its float stream
    const bool fuse_s16 = work_fmt == JT_FMT_FLT && fuse_out_fmt == JT_FMT_S16 && slot_path_ok(p) && in.fmt != JT_FMT_DBL;
    Sig o; o.fmt = fuse_s16 ? JT_FMT_S16 : work_fmt; o.rate = p.out_rate; o.n = n_out;
    o.d = jt_dalloc_bytes(c, (size_t)std::max<int64_t>(n_out, 1) * jt_fmt_bytes(o.fmt));
    if (n_out <= 0) return o;
    const char *kind = work_fmt == JT_FMT_DBL ? (small_path(p) ? "swr_resample:small_f64" : phase_path_ok(p) ? "swr_resample:phase_f64" : qlane_path_ok(p) ? "swr_resample:qlane_f64" : "swr_resample:generic")
                                              : (slot_path_ok(p) ? (p.phase_count > p.div ? "swr_resample:slot_f32_up" : "swr_resample:slot_f32_down") : "swr_resample:generic");
    JtLaunch Lc(c, p.linear ? "swr_resample:linear" : kind);
    if (p.linear) {
        if (work_fmt == JT_FMT_DBL) {
            if (in.fmt == JT_FMT_S16) launch_linear<int16_t, double, SWR_MODE_STORE>(c, in, p, n_out, (double *)o.d, nullptr, 1, 0);
            else if (in.fmt == JT_FMT_FLT) launch_linear<float, double, SWR_MODE_STORE>(c, in, p, n_out, (double *)o.d, nullptr, 1, 0);
            else launch_linear<double, double, SWR_MODE_STORE>(c, in, p, n_out, (double *)o.d, nullptr, 1, 0);
        } else if (work_fmt == JT_FMT_FLT) {
            if (in.fmt == JT_FMT_DBL) JT_THROW(JT_ERR_UNSUPPORTED, "f64 -> f32-internal resample");
            if (in.fmt == JT_FMT_S16) launch_linear<int16_t, float, SWR_MODE_STORE>(c, in, p, n_out, (float *)o.d, nullptr, 1, 0);
            else launch_linear<float, float, SWR_MODE_STORE>(c, in, p, n_out, (float *)o.d, nullptr, 1, 0);
        } else JT_THROW(JT_ERR_UNSUPPORTED, "swr work format %d", work_fmt);
        return o;
    }
    if (work_fmt == JT_FMT_DBL) {
        if (small_path(p)) {
            if (in.fmt == JT_FMT_S16) launch_small<int16_t, SWR_MODE_STORE>(c, in, p, n_out, (double *)o.d, nullptr, 1, 0);
            else if (in.fmt == JT_FMT_FLT) launch_small<float, SWR_MODE_STORE>(c, in, p, n_out, (double *)o.d, nullptr, 1, 0);
            else launch_small<double, SWR_MODE_STORE>(c, in, p, n_out, (double *)o.d, nullptr, 1, 0);
        } else if (phase_path_ok(p)) {
            if (in.fmt == JT_FMT_S16) launch_phase_f64<int16_t, SWR_MODE_STORE>(c, in, p, n_out, (double *)o.d, nullptr, 1, 0);
            else if (in.fmt == JT_FMT_FLT) launch_phase_f64<float, SWR_MODE_STORE>(c, in, p, n_out, (double *)o.d, nullptr, 1, 0);
            else launch_phase_f64<double, SWR_MODE_STORE>(c, in, p, n_out, (double *)o.d, nullptr, 1, 0);
        } else if (qlane_path_ok(p)) {
            if (in.fmt == JT_FMT_S16) launch_qlane_f64<int16_t, SWR_MODE_STORE>(c, in, p, n_out, (double *)o.d, nullptr, 1, 0);
            else if (in.fmt == JT_FMT_FLT) launch_qlane_f64<float, SWR_MODE_STORE>(c, in, p, n_out, (double *)o.d, nullptr, 1, 0);
            else launch_qlane_f64<double, SWR_MODE_STORE>(c, in, p, n_out, (double *)o.d, nullptr, 1, 0);
        } else {
            if (in.fmt == JT_FMT_S16) launch_generic<int16_t, double, SWR_MODE_STORE>(c, in, p, n_out, (double *)o.d, nullptr, 1, 0);
            else if (in.fmt == JT_FMT_FLT) launch_generic<float, double, SWR_MODE_STORE>(c, in, p, n_out, (double *)o.d, nullptr, 1, 0);
            else launch_generic<double, double, SWR_MODE_STORE>(c, in, p, n_out, (double *)o.d, nullptr, 1, 0);
        }
    } else if (work_fmt == JT_FMT_FLT) {
        if (in.fmt == JT_FMT_DBL) JT_THROW(JT_ERR_UNSUPPORTED, "f64 -> f32-internal resample");
        if (slot_path_ok(p)) {
            if (fuse_s16) {
                if (in.fmt == JT_FMT_S16) launch_slot_f32<int16_t, int16_t>(c, in, p, n_out, (int16_t *)o.d);
                else launch_slot_f32<float, int16_t>(c, in, p, n_out, (int16_t *)o.d);
            } else {
                if (in.fmt == JT_FMT_S16) launch_slot_f32<int16_t, float>(c, in, p, n_out, (float *)o.d);
                else launch_slot_f32<float, float>(c, in, p, n_out, (float *)o.d);
            }
        } else {
            if (in.fmt == JT_FMT_S16) launch_generic<int16_t, float, SWR_MODE_STORE>(c, in, p, n_out, (float *)o.d, nullptr, 1, 0);
            else launch_generic<float, float, SWR_MODE_STORE>(c, in, p, n_out, (float *)o.d, nullptr, 1, 0);
        }
    } else JT_THROW(JT_ERR_UNSUPPORTED, "swr work format %d", work_fmt);
    return o;
}

void jt_swr_tick_absmax(jt_ctx *c, const Sig &in, const SwrPlan &p, int tick, int64_t n_ticks, double *d_tick_tp)
{
    JT_CUDA(cudaMemsetAsync(d_tick_tp, 0, sizeof(double) * (size_t)std::max<int64_t>(n_ticks, 1), c->stream));
    if (n_ticks <= 0) return;
    // outputs swr has produced once n_ticks*tick inputs were fed (ebur128 never flushes)
    const int64_t fed = n_ticks * (int64_t)tick;
    Sig v = in.fmt == JT_FMT_S32 ? jt_convert(c, in, JT_FMT_DBL) : in; v.n = std::min(in.n, fed);
    const int64_t n_out = p.out_count(v.n);
    if (n_out <= 0) return;
    JtLaunch Lc(c, p.linear ? "truepeak_oversample:linear" : small_path(p) ? "truepeak_oversample:small_f64" : phase_path_ok(p) ? "truepeak_oversample:phase_f64" : qlane_path_ok(p) ? "truepeak_oversample:qlane_f64" : "truepeak_oversample:generic");
    if (p.linear) {
        if (v.fmt == JT_FMT_S16) launch_linear<int16_t, double, SWR_MODE_TICKMAX>(c, v, p, n_out, (double *)nullptr, d_tick_tp, tick, n_ticks);
        else if (v.fmt == JT_FMT_FLT) launch_linear<float, double, SWR_MODE_TICKMAX>(c, v, p, n_out, (double *)nullptr, d_tick_tp, tick, n_ticks);
        else launch_linear<double, double, SWR_MODE_TICKMAX>(c, v, p, n_out, (double *)nullptr, d_tick_tp, tick, n_ticks);
    } else if (small_path(p)) {
        if (v.fmt == JT_FMT_S16) launch_small<int16_t, SWR_MODE_TICKMAX>(c, v, p, n_out, nullptr, d_tick_tp, tick, n_ticks);
        else if (v.fmt == JT_FMT_FLT) launch_small<float, SWR_MODE_TICKMAX>(c, v, p, n_out, nullptr, d_tick_tp, tick, n_ticks);
        else launch_small<double, SWR_MODE_TICKMAX>(c, v, p, n_out, nullptr, d_tick_tp, tick, n_ticks);
    } else if (phase_path_ok(p)) {
        if (v.fmt == JT_FMT_S16) launch_phase_f64<int16_t, SWR_MODE_TICKMAX>(c, v, p, n_out, nullptr, d_tick_tp, tick, n_ticks);
        else if (v.fmt == JT_FMT_FLT) launch_phase_f64<float, SWR_MODE_TICKMAX>(c, v, p, n_out, nullptr, d_tick_tp, tick, n_ticks);
        else launch_phase_f64<double, SWR_MODE_TICKMAX>(c, v, p, n_out, nullptr, d_tick_tp, tick, n_ticks);
    } else if (qlane_path_ok(p)) {
        if (v.fmt == JT_FMT_S16) launch_qlane_f64<int16_t, SWR_MODE_TICKMAX>(c, v, p, n_out, nullptr, d_tick_tp, tick, n_ticks);
        else if (v.fmt == JT_FMT_FLT) launch_qlane_f64<float, SWR_MODE_TICKMAX>(c, v, p, n_out, nullptr, d_tick_tp, tick, n_ticks);
        else launch_qlane_f64<double, SWR_MODE_TICKMAX>(c, v, p, n_out, nullptr, d_tick_tp, tick, n_ticks);
    } else {
        if (v.fmt == JT_FMT_S16) launch_generic<int16_t, double, SWR_MODE_TICKMAX>(c, v, p, n_out, (double *)nullptr, d_tick_tp, tick, n_ticks);
        else if (v.fmt == JT_FMT_FLT) launch_generic<float, double, SWR_MODE_TICKMAX>(c, v, p, n_out, (double *)nullptr, d_tick_tp, tick, n_ticks);
        else launch_generic<double, double, SWR_MODE_TICKMAX>(c, v, p, n_out, (double *)nullptr, d_tick_tp, tick, n_ticks);
    }
}
