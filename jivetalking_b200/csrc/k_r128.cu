// BS.1770 meters.
//   jt_ebur128        : FFmpeg's ebur128 filter (libavfilter/f_ebur128.c) as instantiated by
//                       "ebur128=metadata=1:peak=sample+true:dualmono=true" (reference:
//                       internal/processor/filters.go:626, analyser_output.go:18)
//   jt_loudnorm_meter : the libebur128 port loudnorm uses (libavfilter/ebur128.c) for its
//                       input_i / input_tp / input_lra / input_thresh (reference:
//                       internal/processor/normalise.go:257-264,328-343)
// Device: K-weighting + 100 ms mean-square / sample-peak partials, one sequential lane per
// group of ticks with a 200 ms warm-up (the RLB high-pass forgets its state to < 1e-19 in
// that time); true peak via k_swr.cu.  Host: windowing, gating, histograms, LRA on the
// ~10 values per second the device produced (same split as the reference: FFmpeg emits
// values, Go accumulates them).
#include "jt_internal.h"
#include "jt_device.cuh"
#include "jt_lanes.cuh"
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <mutex>

struct KWeight { double pb0, pb1, pb2, pa1, pa2, rb0, rb1, rb2, ra1, ra2; double b[5], a[5]; };

static KWeight kweight_design(int rate)
{
    KWeight k;
    double f0 = 1681.974450955533, G = 3.999843853973347, Q = 0.7071752369554196;
    double K = tan(M_PI * f0 / (double)rate);
    double Vh = pow(10.0, G / 20.0), Vb = pow(Vh, 0.4996667741545416);
    double a0 = 1.0 + K / Q + K * K;
    k.pb0 = (Vh + Vb * K / Q + K * K) / a0;
    k.pb1 = 2.0 * (K * K - Vh) / a0;
    k.pb2 = (Vh - Vb * K / Q + K * K) / a0;
    k.pa1 = 2.0 * (K * K - 1.0) / a0;
    k.pa2 = (1.0 - K / Q + K * K) / a0;
    f0 = 38.13547087602444; Q = 0.5003270373238773;
    K = tan(M_PI * f0 / (double)rate);
    k.rb0 = 1.0; k.rb1 = -2.0; k.rb2 = 1.0;
    k.ra1 = 2.0 * (K * K - 1.0) / (1.0 + K / Q + K * K);
    k.ra2 = (1.0 - K / Q + K * K) / (1.0 + K / Q + K * K);
    const double pb[3] = {k.pb0, k.pb1, k.pb2}, pa[3] = {1.0, k.pa1, k.pa2};
    const double rb[3] = {1.0, -2.0, 1.0}, ra[3] = {1.0, k.ra1, k.ra2};
    k.b[0] = pb[0] * rb[0];
    k.b[1] = pb[0] * rb[1] + pb[1] * rb[0];
    k.b[2] = pb[0] * rb[2] + pb[1] * rb[1] + pb[2] * rb[0];
    k.b[3] = pb[1] * rb[2] + pb[2] * rb[1];
    k.b[4] = pb[2] * rb[2];
    k.a[0] = pa[0] * ra[0];
    k.a[1] = pa[0] * ra[1] + pa[1] * ra[0];
    k.a[2] = pa[0] * ra[2] + pa[1] * ra[1] + pa[2] * ra[0];
    k.a[3] = pa[1] * ra[2] + pa[2] * ra[1];
    k.a[4] = pa[2] * ra[2];
    return k;
}

// One lane = G consecutive UNITS plus a warm-up of `warm` samples (1.5 ticks: 36 time constants of the RLB high-pass, whose
// state is forgotten to 2e-16 by then).  A unit is a whole 100 ms tick (S = 1: long streams, where there are more ticks than
// lanes to fill the machine) or one of S equal parts of a tick (short streams -- a 60 s region, a 10 min file -- where whole-
// tick lanes would leave most SMs idle and the launch would cost a fixed (warm + G ticks) walk whatever the length).
// STRUCT 0: cascade of two direct form I biquads (f_ebur128.c FILTER macro); STRUCT 1: 4th-order direct form II (ebur128.c).
__device__ __forceinline__ int64_t r128_unit_start(int64_t u, int S, int L, int tick) { const int64_t k = u / S; return k * tick + (u - k * S) * (int64_t)L; }
__device__ __forceinline__ int64_t r128_unit_end(int64_t u, int S, int L, int tick) { const int64_t k = u / S; return k * tick + min((u - k * S + 1) * (int64_t)L, (int64_t)tick); }

template <class TIN, int STRUCT>
__global__ void __launch_bounds__(64)
k_r128_ticks(const TIN *__restrict__ x, int64_t n, int tick, int64_t n_units_total, int G, int warm, int S, int L,
             const __grid_constant__ KWeight kw, double *__restrict__ unit_pow, double *__restrict__ unit_peak)
{
    constexpr int R = 512 / (int)sizeof(TIN);                    // 512-byte rows, double-buffered
    extern __shared__ __align__(16) unsigned char smem[];
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = t * G < n_units_total;
    const int64_t k0 = min(t * G, n_units_total);
    const int64_t k1 = min(k0 + (int64_t)G, n_units_total);
    const int64_t warm_end = min(r128_unit_start(k0, S, L, tick), n);
    const int64_t begin = max((int64_t)0, warm_end - warm);
    const int64_t end = live ? min(r128_unit_end(k1 - 1, S, L, tick), n) : begin;
    LaneStage<TIN, R, 2> in;
    in.init(smem + (size_t)(threadIdx.x >> 5) * LaneStage<TIN, R, 2>::WARP_BYTES, x + begin, max(end - begin, (int64_t)0));
    double x1 = 0, x2 = 0, y1 = 0, y2 = 0, z1 = 0, z2 = 0;      // STRUCT 0
    double v1 = 0, v2 = 0, v3 = 0, v4 = 0;                        // STRUCT 1
    // One warp per scheduler: the walk is bound by the dependent-issue latency of f64 (8 cycles), so the
    // recurrences are written with the newest state entering LAST -- the feed-forward part and the older
    // feedback terms are off the carried chain, which is then one DFMA per biquad per sample.
    auto step = [&](double x0) -> double {
        if (STRUCT == 0) {
            const double t = fma(x2, kw.pb2, fma(x1, kw.pb1, x0 * kw.pb0));
            const double y0 = fma(-y1, kw.pa1, fma(-y2, kw.pa2, t));
            x2 = x1; x1 = x0;
            const double s = fma(y2, kw.rb2, fma(y1, kw.rb1, y0 * kw.rb0));
            const double z0 = fma(-z1, kw.ra1, fma(-z2, kw.ra2, s));
            y2 = y1; y1 = y0; z2 = z1; z1 = z0;
            return z0;
        } else {
            const double v0 = fma(-kw.a[1], v1, fma(-kw.a[2], v2, fma(-kw.a[3], v3, fma(-kw.a[4], v4, x0))));
            const double o = fma(kw.b[0], v0, fma(kw.b[1], v1, fma(kw.b[2], v2, fma(kw.b[3], v3, kw.b[4] * v4))));
            v4 = v3; v3 = v2; v2 = v1; v1 = v0;
            return o;
        }
    };
    // sample peak as an integer maximum of the magnitude bits (non-negative doubles order like their bit
    // patterns; f64 min / max issue at a third of the add rate, profiles/ubench_r1.txt)
    unsigned long long pkb = 0ull;
    auto peak_in = [&](double v) { const unsigned long long b = (unsigned long long)__double_as_longlong(v) & 0x7fffffffffffffffull; pkb = b > pkb ? b : pkb; };
    int64_t k = k0, unit_end = min(r128_unit_end(k0, S, L, tick), n);
    double acc = 0.0;
    in.prime();
    for (int tile = 0; tile < in.ntiles; tile++) {
        in.prefetch();
        const TIN *row = in.wait(tile);
        const int nv = in.valid(tile);
        int64_t i = begin + (int64_t)tile * R;
        int q = 0;
        // branch-free runs: [warm-up run] then runs that end at a unit boundary
        while (q < nv) {
            if (i < warm_end) {
                const int run = (int)min((int64_t)(nv - q), warm_end - i);
                int r = 0;
                for (; r + 8 <= run; r += 8) {          // loads first, then the recurrence
                    double v[8];
#pragma unroll
                    for (int j = 0; j < 8; j++) v[j] = jt_as_f64(row[q + r + j]);
#pragma unroll
                    for (int j = 0; j < 8; j++) (void)step(v[j]);
                }
                for (; r < run; r++) (void)step(jt_as_f64(row[q + r]));
                q += run; i += run;
            } else {
                const int run = (int)min((int64_t)(nv - q), unit_end - i);
                int r = 0;
                for (; r + 8 <= run; r += 8) {
                    double v[8];
#pragma unroll
                    for (int j = 0; j < 8; j++) v[j] = jt_as_f64(row[q + r + j]);
#pragma unroll
                    for (int j = 0; j < 8; j++) { const double z = step(v[j]); acc = fma(z, z, acc); peak_in(v[j]); }
                }
                for (; r < run; r++) {
                    const double x0 = jt_as_f64(row[q + r]);
                    const double z = step(x0);
                    acc = fma(z, z, acc);
                    peak_in(x0);
                }
                q += run; i += run;
                if (i == unit_end) {
                    unit_pow[k] = acc; unit_peak[k] = __longlong_as_double((long long)pkb);
                    acc = 0.0; pkb = 0ull; k++;
                    // (units past the stream's end stay at the zeros the buffers were cleared to)
                    unit_end = k < k1 ? min(r128_unit_end(k, S, L, tick), n) : n + 1;
                }
            }
        }
        in.release();
    }
}

// per-tick values from the S parts of each tick, summed in part order (deterministic)
__global__ void k_r128_fold(const double *__restrict__ unit_pow, const double *__restrict__ unit_peak, int64_t n_ticks, int S,
                            double *__restrict__ tick_pow, double *__restrict__ tick_peak)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_ticks) return;
    double p = 0.0, m = 0.0;
    for (int j = 0; j < S; j++) { p += unit_pow[k * S + j]; m = fmax(m, unit_peak[k * S + j]); }
    tick_pow[k] = p; tick_peak[k] = m;
}

// ---------------------------------------------------------------------------------------------------------------------------
// The same per-tick values WITHOUT lanes and warm-up walks: K-weighting is linear, so the filter state at any sample is
//   (end state of a zero-state run over the block before it)  +  Phi * (state at that block's start)
// with Phi the 4 x 4 state transition over one block.  Three passes over blocks of B samples (B divides the tick, ~96):
//   k_kw_blocks<.., 0>  thread per block: zero-state run -> the block's forced end state e_b            (9 DP per sample)
//   k_kw_scan           thread per tick : s_{b+1} = Phi s_b + e_b, started two ticks (48 RLB time constants) early from
//                       rest -> the true state at the start of each of the tick's blocks                (20 DP per block)
//   k_kw_blocks<.., 1>  thread per block: the filter from that state -> energy and sample peak           (10 DP per sample)
// and k_kw_fold sums a tick's block energies in block order.  A block's samples are walked by ONE thread with exactly the
// recurrence of the sequential filter, so given its start state the result is the sequential one; the start state itself
// differs from the sequential state by rounding only (1e-16 relative).  Every block is independent: 1.8 million threads for an
// hour of audio, 5000 for a 10 s region -- the lane kernel above spent the same (warm-up + two ticks) walk whatever the length
// and, at 6 warps per SM, 112 cycles per sample.  The CTA stages its 128 blocks in shared memory with coalesced row loads
// (pitch B + 1: the per-thread walks are conflict free).
struct KwPhi { double m[16]; };

template <int STRUCT>
__host__ __device__ __forceinline__ double kw_step(const KWeight &kw, double x0, double &x1, double &x2, double &s0, double &s1, double &s2, double &s3)
{
    if (STRUCT == 0) {                          // state: y1, y2, z1, z2 (+ the input history x1, x2)
        const double t = fma(x2, kw.pb2, fma(x1, kw.pb1, x0 * kw.pb0));
        const double y0 = fma(-s0, kw.pa1, fma(-s1, kw.pa2, t));
        x2 = x1; x1 = x0;
        const double u = fma(s1, kw.rb2, fma(s0, kw.rb1, y0 * kw.rb0));
        const double z0 = fma(-s2, kw.ra1, fma(-s3, kw.ra2, u));
        s1 = s0; s0 = y0; s3 = s2; s2 = z0;
        return z0;
    } else {                                    // state: v1 .. v4
        const double v0 = fma(-kw.a[1], s0, fma(-kw.a[2], s1, fma(-kw.a[3], s2, fma(-kw.a[4], s3, x0))));
        const double o = fma(kw.b[0], v0, fma(kw.b[1], s0, fma(kw.b[2], s1, fma(kw.b[3], s2, kw.b[4] * s3))));
        s3 = s2; s2 = s1; s1 = s0; s0 = v0;
        return o;
    }
}

// THREADS * B contiguous samples -> tile rows of pitch B + 1, as vectors of VW elements (a vector never straddles two rows: VW | B)
template <class TIN, int VW, int THREADS>
__device__ __forceinline__ void kw_stage_vec(const TIN *__restrict__ src, TIN *__restrict__ tile, int B, int pitch, int tid)
{
    struct alignas(sizeof(TIN) * VW) Vec { TIN v[VW]; };
    const Vec *xv = (const Vec *)src;
    const int total = THREADS * B, step = THREADS * VW, dq = step / B, dr = step % B;
    int idx = tid * VW, r = idx / B, col = idx % B;
    while (idx < total) {
        Vec v[4]; int off[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            off[j] = -1;
            if (idx < total) {
                v[j] = xv[idx / VW]; off[j] = r * pitch + col;
                idx += step; r += dq; col += dr; if (col >= B) { col -= B; r++; }
            }
        }
#pragma unroll
        for (int j = 0; j < 4; j++) if (off[j] >= 0) {
#pragma unroll
            for (int e = 0; e < VW; e++) tile[off[j] + e] = v[j].v[e];
        }
    }
}

template <class TIN, int STRUCT, int PHASE, int THREADS>
__global__ void __launch_bounds__(THREADS)
k_kw_blocks(const TIN *__restrict__ x, int64_t n, int B, int64_t n_blocks, const __grid_constant__ KWeight kw,
            const double *__restrict__ s_in, double *__restrict__ s_out, double *__restrict__ e_out, double *__restrict__ pk_out)
{
    extern __shared__ __align__(16) unsigned char smem[];
    TIN *tile = (TIN *)smem;
    const int pitch = B + 1, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t b0 = (int64_t)blockIdx.x * THREADS, g0 = b0 * B;
    {   // the CTA's THREADS * B samples are contiguous: flat coalesced loads, several in flight per thread, (row, col) tracked
        // incrementally (a load-store-load chain per row left the kernel waiting on DRAM latency: ncu r2q).  CTAs that lie
        // inside the stream and start on a 16-byte boundary (all but the last, for the usual block sizes) load 16-byte vectors.
        constexpr int VMAX = 16 / (int)sizeof(TIN);
        const int total = THREADS * B;
        const bool inside = g0 + total <= n;
        const uintptr_t addr = (uintptr_t)(x + g0);
        if (inside && B % VMAX == 0 && (addr & 15) == 0) kw_stage_vec<TIN, VMAX, THREADS>(x + g0, tile, B, pitch, tid);
        else if (VMAX >= 4 && inside && B % (VMAX / 2) == 0 && (addr & 7) == 0) kw_stage_vec<TIN, (VMAX >= 4 ? VMAX / 2 : 1), THREADS>(x + g0, tile, B, pitch, tid);
        else {
            const int dq = THREADS / B, dr = THREADS % B;
            int idx = tid, r = tid / B, col = tid % B;
            while (idx < total) {
                TIN v[8]; int off[8];
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    off[j] = -1;
                    if (idx < total) {
                        const int64_t g = g0 + idx;
                        v[j] = g < n ? x[g] : (TIN)0; off[j] = r * pitch + col;
                        idx += THREADS; r += dq; col += dr; if (col >= B) { col -= B; r++; }
                    }
                }
#pragma unroll
                for (int j = 0; j < 8; j++) if (off[j] >= 0) tile[off[j]] = v[j];
            }
        }
    }
    (void)lane; (void)warp;
    __syncthreads();
    const int64_t b = b0 + tid;
    if (b >= n_blocks) return;
    const TIN *row = tile + tid * pitch;
    double x1 = 0, x2 = 0;
    if (STRUCT == 0) {
        const int64_t t0 = b * B;
        if (tid > 0) { x1 = jt_as_f64(row[-pitch + B - 1]); x2 = B >= 2 ? jt_as_f64(row[-pitch + B - 2]) : 0.0; }
        else { if (t0 >= 1 && t0 - 1 < n) x1 = jt_as_f64(x[t0 - 1]); if (t0 >= 2 && t0 - 2 < n) x2 = jt_as_f64(x[t0 - 2]); }
    }
    double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    if (PHASE == 1) { s0 = s_in[4 * b]; s1 = s_in[4 * b + 1]; s2 = s_in[4 * b + 2]; s3 = s_in[4 * b + 3]; }
    double acc = 0.0; unsigned long long pkb = 0ull;
    int k = 0;
    for (; k + 8 <= B; k += 8) {
        double v[8];
#pragma unroll
        for (int j = 0; j < 8; j++) v[j] = jt_as_f64(row[k + j]);
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const double z = kw_step<STRUCT>(kw, v[j], x1, x2, s0, s1, s2, s3);
            if (PHASE == 1) {
                acc = fma(z, z, acc);
                const unsigned long long bits = (unsigned long long)__double_as_longlong(v[j]) & 0x7fffffffffffffffull; pkb = bits > pkb ? bits : pkb;
            }
        }
    }
    for (; k < B; k++) {
        const double v = jt_as_f64(row[k]);
        const double z = kw_step<STRUCT>(kw, v, x1, x2, s0, s1, s2, s3);
        if (PHASE == 1) { acc = fma(z, z, acc); const unsigned long long bits = (unsigned long long)__double_as_longlong(v) & 0x7fffffffffffffffull; pkb = bits > pkb ? bits : pkb; }
    }
    if (PHASE == 0) { s_out[4 * b] = s0; s_out[4 * b + 1] = s1; s_out[4 * b + 2] = s2; s_out[4 * b + 3] = s3; }
    else { e_out[b] = acc; pk_out[b] = __longlong_as_double((long long)pkb); }
}

// thread per tick: the state at the start of each of its P blocks (in place: e[] holds forced end states on entry of a block's
// turn and is only read; the states go to s[])
__global__ void __launch_bounds__(128)
k_kw_scan(const double *__restrict__ e, int64_t n_blocks, int P, int64_t n_ticks, int warm_blocks, const __grid_constant__ KwPhi phi, double *__restrict__ s)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_ticks) return;
    const int64_t b0 = k * P, b1 = min(b0 + (int64_t)P, n_blocks);
    double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    for (int64_t b = max((int64_t)0, b0 - warm_blocks); b < b1; b++) {
        if (b >= b0) { s[4 * b] = s0; s[4 * b + 1] = s1; s[4 * b + 2] = s2; s[4 * b + 3] = s3; }
        const double n0 = fma(phi.m[0], s0, fma(phi.m[1], s1, fma(phi.m[2], s2, fma(phi.m[3], s3, e[4 * b]))));
        const double n1 = fma(phi.m[4], s0, fma(phi.m[5], s1, fma(phi.m[6], s2, fma(phi.m[7], s3, e[4 * b + 1]))));
        const double n2 = fma(phi.m[8], s0, fma(phi.m[9], s1, fma(phi.m[10], s2, fma(phi.m[11], s3, e[4 * b + 2]))));
        const double n3 = fma(phi.m[12], s0, fma(phi.m[13], s1, fma(phi.m[14], s2, fma(phi.m[15], s3, e[4 * b + 3]))));
        s0 = n0; s1 = n1; s2 = n2; s3 = n3;
    }
}

__global__ void k_kw_fold(const double *__restrict__ e_blk, const double *__restrict__ pk_blk, int64_t n_blocks, int P, int64_t n_ticks,
                          double *__restrict__ tick_pow, double *__restrict__ tick_peak)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_ticks) return;
    double p = 0.0, m = 0.0;
    for (int64_t b = k * P; b < min((k + 1) * (int64_t)P, n_blocks); b++) { p += e_blk[b]; m = fmax(m, pk_blk[b]); }
    tick_pow[k] = p; tick_peak[k] = m;
}

static int kw_block_size(int tick)
{
    int best = 0;
    for (int d = 64; d <= 160; d++) if (tick % d == 0 && (!best || std::abs(d - 96) < std::abs(best - 96))) best = d;
    return best;
}

template <int STRUCT>
static KwPhi kw_transition(const KWeight &kw, int B)
{
    KwPhi phi;
    for (int col = 0; col < 4; col++) {
        double st[4] = {0, 0, 0, 0}; st[col] = 1.0;
        double x1 = 0, x2 = 0;
        for (int i = 0; i < B; i++) (void)kw_step<STRUCT>(kw, 0.0, x1, x2, st[0], st[1], st[2], st[3]);
        for (int r = 0; r < 4; r++) phi.m[4 * r + col] = st[r];
    }
    return phi;
}

template <class TIN, int STRUCT>
static void run_ticks_blocks_t(jt_ctx *c, const Sig &in, int tick, int64_t n_ticks, int B, const KWeight &kw, double *d_pow, double *d_peak)
{
    constexpr int THREADS = sizeof(TIN) == 8 ? 64 : 128;
    const int P = tick / B;
    const int64_t n_blocks = n_ticks * P;
    double *d_e = jt_dalloc<double>(c, (size_t)n_blocks * 4), *d_s = jt_dalloc<double>(c, (size_t)n_blocks * 4);
    double *d_eb = jt_dalloc<double>(c, (size_t)n_blocks), *d_pk = jt_dalloc<double>(c, (size_t)n_blocks);
    const size_t smem = (size_t)THREADS * (B + 1) * sizeof(TIN);
    const int grid = (int)((n_blocks + THREADS - 1) / THREADS);
    const KwPhi phi = kw_transition<STRUCT>(kw, B);
    jt_smem_optin((const void *)k_kw_blocks<TIN, STRUCT, 0, THREADS>, smem);
    jt_smem_optin((const void *)k_kw_blocks<TIN, STRUCT, 1, THREADS>, smem);
    k_kw_blocks<TIN, STRUCT, 0, THREADS><<<grid, THREADS, smem, c->stream>>>((const TIN *)in.d, in.n, B, n_blocks, kw, nullptr, d_e, nullptr, nullptr);
    k_kw_scan<<<(int)((n_ticks + 127) / 128), 128, 0, c->stream>>>(d_e, n_blocks, P, n_ticks, 2 * P, phi, d_s);
    k_kw_blocks<TIN, STRUCT, 1, THREADS><<<grid, THREADS, smem, c->stream>>>((const TIN *)in.d, in.n, B, n_blocks, kw, d_s, nullptr, d_eb, d_pk);
    k_kw_fold<<<(int)((n_ticks + 255) / 256), 256, 0, c->stream>>>(d_eb, d_pk, n_blocks, P, n_ticks, d_pow, d_peak);
}

template <int STRUCT>
static void run_ticks(jt_ctx *c, const Sig &in0, int tick, int64_t n_ticks_total, const KWeight &kw,
                      double *d_pow, double *d_peak)
{
    if (n_ticks_total <= 0) return;
    const Sig in = in0.fmt == JT_FMT_S32 ? jt_convert(c, in0, JT_FMT_DBL) : in0;      // the filter's link is dbl: s32 widens exactly
    {
        static const char *lanes_only = getenv("JT_R128_LANES");
        const int B = kw_block_size(tick);
        // (the 4th-order direct form II of loudnorm's meter, STRUCT 1, keeps the lanes: its internal state is ~tau^2 = 6e5 times
        //  the signal and the output a second difference of it, so a state re-derived to 1e-16 relative moves the output by
        //  1e-8 -- the sequential rounding trajectory af_loudnorm's dynamic gains are tested against to 2e-9 is only reproduced by
        //  walking it; the cascade of direct form I biquads of f_ebur128.c carries states of the signal's own size)
        if (B && STRUCT == 0 && !(lanes_only && *lanes_only == '1')) {
            JtLaunch Lb(c, in.rate >= 176400 ? "r128_kweight_ticks:192k" : "r128_kweight_ticks", 4);
            if (in.fmt == JT_FMT_S16) run_ticks_blocks_t<int16_t, STRUCT>(c, in, tick, n_ticks_total, B, kw, d_pow, d_peak);
            else if (in.fmt == JT_FMT_FLT) run_ticks_blocks_t<float, STRUCT>(c, in, tick, n_ticks_total, B, kw, d_pow, d_peak);
            else run_ticks_blocks_t<double, STRUCT>(c, in, tick, n_ticks_total, B, kw, d_pow, d_peak);
            return;
        }
    }
    // 34 KB of staging per warp: 6 warps per SM
    const int64_t slots = (int64_t)c->num_sms * 6 * 32;
    const int warm = tick + tick / 2;
    int S = 1, G;
    if (n_ticks_total * 4 < slots) {
        // a short stream: cut every tick into S parts so the lanes fill the machine (parts of at least 256 samples)
        S = (int)std::min<int64_t>(std::min<int64_t>(16, slots / (2 * n_ticks_total)), std::max(1, tick / 256));
        if (S < 1) S = 1;
        G = 1;
    } else G = (int)std::max<int64_t>(2, (n_ticks_total + slots - 1) / slots);      // ticks per lane; grows instead of a second wave of CTAs
    const int L = (tick + S - 1) / S;
    const int64_t n_units = n_ticks_total * S;
    const int64_t lanes = (n_units + G - 1) / G;
    const int grid = (int)((lanes + 63) / 64);
    double *u_pow = d_pow, *u_peak = d_peak;
    if (S > 1) { u_pow = jt_dalloc<double>(c, (size_t)n_units); u_peak = jt_dalloc<double>(c, (size_t)n_units); }
    JT_CUDA(cudaMemsetAsync(u_pow, 0, sizeof(double) * (size_t)n_units, c->stream));
    JT_CUDA(cudaMemsetAsync(u_peak, 0, sizeof(double) * (size_t)n_units, c->stream));
    JtLaunch Lc(c, in.rate >= 176400 ? "r128_kweight_ticks:192k" : "r128_kweight_ticks", S > 1 ? 2 : 1);
#define R128_LAUNCH(T) do { \
        const size_t smem = 2 * LaneStage<T, 512 / (int)sizeof(T), 2>::WARP_BYTES; \
        jt_smem_optin((const void *)k_r128_ticks<T, STRUCT>, (size_t)(smem)); \
        k_r128_ticks<T, STRUCT><<<grid, 64, smem, c->stream>>>((const T *)in.d, in.n, tick, n_units, G, warm, S, L, kw, u_pow, u_peak); } while (0)
    if (in.fmt == JT_FMT_S16) R128_LAUNCH(int16_t);
    else if (in.fmt == JT_FMT_FLT) R128_LAUNCH(float);
    else R128_LAUNCH(double);
#undef R128_LAUNCH
    if (S > 1) k_r128_fold<<<(int)((n_ticks_total + 255) / 256), 256, 0, c->stream>>>(u_pow, u_peak, n_ticks_total, S, d_pow, d_peak);
}

// ---------------------------------------------------------------------------------------
// ebur128 filter
// ---------------------------------------------------------------------------------------
#define ABS_THRES (-70)
#define HIST_RES 100
#define HIST_SIZE ((10 - ABS_THRES) * HIST_RES + 1)
static inline double loudness_of(double e) { return -0.691 + 10 * log10(e); }
static inline int hist_pos(double l) { int p = (int)((l - ABS_THRES) * HIST_RES); return p < 0 ? 0 : p > HIST_SIZE - 1 ? HIST_SIZE - 1 : p; }

void jt_ebur128_launch(jt_ctx *c, const Sig &in, bool dualmono, bool true_peak, R128Pending &pd)
{
    if (in.rate % 10) JT_THROW(JT_ERR_UNSUPPORTED, "ebur128 at %d Hz (rate must be a multiple of 10)", in.rate);
    pd = R128Pending();
    pd.tick = in.rate / 10; pd.nt = in.n / pd.tick; pd.dualmono = dualmono; pd.true_peak = true_peak;
    const int tick = pd.tick; const int64_t nt = pd.nt;
    if (nt <= 0) return;
    double *d_pow = jt_dalloc<double>(c, nt), *d_peak = jt_dalloc<double>(c, nt), *d_tp = jt_dalloc<double>(c, nt);
    KWeight kw = kweight_design(in.rate);
    run_ticks<0>(c, in, tick, nt, kw, d_pow, d_peak);
    if (true_peak) {
        if (in.rate == 192000) JT_CUDA(cudaMemcpyAsync(d_tp, d_peak, sizeof(double) * nt, cudaMemcpyDeviceToDevice, c->stream));
        else { SwrPlan p = jt_swr_plan(in.rate, 192000); jt_swr_tick_absmax(c, in.fmt == JT_FMT_S32 ? jt_convert(c, in, JT_FMT_DBL) : in, p, tick, nt, d_tp); }
    }
    pd.hp = jt_pinned<double>(c, nt); pd.hk = jt_pinned<double>(c, nt); pd.ht = jt_pinned<double>(c, nt);
    jt_copy_small(c, pd.hp, d_pow, sizeof(double) * nt);
    jt_copy_small(c, pd.hk, d_peak, sizeof(double) * nt);
    if (true_peak) jt_copy_small(c, pd.ht, d_tp, sizeof(double) * nt);
    pd.ev = jt_record_event(c);
}

void jt_ebur128_finish(jt_ctx *c, R128Pending &pd, R128Result &out)
{
    out = R128Result();
    out.n_ticks = pd.nt;
    if (pd.nt <= 0) return;
    JT_CUDA(cudaEventSynchronize(pd.ev));
    jt_ebur128_host_finalize(c, pd.hp, pd.hk, pd.true_peak ? pd.ht : nullptr, pd.nt, pd.tick, pd.dualmono, out);
}

// f_ebur128.c filter_frame() tail over the per-tick values (K-weighted energy, sample peak, true peak of each
// 100 ms tick): windowing, gating histograms, LRA.  Pure host code: also the merge step when the ticks of one
// stream were produced by several GPUs.
void jt_ebur128_host_finalize(jt_ctx *c, const double *hp, const double *hk, const double *ht_or_null, int64_t nt, int tick,
                              bool dualmono, R128Result &out)
{
    out = R128Result();
    out.n_ticks = nt;
    if (nt <= 0) return;
    std::vector<double> zero_tp;
    const double *ht = ht_or_null;
    if (!ht) { zero_tp.assign(nt, 0.0); ht = zero_tp.data(); }

    // host: f_ebur128.c filter_frame() tail, once per 100 ms
    JtHost hfin(c, "r128_finalize");
    out.M.resize(nt); out.S.resize(nt); out.sp_cum.resize(nt); out.tp_cum.resize(nt);
    std::vector<uint32_t> h400(HIST_SIZE, 0), h3000(HIST_SIZE, 0);
    double kept400 = 0, kept3000 = 0; uint64_t nk400 = 0, nk3000 = 0;
    double rel400 = 0, rel3000 = 0; bool any400 = false, any3000 = false;
    const double pan_law = -3.01029995663978;
    double sp = 0, tp = 0, w400 = 0, w3000 = 0;
    for (int64_t k = 0; k < nt; k++) {
        sp = std::max(sp, hk[k]); tp = std::max(tp, ht[k]);
        out.sp_cum[k] = sp; out.tp_cum[k] = tp;
        w400 += hp[k]; if (k >= 4) w400 -= hp[k - 4];
        w3000 += hp[k]; if (k >= 30) w3000 -= hp[k - 30];
        double p400 = 1e-12, p3000 = 1e-12;
        if (k >= 3) { double s = 0; for (int j = 3; j >= 0; j--) s += hp[k - j]; p400 = (p400 + s) / (4.0 * tick); }
        if (k >= 29) { double s = 0; for (int j = 29; j >= 0; j--) s += hp[k - j]; p3000 = (p3000 + s) / (30.0 * tick); }
        double l400 = loudness_of(p400), l3000 = loudness_of(p3000);
        if (l400 >= ABS_THRES) {
            h400[hist_pos(l400)]++; kept400 += p400; nk400++;
            double rt = kept400 / nk400; if (!rt) rt = 1e-12;
            rel400 = loudness_of(rt) - 10; any400 = true;
        }
        if (l3000 >= ABS_THRES) {
            h3000[hist_pos(l3000)]++; kept3000 += p3000; nk3000++;
            double rt = kept3000 / nk3000; if (!rt) rt = 1e-12;
            rel3000 = loudness_of(rt) - 20; any3000 = true;
        }
        if (dualmono) { l400 -= pan_law; l3000 -= pan_law; }
        out.M[k] = l400; out.S[k] = l3000;
    }
    (void)w400; (void)w3000;
    if (any400) {
        double isum = 0; uint64_t nb = 0;
        for (int i = hist_pos(rel400); i < HIST_SIZE; i++) {
            const double loud = i / (double)HIST_RES + ABS_THRES;
            nb += h400[i]; isum += h400[i] * exp2(3.32192809488736234787 * ((loud + 0.691) / 10.));
        }
        if (nb) { out.I = loudness_of(isum / nb); if (dualmono) out.I -= pan_law; }
    }
    if (any3000) {
        uint64_t nb_powers = 0; const int pos = hist_pos(rel3000);
        for (int i = pos; i < HIST_SIZE; i++) nb_powers += h3000[i];
        if (nb_powers) {
            uint64_t nn = 0, nb_pow = (uint64_t)(10 * nb_powers * 0.01 + 0.5);
            for (int i = pos; i < HIST_SIZE; i++) { nn += h3000[i]; if (nn >= nb_pow) { out.LRA_low = i / (double)HIST_RES + ABS_THRES; break; } }
            nn = nb_powers; nb_pow = (uint64_t)(95 * nb_powers * 0.01 + 0.5);
            for (int i = HIST_SIZE - 1; i >= 0; i--) {
                uint64_t cc = h3000[i]; nn -= std::min(nn, cc);
                if (nn < nb_pow) { out.LRA_high = i / (double)HIST_RES + ABS_THRES; break; }
            }
            out.LRA = out.LRA_high - out.LRA_low;
        }
    }
}

// ---------------------------------------------------------------------------------------
// loudnorm's meter: libavfilter/ebur128.c (FFmpeg's cut-down libebur128 port), histogram gating
// ---------------------------------------------------------------------------------------
void jt_loudnorm_meter_launch(jt_ctx *c, const Sig &in, bool dual_mono, LoudnormPending &pd)
{
    pd = LoudnormPending();
    pd.s100 = (in.rate + 5) / 10; pd.dual_mono = dual_mono;
    pd.nfull = in.n / pd.s100;
    pd.nt = (in.n + pd.s100 - 1) / pd.s100;                // last partial tick only feeds the sample peak
    if (pd.nt <= 0) return;
    double *d_pow = jt_dalloc<double>(c, pd.nt), *d_peak = jt_dalloc<double>(c, pd.nt);
    KWeight kw = kweight_design(in.rate);
    run_ticks<1>(c, in, pd.s100, pd.nt, kw, d_pow, d_peak);
    pd.hp = jt_pinned<double>(c, pd.nt); pd.hk = jt_pinned<double>(c, pd.nt);
    jt_copy_small(c, pd.hp, d_pow, sizeof(double) * pd.nt);
    jt_copy_small(c, pd.hk, d_peak, sizeof(double) * pd.nt);
    pd.ev = jt_record_event(c);
}

void jt_loudnorm_meter_finish(jt_ctx *c, LoudnormPending &pd, LoudnormMeter &out)
{
    out.I = -HUGE_VAL; out.LRA = 0; out.thresh = -70.0; out.sample_peak = 0;
    if (pd.nt <= 0) return;
    JT_CUDA(cudaEventSynchronize(pd.ev));
    jt_loudnorm_meter_host_finalize(pd.hp, pd.hk, pd.nt, pd.nfull, pd.s100, pd.dual_mono, out);
}

// libavfilter/ebur128.c keeps its 400 ms gating blocks and its 3 s short-term blocks in two 1000-bin histograms of
// 0.1 LU (the port dropped libebur128's block lists): integrated loudness, relative threshold and LRA are functions
// of the histograms, i.e. quantised to bin centres.  LnMeterState replays that state machine over per-100 ms
// K-weighted energies; af_loudnorm's dynamic mode queries it after every 100 ms frame (k_loudnorm.cu).
static double g_hist_energy[1000], g_hist_bound[1001];
static void hist_tables()
{
    static std::once_flag once;
    std::call_once(once, []() {
        g_hist_bound[0] = pow(10.0, (-70.0 + 0.691) / 10.0);
        for (int i = 0; i < 1000; i++) g_hist_energy[i] = pow(10.0, ((double)i / 10.0 - 69.95 + 0.691) / 10.0);
        for (int i = 1; i <= 1000; i++) g_hist_bound[i] = pow(10.0, ((double)i / 10.0 - 70.0 + 0.691) / 10.0);
    });
}
static size_t hist_index(double energy)
{
    size_t lo = 0, hi = 1000, mid;
    do { mid = (lo + hi) / 2; if (energy >= g_hist_bound[mid]) lo = mid; else hi = mid; } while (hi - lo != 1);
    return lo;
}
static inline double energy_to_loudness(double e) { return 10.0 * (log(e) / log(10.0)) - 0.691; }

LnMeterState::LnMeterState(int s100_, bool dual_mono) : s100(s100_), wgt(dual_mono ? 2.0 : 1.0), h400(1000, 0), h3000(1000, 0) { hist_tables(); }

void LnMeterState::add_tick(const double *hp, int64_t k)
{
    // tick k has just completed: k + 1 ticks of audio have been seen
    if (k >= 3) {
        const double e = wgt * (hp[k - 3] + hp[k - 2] + hp[k - 1] + hp[k]) / (4.0 * s100);
        if (e >= g_hist_bound[0]) { const size_t j = hist_index(e); h400[j]++; sum400 += g_hist_energy[j]; cnt400++; }
    }
    if (k >= 29 && (k - 29) % 10 == 0) {               // short_term_frame_counter: first at 3 s, then every 1 s
        const double e = shortterm_energy(hp, k);
        if (e >= g_hist_bound[0]) h3000[hist_index(e)]++;
    }
}
double LnMeterState::shortterm_energy(const double *hp, int64_t k) const
{
    double s = 0; for (int64_t j = std::max<int64_t>(0, k - 29); j <= k; j++) s += hp[j];
    return wgt * s / (30.0 * s100);
}
double LnMeterState::shortterm(const double *hp, int64_t k) const
{
    const double e = shortterm_energy(hp, k);
    return e <= 0.0 ? -HUGE_VAL : energy_to_loudness(e);
}
// sums run in bin order like ebur128_calc_relative_threshold, so the values are those of the upstream loops bit for bit
static double rel_threshold_energy(const std::vector<uint32_t> &h, uint64_t cnt)
{
    double sum = 0; for (size_t j = 0; j < 1000; j++) sum += h[j] * g_hist_energy[j];
    return sum / (double)cnt * pow(10.0, -10.0 / 10.0);
}
double LnMeterState::relative_threshold() const
{
    if (!cnt400) return -70.0;
    return energy_to_loudness(rel_threshold_energy(h400, cnt400));
}
double LnMeterState::global() const
{
    if (!cnt400) return -HUGE_VAL;
    const double rel = rel_threshold_energy(h400, cnt400);
    size_t start;
    if (rel < g_hist_bound[0]) start = 0; else { start = hist_index(rel); if (rel > g_hist_energy[start]) ++start; }
    double g = 0; uint64_t cnt = 0;
    for (size_t j = start; j < 1000; j++) { g += h400[j] * g_hist_energy[j]; cnt += h400[j]; }
    if (!cnt) return -HUGE_VAL;
    return energy_to_loudness(g / (double)cnt);
}
double LnMeterState::lra() const
{
    uint64_t size = 0; double power = 0;
    for (size_t j = 0; j < 1000; j++) { size += h3000[j]; power += h3000[j] * g_hist_energy[j]; }
    if (!size) return 0.0;
    power /= (double)size;
    const double integ = pow(10.0, -20.0 / 10.0) * power;
    size_t index;
    if (integ < g_hist_bound[0]) index = 0; else { index = hist_index(integ); if (integ > g_hist_energy[index]) ++index; }
    size = 0;
    for (size_t j = index; j < 1000; j++) size += h3000[j];
    if (!size) return 0.0;
    const uint64_t plo = (uint64_t)((size - 1) * 0.1 + 0.5), phi = (uint64_t)((size - 1) * 0.95 + 0.5);
    size = 0; size_t j = index;
    while (size <= plo) size += h3000[j++];
    const double l_en = g_hist_energy[j - 1];
    while (size <= phi) size += h3000[j++];
    const double h_en = g_hist_energy[j - 1];
    return energy_to_loudness(h_en) - energy_to_loudness(l_en);
}

void jt_loudnorm_meter_host_finalize(const double *hp, const double *hk, int64_t nt, int64_t nfull, int s100, bool dual_mono, LoudnormMeter &out)
{
    out.I = -HUGE_VAL; out.LRA = 0; out.thresh = -70.0; out.sample_peak = 0;
    if (nt <= 0) return;
    for (int64_t k = 0; k < nt; k++) out.sample_peak = std::max(out.sample_peak, hk[k]);
    LnMeterState st(s100, dual_mono);
    for (int64_t k = 0; k < nfull; k++) st.add_tick(hp, k);
    out.I = st.global(); out.LRA = st.lra(); out.thresh = st.relative_threshold();
}

void jt_ebur128(jt_ctx *c, const Sig &in, bool dualmono, bool true_peak, R128Result &out)
{
    R128Pending pd; jt_ebur128_launch(c, in, dualmono, true_peak, pd); jt_ebur128_finish(c, pd, out);
}
void jt_loudnorm_meter(jt_ctx *c, const Sig &in, bool dual_mono, LoudnormMeter &out)
{
    LoudnormPending pd; jt_loudnorm_meter_launch(c, in, dual_mono, pd); jt_loudnorm_meter_finish(c, pd, out);
}
