/*
 * ORACLE (test infrastructure only) -- restatement of FFmpeg libswresample's native
 * resampler as the reference uses it implicitly:
 *   - ebur128 true-peak oversampler  (filters.go:626 "peak=sample+true"; swr -> 192 kHz, DBL)
 *   - aformat=sample_rates=44100 output stage (filters.go:706-710)
 *   - loudnorm's forced 192 kHz input in dynamic mode (normalise.go:257-264)
 * Follows libswresample/resample.c (build_filter, resample_init, invert_initial_buffer,
 * resample_flush, swri_resample) and resample_template.c (resample_common) with swr's
 * default options.  PINNED against the real libswresample 6.1.100 (FFmpeg 8.0.1) in
 * this image: tests/test_oracle_swr.py, tests/golden/swr_*.npz.
 */
#include "orc.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* modified Bessel I0 by power series (upstream uses a boost-derived rational fit;
 * both are accurate to ~1 ulp of double) */
static double bessel_i0(double x)
{
    double y = x * x / 4.0, t = 1.0, sum = 1.0;
    for (int k = 1; k < 200; k++) {
        t *= y / ((double)k * (double)k);
        sum += t;
        if (t < sum * 1e-18) break;
    }
    return sum;
}

static int64_t gcd64(int64_t a, int64_t b) { while (b) { int64_t t = a % b; a = b; b = t; } return a; }

typedef struct {
    int phase_count, filter_length, filter_alloc;
    int64_t inc_num, inc_den; /* index advance per output sample = inc_num / inc_den, in 1/phase_count input samples
                              * (swr's dst_incr / src_incr, reduced): output m reads from index0 + floor(m inc_num / inc_den) */
    double *bank;          /* phase_count x filter_alloc */
    int in_rate, out_rate;
} swr_plan;

static int plan_init2(swr_plan *p, int in_rate, int out_rate, int build_bank)
{
    const int filter_size = 32, phase_shift = 10;
    const double cutoff = 0.97, kaiser_beta = 9.0;
    double factor = fmin((double)out_rate * cutoff / in_rate, 1.0);
    int phase_count = 1 << phase_shift;
    int64_t g = gcd64(out_rate, in_rate);
    int64_t pc_exact = out_rate / g;
    /* exact_rational (default on): the reduced phase count when it fits, else the full 1 << phase_shift phases
     * with a fractional index advance (index += dst_incr_div; frac += dst_incr_mod; carry at src_incr) and -- swr's
     * linear_interp option defaults to on, and swri_resample() picks resample_linear whenever frac or dst_incr_mod is
     * non-zero -- a linear interpolation between the two neighbouring phases by frac / src_incr
     * (resample_template.c resample_linear; checked live against the real library, tests/test_oracle_swr.py) */
    if (pc_exact <= phase_count) phase_count = (int)pc_exact;
    p->phase_count = phase_count;
    p->filter_length = (int)ceil(filter_size / factor);
    if (p->filter_length < 1) p->filter_length = 1;
    p->filter_alloc = (p->filter_length + 7) & ~7;
    p->in_rate = in_rate; p->out_rate = out_rate;
    /* dst_incr / src_incr = in_rate * phase_count / out_rate (av_reduce; the power-of-two scaling that follows in
     * swri_resample_init changes neither quotient nor carries) */
    {
        int64_t num = (int64_t)in_rate * phase_count, den = out_rate, gg = gcd64(num, den);
        p->inc_num = num / gg; p->inc_den = den / gg;
    }
    p->bank = NULL;
    if (!build_bank) return 0;
    p->bank = (double *)calloc((size_t)(phase_count + 1) * p->filter_alloc, sizeof(double));
    if (!p->bank) return -2;

    /* build_filter(), SWR_FILTER_TYPE_KAISER, scale = 1 */
    const int tap_count = p->filter_length, alloc = p->filter_alloc;
    const int ph_nb = phase_count % 2 ? phase_count : phase_count / 2 + 1;
    const int center = (tap_count - 1) / 2;
    double *tab = (double *)malloc(sizeof(double) * (tap_count + 1));
    double *sin_lut = (double *)calloc(ph_nb, sizeof(double));
    double norm = 0.0;
    if (factor > 1.0) factor = 1.0;
    if (factor == 1.0)
        for (int ph = 0; ph < ph_nb; ph++)
            sin_lut[ph] = sin(M_PI * ph / phase_count) * (center & 1 ? 1 : -1);
    for (int ph = 0; ph < ph_nb; ph++) {
        double s = sin_lut[ph];
        for (int i = 0; i < tap_count; i++) {
            double x = M_PI * ((double)(i - center) - (double)ph / phase_count) * factor;
            double y, w;
            if (x == 0) y = 1.0;
            else if (factor == 1.0) y = s / x;
            else y = sin(x) / x;
            w = 2.0 * x / (factor * tap_count * M_PI);
            y *= bessel_i0(kaiser_beta * sqrt(fmax(1 - w * w, 0)));
            tab[i] = y;
            s = -s;
            if (!ph) norm += y;
        }
        for (int i = 0; i < tap_count; i++)
            p->bank[ph * alloc + i] = tab[i] * 1.0 / norm;
        if (phase_count % 2 == 0)
            for (int i = 0; i < tap_count; i++)
                p->bank[(phase_count - ph) * alloc + tap_count - 1 - i] = p->bank[ph * alloc + i];
    }
    free(tab); free(sin_lut);
    /* row phase_count, the upper neighbour of the last phase = phase 0 one sample later (swri_resample_init:
     * memcpy of row 0 shifted by one tap, its first tap taken from row 0's last allocated slot) */
    for (int i = 0; i + 1 < alloc; i++) p->bank[phase_count * alloc + i + 1] = p->bank[i];
    p->bank[phase_count * alloc] = p->bank[alloc - 1];
    return 0;
}

static int plan_init(swr_plan *p, int in_rate, int out_rate) { return plan_init2(p, in_rate, out_rate, 1); }

static void plan_free(swr_plan *p) { free(p->bank); p->bank = NULL; }

/* number of outputs whose taps lie fully inside [.., n_avail) : swri_resample()'s
 * end_index/delta_n bound with index0 = -phase_count*((L-1)/2) */
static int64_t outputs_available(const swr_plan *p, int64_t n_avail)
{
    const int64_t pc = p->phase_count, L = p->filter_length, c = (L - 1) / 2;
    int64_t end_index = (1 + n_avail - L) * pc;         /* in phase units relative to sample 0 */
    int64_t index0 = -pc * c;
    int64_t span = end_index - index0;
    if (span <= 0) return 0;
    /* outputs m with floor(m inc_num / inc_den) < span */
    return (span * p->inc_den + p->inc_num - 1) / p->inc_num;
}

/* swr holds back output until filter_length+1 input samples have arrived
 * (invert_initial_buffer needs them to build the mirrored history) */
static int64_t out_count_noflush(const swr_plan *p, int64_t n_in)
{
    if (n_in < p->filter_length + 1) return 0;
    return outputs_available(p, n_in);
}

static int64_t reflection_len(const swr_plan *p, int64_t n_in, int64_t m_done)
{
    /* in_buffer_count at flush time = samples from the next output's first tap to the end */
    const int64_t pc = p->phase_count, L = p->filter_length, c = (L - 1) / 2;
    int64_t pos = -pc * c + (m_done * p->inc_num) / p->inc_den;
    int64_t sidx = pos >= 0 ? pos / pc : -((-pos + pc - 1) / pc);
    int64_t cnt = n_in - sidx;
    if (cnt > L) cnt = L;
    if (cnt < 0) cnt = 0;
    return (cnt + 1) / 2;
}

int64_t orc_swr_out_count(int64_t n_in, int in_rate, int out_rate)
{
    swr_plan p;
    if (in_rate == out_rate) return n_in;
    if (plan_init2(&p, in_rate, out_rate, 0)) return -1;
    int64_t r = out_count_noflush(&p, n_in);
    plan_free(&p);
    return r;
}

int64_t orc_swr_out_count_flush(int64_t n_in, int in_rate, int out_rate)
{
    swr_plan p;
    if (in_rate == out_rate) return n_in;
    if (plan_init2(&p, in_rate, out_rate, 0)) return -1;
    int64_t m0 = out_count_noflush(&p, n_in);
    int64_t r = reflection_len(&p, n_in, m0);
    int64_t tot = outputs_available(&p, n_in + r);
    plan_free(&p);
    return tot;
}

int orc_swr_filter_bank(int in_rate, int out_rate, int *filter_length, double *bank, int cap)
{
    swr_plan p;
    if (plan_init(&p, in_rate, out_rate)) return -1;
    *filter_length = p.filter_length;
    for (int ph = 0; ph < p.phase_count; ph++)
        for (int i = 0; i < p.filter_length; i++)
            if (ph * p.filter_length + i < cap)
                bank[ph * p.filter_length + i] = p.bank[ph * p.filter_alloc + i];
    int pc = p.phase_count;
    plan_free(&p);
    return pc;
}

/* sample accessor with swr's start mirror x[-k] = x[k] and end reflection
 * x[n+j] = x[n-1-j] (resample_flush) */
#define GEN_RESAMPLE(NAME, T, ACC)                                                           \
int64_t NAME(const T *in, int64_t n, int in_rate, int out_rate, int flush, T *out, int64_t cap) \
{                                                                                            \
    if (in_rate == out_rate) {                                                               \
        int64_t m = n < cap ? n : cap;                                                       \
        if (out) memcpy(out, in, (size_t)m * sizeof(T));                                     \
        return m;                                                                            \
    }                                                                                        \
    swr_plan p;                                                                              \
    if (plan_init(&p, in_rate, out_rate)) return -1;                                         \
    const int64_t pc = p.phase_count, L = p.filter_length, c = (L - 1) / 2;                  \
    int64_t m0 = out_count_noflush(&p, n);                                                   \
    int64_t total = m0, refl = 0;                                                            \
    if (flush) {                                                                             \
        refl = reflection_len(&p, n, m0);                                                    \
        total = outputs_available(&p, n + refl);                                             \
    }                                                                                        \
    if (total > cap) total = cap;                                                            \
    /* the bank in the working precision */                                                  \
    T *bank = (T *)malloc(sizeof(T) * (size_t)(pc + 1) * L);                                 \
    for (int64_t ph = 0; ph <= pc; ph++)                                                     \
        for (int64_t i = 0; i < L; i++) bank[ph * L + i] = (T)p.bank[ph * p.filter_alloc + i]; \
    for (int64_t m = 0; m < total; m++) {                                                    \
        int64_t pos = -pc * c + (m * p.inc_num) / p.inc_den;                                 \
        int64_t sidx = pos >= 0 ? pos / pc : -((-pos + pc - 1) / pc);                        \
        int64_t ph = pos - sidx * pc;                                                        \
        const T *f = bank + ph * L;                                                          \
        ACC val = 0, val2 = 0;                                                               \
        int64_t i;                                                                           \
        if (p.inc_den > 1) {                      /* resample_linear */                      \
            const int64_t frac = (m * p.inc_num) % p.inc_den;                                \
            for (i = 0; i < L; i++) {                                                        \
                int64_t a = sidx + i;                                                        \
                if (a < 0) a = -a; else if (a >= n) a = 2 * n - 1 - a;                       \
                val  += in[a] * (ACC)f[i];                                                   \
                val2 += in[a] * (ACC)f[i + L];                                               \
            }                                                                                \
            val += (val2 - val) * (ACC)frac / (ACC)p.inc_den;                                \
            out[m] = (T)val;                                                                 \
            continue;                                                                        \
        }                                                                                    \
        for (i = 0; i + 1 < L; i += 2) {                                                     \
            int64_t a = sidx + i, b = sidx + i + 1;                                          \
            if (a < 0) a = -a; else if (a >= n) a = 2 * n - 1 - a;                           \
            if (b < 0) b = -b; else if (b >= n) b = 2 * n - 1 - b;                           \
            val  += in[a] * (ACC)f[i];                                                       \
            val2 += in[b] * (ACC)f[i + 1];                                                   \
        }                                                                                    \
        if (i < L) {                                                                         \
            int64_t a = sidx + i;                                                            \
            if (a < 0) a = -a; else if (a >= n) a = 2 * n - 1 - a;                           \
            val += in[a] * (ACC)f[i];                                                        \
        }                                                                                    \
        out[m] = (T)(val + val2);                                                            \
    }                                                                                        \
    free(bank);                                                                              \
    plan_free(&p);                                                                           \
    return total;                                                                            \
}

GEN_RESAMPLE(orc_swr_resample_f64, double, double)
GEN_RESAMPLE(orc_swr_resample_f32, float, float)

/* ---------------- audioconvert.c ---------------- */
static inline int16_t clip16(long v) { return v < -32768 ? -32768 : v > 32767 ? 32767 : (int16_t)v; }

void orc_conv_s16_to_f64(const int16_t *in, int64_t n, double *out)
{ for (int64_t i = 0; i < n; i++) out[i] = in[i] * (1.0 / (1 << 15)); }
void orc_conv_s16_to_f32(const int16_t *in, int64_t n, float *out)
{ for (int64_t i = 0; i < n; i++) out[i] = in[i] * (1.0f / (1 << 15)); }
void orc_conv_f64_to_s16(const double *in, int64_t n, int16_t *out)
{ for (int64_t i = 0; i < n; i++) out[i] = clip16(lrint(in[i] * (1 << 15))); }
void orc_conv_f32_to_s16(const float *in, int64_t n, int16_t *out)
{ for (int64_t i = 0; i < n; i++) out[i] = clip16(lrintf(in[i] * (1 << 15))); }

void orc_downmix_stereo_f32(const float *in, int64_t n, float *out)
{
    const float c = (float)M_SQRT1_2;
    for (int64_t i = 0; i < n; i++) out[i] = in[2 * i] * c + in[2 * i + 1] * c;
}
void orc_downmix_stereo_s16(const int16_t *in, int64_t n, int16_t *out)
{
    /* rematrix s16: coefficients 0.5/0.5 in Q15 (16384), rounding add 16384, >> 15 */
    for (int64_t i = 0; i < n; i++)
        out[i] = clip16(((long)in[2 * i] * 16384 + (long)in[2 * i + 1] * 16384 + 16384) >> 15);
}
