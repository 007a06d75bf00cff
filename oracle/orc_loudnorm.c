/*
 * ORACLE (test infrastructure only) -- FFmpeg's loudnorm filter, mono.
 *
 *   orc_loudnorm_meter : the BS.1770 meter behind loudnorm's input_* / output_* JSON fields
 *                        (internal/processor/normalise.go:64-75, 328-343): libavfilter/ebur128.c, FFmpeg's cut-down port of
 *                        libebur128 -- 4th-order direct-form-II K-weighting, 400 ms gating blocks every 100 ms, 3 s
 *                        short-term blocks every 1 s, BOTH kept as 1000-bin HISTOGRAMS of 0.1 LU (the port dropped
 *                        libebur128's block lists), relative gate -10 LU, LRA from the short-term histogram, sample peak.
 *                        FF_EBUR128_DUAL_MONO doubles the energy.
 *   orc_loudnorm       : the filter itself (libavfilter/af_loudnorm.c): linear mode (one gain) or dynamic mode at 192 kHz
 *                        (3 s look-ahead, 21-tap Gaussian over per-100 ms gain targets, the peak limiter with its
 *                        10 ms attack / 100 ms release state machine, first / inner / final frame handling, and the
 *                        short-input fall-back to a single gain).  The reference reaches dynamic mode whenever the
 *                        linear-mode preconditions fail (normalise.go:683-693, 1294-1304).
 *
 * Two behaviours of the upstream code that are easy to miss and are restated on purpose:
 *   - filter_frame() feeds EVERY frame to r128_in first, including the flush frame that flush_frame() rebuilds from the
 *     3 s delay line: in dynamic mode the last 2.9 s of the stream are metered twice (input_i / input_lra /
 *     input_thresh see them again, continuing the K-weighting state);
 *   - the limiter works in a 210 ms ring that the FINAL frame refills from scratch with one constant gain.
 *
 * Parity unpinned (see orc.h): restated from recollection of the upstream source; BS.1770 known answers and a
 * torchaudio cross-check are in tests/test_oracle_graph.py.
 */
#include "orc.h"
#include <math.h>
#include <float.h>
#include <stdlib.h>
#include <string.h>

/* ---------------------------------------------------------------- libavfilter/ebur128.c ---- */
#define HIST_N 1000
static double hist_energy[HIST_N], hist_bound[HIST_N + 1];
static int hist_ready = 0;
static void hist_init(void)
{
    if (hist_ready) return;
    hist_bound[0] = pow(10.0, (-70.0 + 0.691) / 10.0);
    for (int i = 0; i < HIST_N; i++) hist_energy[i] = pow(10.0, ((double)i / 10.0 - 69.95 + 0.691) / 10.0);
    for (int i = 1; i <= HIST_N; i++) hist_bound[i] = pow(10.0, ((double)i / 10.0 - 70.0 + 0.691) / 10.0);
    hist_ready = 1;
}
static size_t hist_index(double energy)
{
    size_t lo = 0, hi = HIST_N, mid;
    do { mid = (lo + hi) / 2; if (energy >= hist_bound[mid]) lo = mid; else hi = mid; } while (hi - lo != 1);
    return lo;
}
static double e2l(double e) { return 10.0 * (log(e) / log(10.0)) - 0.691; }

typedef struct r128 {
    int s100; double a[5], b[5], v[5];
    double *z2;                 /* squared K-weighted samples of the last 3 s (audio_data ring, starts zero-filled) */
    int64_t n, ring, pos;       /* samples added so far; ring length; next write position */
    int64_t next_block;         /* sample count at which the next 400 ms gating block closes */
    int64_t st_counter;         /* short_term_frame_counter */
    unsigned long h400[HIST_N], h3000[HIST_N];
    double peak, wgt;
} r128;

static void r128_init(r128 *s, int rate, int dual_mono)
{
    hist_init();
    memset(s, 0, sizeof(*s));
    double f0 = 1681.974450955533, G = 3.999843853973347, Q = 0.7071752369554196;
    double K = tan(M_PI * f0 / (double)rate), Vh = pow(10.0, G / 20.0), Vb = pow(Vh, 0.4996667741545416);
    double pb[3] = {0.0, 0.0, 0.0}, pa[3] = {1.0, 0.0, 0.0}, rb[3] = {1.0, -2.0, 1.0}, ra[3] = {1.0, 0.0, 0.0};
    double a0 = 1.0 + K / Q + K * K;
    pb[0] = (Vh + Vb * K / Q + K * K) / a0; pb[1] = 2.0 * (K * K - Vh) / a0; pb[2] = (Vh - Vb * K / Q + K * K) / a0;
    pa[1] = 2.0 * (K * K - 1.0) / a0; pa[2] = (1.0 - K / Q + K * K) / a0;
    f0 = 38.13547087602444; Q = 0.5003270373238773; K = tan(M_PI * f0 / (double)rate);
    ra[1] = 2.0 * (K * K - 1.0) / (1.0 + K / Q + K * K); ra[2] = (1.0 - K / Q + K * K) / (1.0 + K / Q + K * K);
    s->b[0] = pb[0] * rb[0]; s->b[1] = pb[0] * rb[1] + pb[1] * rb[0]; s->b[2] = pb[0] * rb[2] + pb[1] * rb[1] + pb[2] * rb[0];
    s->b[3] = pb[1] * rb[2] + pb[2] * rb[1]; s->b[4] = pb[2] * rb[2];
    s->a[0] = pa[0] * ra[0]; s->a[1] = pa[0] * ra[1] + pa[1] * ra[0]; s->a[2] = pa[0] * ra[2] + pa[1] * ra[1] + pa[2] * ra[0];
    s->a[3] = pa[1] * ra[2] + pa[2] * ra[1]; s->a[4] = pa[2] * ra[2];
    s->s100 = (rate + 5) / 10;
    s->next_block = 4 * (int64_t)s->s100;
    s->ring = 30 * (int64_t)s->s100;                        /* 3000 ms window, a multiple of samples_in_100ms */
    s->z2 = (double *)calloc((size_t)s->ring, sizeof(double));
    s->wgt = dual_mono ? 2.0 : 1.0;
}
static void r128_free(r128 *s) { free(s->z2); s->z2 = NULL; }

/* mean K-weighted energy of the last `w` samples (ebur128_calc_gating_block: the ring starts zero-filled) */
static double r128_window(const r128 *s, int64_t w)
{
    double sum = 0.0;
    int64_t i = s->pos - w;
    if (i < 0) i += s->ring;
    for (int64_t k = 0; k < w; k++) { sum += s->z2[i]; if (++i == s->ring) i = 0; }
    return s->wgt * sum / (double)w;
}

static void r128_add(r128 *s, const double *x, int64_t n)
{
    for (int64_t i = 0; i < n; i++) {
        if (fabs(x[i]) > s->peak) s->peak = fabs(x[i]);
        double *v = s->v;
        v[0] = x[i] - s->a[1] * v[1] - s->a[2] * v[2] - s->a[3] * v[3] - s->a[4] * v[4];
        const double o = s->b[0] * v[0] + s->b[1] * v[1] + s->b[2] * v[2] + s->b[3] * v[3] + s->b[4] * v[4];
        v[4] = v[3]; v[3] = v[2]; v[2] = v[1]; v[1] = v[0];
        s->z2[s->pos] = o * o; s->n++;
        if (++s->pos == s->ring) s->pos = 0;
        s->st_counter++;
        if (s->n == s->next_block) {                       /* a gating block closes every 100 ms after the first 400 ms */
            const double e = r128_window(s, 4 * (int64_t)s->s100);
            if (e >= hist_bound[0]) s->h400[hist_index(e)]++;
            /* short_term_frame_counter is bumped by needed_frames at the same moment */
            if (s->st_counter >= 30 * (int64_t)s->s100) {
                const double st = r128_window(s, 30 * (int64_t)s->s100);
                if (st >= hist_bound[0]) s->h3000[hist_index(st)]++;
                s->st_counter = 20 * (int64_t)s->s100;
            }
            s->next_block += s->s100;
        }
    }
}
/* upstream adds needed_frames (400 ms for the first block) to the counter when a block closes and compares with ==:
 * the counter reaches 30 * s100 exactly at 3 s, then every 1 s; counting samples one by one and testing at block
 * boundaries is the same sequence. */

static int r128_relative_threshold_energy(const r128 *s, double *out)
{
    double sum = 0.0; unsigned long cnt = 0;
    for (int j = 0; j < HIST_N; j++) { sum += s->h400[j] * hist_energy[j]; cnt += s->h400[j]; }
    if (!cnt) { *out = 0.0; return 0; }
    *out = sum / (double)cnt * pow(10.0, -10.0 / 10.0);
    return 1;
}
static double r128_relative_threshold(const r128 *s)
{
    double t;
    if (!r128_relative_threshold_energy(s, &t)) return -70.0;
    return e2l(t);
}
static double r128_global(const r128 *s)
{
    double rel, g = 0.0; unsigned long cnt = 0; size_t start;
    if (!r128_relative_threshold_energy(s, &rel)) return -HUGE_VAL;
    if (rel < hist_bound[0]) start = 0;
    else { start = hist_index(rel); if (rel > hist_energy[start]) ++start; }
    for (size_t j = start; j < HIST_N; j++) { g += s->h400[j] * hist_energy[j]; cnt += s->h400[j]; }
    if (!cnt) return -HUGE_VAL;
    return e2l(g / (double)cnt);
}
static double r128_shortterm(const r128 *s)
{
    const double e = r128_window(s, 30 * (int64_t)s->s100);
    return e <= 0.0 ? -HUGE_VAL : e2l(e);
}
static double r128_lra(const r128 *s)
{
    size_t size = 0, index, j; double power = 0.0, integ;
    for (j = 0; j < HIST_N; j++) { size += s->h3000[j]; power += s->h3000[j] * hist_energy[j]; }
    if (!size) return 0.0;
    power /= (double)size;
    integ = pow(10.0, -20.0 / 10.0) * power;
    if (integ < hist_bound[0]) index = 0;
    else { index = hist_index(integ); if (integ > hist_energy[index]) ++index; }
    size = 0;
    for (j = index; j < HIST_N; j++) size += s->h3000[j];
    if (!size) return 0.0;
    const size_t plo = (size_t)((size - 1) * 0.1 + 0.5), phi = (size_t)((size - 1) * 0.95 + 0.5);
    size = 0; j = index;
    while (size <= plo) size += s->h3000[j++];
    const double l_en = hist_energy[j - 1];
    while (size <= phi) size += s->h3000[j++];
    const double h_en = hist_energy[j - 1];
    return e2l(h_en) - e2l(l_en);
}

int orc_loudnorm_meter(const double *x, int64_t n, int rate, int dual_mono, double *out4 /* I, LRA, thresh, sample_peak */)
{
    r128 s; r128_init(&s, rate, dual_mono);
    r128_add(&s, x, n);
    out4[0] = r128_global(&s); out4[1] = r128_lra(&s); out4[2] = r128_relative_threshold(&s); out4[3] = s.peak;
    r128_free(&s);
    return 0;
}

/* ---------------------------------------------------------------- libavfilter/af_loudnorm.c ---- */
enum { FIRST_FRAME, INNER_FRAME, FINAL_FRAME, LINEAR_MODE };
enum { OUT, ATTACK, SUSTAIN, RELEASE };

typedef struct LN {
    double target_i, target_lra, target_tp, measured_i, measured_lra, measured_tp, measured_thresh, offset;
    double *buf; int buf_size, buf_index, prev_buf_index;
    double delta[30], weights[21], prev_delta; int index;
    double gain_reduction[2]; double *limiter_buf; double prev_smp;
    int limiter_buf_index, limiter_buf_size, limiter_state, peak_index, env_index, env_cnt, attack_length, release_length;
    int frame_type, above_threshold, prev_nb_samples;
    r128 in, out;
} LN;

static int frame_size(int rate, int ms) { const int size = (int)round((double)rate * (ms / 1000.0)); return size + (size % 2); }

static void init_gaussian_filter(LN *s)
{
    double total = 0.0; const double sigma = 3.5;
    const int offset = 21 / 2;
    const double c1 = 1.0 / (sigma * sqrt(2.0 * M_PI)), c2 = 2.0 * pow(sigma, 2.0);
    for (int i = 0; i < 21; i++) { const int x = i - offset; s->weights[i] = c1 * exp(-(pow(x, 2.0) / c2)); total += s->weights[i]; }
    const double adjust = 1.0 / total;
    for (int i = 0; i < 21; i++) s->weights[i] *= adjust;
}
static double gaussian_filter(LN *s, int index)
{
    double result = 0.;
    index = index - 10 > 0 ? index - 10 : index + 20;
    for (int i = 0; i < 21; i++) result += s->delta[((index + i) < 30) ? (index + i) : (index + i - 30)] * s->weights[i];
    return result;
}

static void detect_peak(LN *s, int offset, int nb_samples, int *peak_delta, double *peak_value)
{
    int n, i, index; double ceiling; double *buf;
    *peak_delta = -1;
    buf = s->limiter_buf; ceiling = s->target_tp;
    index = s->limiter_buf_index + offset + 1920;
    if (index >= s->limiter_buf_size) index -= s->limiter_buf_size;
    if (s->frame_type == FIRST_FRAME) s->prev_smp = fabs(buf[index - 1]);
    for (n = 0; n < nb_samples; n++) {
        double this, next, max_peak = 0;
        this = fabs(buf[index < s->limiter_buf_size ? index : index - s->limiter_buf_size]);
        next = fabs(buf[(index + 1) < s->limiter_buf_size ? (index + 1) : (index + 1 - s->limiter_buf_size)]);
        if ((s->prev_smp <= this) && (next <= this) && (this > ceiling) && (n > 0)) {
            int detected = 1;
            for (i = 2; i < 12; i++) {
                next = fabs(buf[(index + i) < s->limiter_buf_size ? (index + i) : (index + i - s->limiter_buf_size)]);
                if (next > this) { detected = 0; break; }
            }
            if (detected) {
                max_peak = fabs(buf[index]);
                s->prev_smp = fabs(buf[index < s->limiter_buf_size ? index : index - s->limiter_buf_size]);
                *peak_delta = n; s->peak_index = index; *peak_value = max_peak;
                return;
            }
            /* upstream `continue`s the channel loop here: with one channel that skips the prev_smp update below
             * but still advances to the next sample */
        } else
            s->prev_smp = this;
        index += 1;
        if (index >= s->limiter_buf_size) index -= s->limiter_buf_size;
    }
}

static void true_peak_limiter(LN *s, double *out, int nb_samples)
{
    int n, index, peak_delta, smp_cnt; double ceiling, peak_value = 0; double *buf;
    buf = s->limiter_buf; ceiling = s->target_tp; index = s->limiter_buf_index; smp_cnt = 0;
    if (s->frame_type == FIRST_FRAME) {
        double max = 0.;
        for (n = 0; n < 1920; n++) max = fabs(buf[n]) > max ? fabs(buf[n]) : max;
        if (max > ceiling) {
            s->gain_reduction[1] = ceiling / max; s->limiter_state = SUSTAIN;
            for (n = 0; n < 1920; n++) buf[n] *= s->gain_reduction[1];
        }
    }
    do {
        switch (s->limiter_state) {
        case OUT:
            detect_peak(s, smp_cnt, nb_samples - smp_cnt, &peak_delta, &peak_value);
            if (peak_delta != -1) {
                s->env_cnt = 0;
                smp_cnt += (peak_delta - s->attack_length);
                s->gain_reduction[0] = 1.; s->gain_reduction[1] = ceiling / peak_value;
                s->limiter_state = ATTACK;
                s->env_index = s->peak_index - s->attack_length;
                if (s->env_index < 0) s->env_index += s->limiter_buf_size;
                s->env_index += s->env_cnt;
                if (s->env_index > s->limiter_buf_size) s->env_index -= s->limiter_buf_size;
            } else smp_cnt = nb_samples;
            break;
        case ATTACK:
            for (; s->env_cnt < s->attack_length; s->env_cnt++) {
                const double env = s->gain_reduction[0] - ((double)s->env_cnt / (s->attack_length - 1) * (s->gain_reduction[0] - s->gain_reduction[1]));
                buf[s->env_index] *= env;
                s->env_index += 1;
                if (s->env_index >= s->limiter_buf_size) s->env_index -= s->limiter_buf_size;
                smp_cnt++;
                if (smp_cnt >= nb_samples) { s->env_cnt++; break; }
            }
            if (smp_cnt < nb_samples) { s->env_cnt = 0; s->attack_length = 1920; s->limiter_state = SUSTAIN; }
            break;
        case SUSTAIN:
            detect_peak(s, smp_cnt, nb_samples, &peak_delta, &peak_value);
            if (peak_delta == -1) {
                s->limiter_state = RELEASE;
                s->gain_reduction[0] = s->gain_reduction[1]; s->gain_reduction[1] = 1.; s->env_cnt = 0;
                break;
            } else {
                const double gain_reduction = ceiling / peak_value;
                if (gain_reduction < s->gain_reduction[1]) {
                    s->limiter_state = ATTACK;
                    s->attack_length = peak_delta;
                    if (s->attack_length <= 1) s->attack_length = 2;
                    s->gain_reduction[0] = s->gain_reduction[1]; s->gain_reduction[1] = gain_reduction; s->env_cnt = 0;
                    break;
                }
                for (s->env_cnt = 0; s->env_cnt < peak_delta; s->env_cnt++) {
                    buf[s->env_index] *= s->gain_reduction[1];
                    s->env_index += 1;
                    if (s->env_index >= s->limiter_buf_size) s->env_index -= s->limiter_buf_size;
                    smp_cnt++;
                    if (smp_cnt >= nb_samples) { s->env_cnt++; break; }
                }
            }
            break;
        case RELEASE:
            for (; s->env_cnt < s->release_length; s->env_cnt++) {
                const double env = s->gain_reduction[0] + (((double)s->env_cnt / (s->release_length - 1)) * (s->gain_reduction[1] - s->gain_reduction[0]));
                buf[s->env_index] *= env;
                s->env_index += 1;
                if (s->env_index >= s->limiter_buf_size) s->env_index -= s->limiter_buf_size;
                smp_cnt++;
                if (smp_cnt >= nb_samples) { s->env_cnt++; break; }
            }
            if (smp_cnt < nb_samples) { s->env_cnt = 0; s->limiter_state = OUT; }
            break;
        }
    } while (smp_cnt < nb_samples);
    for (n = 0; n < nb_samples; n++) {
        out[n] = buf[index];
        if (fabs(out[n]) > ceiling) out[n] = ceiling * (out[n] < 0 ? -1 : 1);
        index += 1;
        if (index >= s->limiter_buf_size) index -= s->limiter_buf_size;
    }
}

/* one call of filter_frame(): consumes `nb` samples at src, writes the frame's output at dst, returns its length */
static int ln_filter_frame(LN *s, int rate, const double *src, int nb, double *dst)
{
    double *buf = s->buf, *limiter_buf = s->limiter_buf;
    int n, i, subframe_length, src_index, out_nb = nb;
    double gain, gain_next, env_global, env_shortterm, global, shortterm, relative_threshold;

    r128_add(&s->in, src, nb);
    if (s->frame_type == FIRST_FRAME && nb < frame_size(rate, 3000)) {
        double offset, offset_tp, true_peak;
        global = r128_global(&s->in); true_peak = s->in.peak;
        offset = pow(10., (s->target_i - global) / 20.);
        offset_tp = true_peak * offset;
        s->offset = offset_tp < s->target_tp ? offset : s->target_tp / true_peak;
        s->frame_type = LINEAR_MODE;
    }
    switch (s->frame_type) {
    case FIRST_FRAME:
        for (n = 0; n < nb; n++) { buf[s->buf_index] = src[n]; s->buf_index += 1; }
        shortterm = r128_shortterm(&s->in);
        if (shortterm < s->measured_thresh) { s->above_threshold = 0; env_shortterm = shortterm <= -70. ? 0. : s->target_i - s->measured_i; }
        else { s->above_threshold = 1; env_shortterm = shortterm <= -70. ? 0. : s->target_i - shortterm; }
        for (n = 0; n < 30; n++) s->delta[n] = pow(10., env_shortterm / 20.);
        s->prev_delta = s->delta[s->index];
        s->buf_index = s->limiter_buf_index = 0;
        for (n = 0; n < s->limiter_buf_size; n++) {
            limiter_buf[s->limiter_buf_index] = buf[s->buf_index] * s->delta[s->index] * s->offset;
            s->limiter_buf_index += 1;
            if (s->limiter_buf_index == s->limiter_buf_size) s->limiter_buf_index -= s->limiter_buf_size;
            s->buf_index += 1;
        }
        subframe_length = frame_size(rate, 100);
        true_peak_limiter(s, dst, subframe_length);
        r128_add(&s->out, dst, subframe_length);
        out_nb = subframe_length;
        s->frame_type = INNER_FRAME;
        break;
    case INNER_FRAME:
        gain = gaussian_filter(s, s->index + 10 < 30 ? s->index + 10 : s->index + 10 - 30);
        gain_next = gaussian_filter(s, s->index + 11 < 30 ? s->index + 11 : s->index + 11 - 30);
        for (n = 0; n < nb; n++) {
            buf[s->prev_buf_index] = src[n];
            limiter_buf[s->limiter_buf_index] = buf[s->buf_index] * (gain + (((double)n / nb) * (gain_next - gain))) * s->offset;
            s->limiter_buf_index += 1;
            if (s->limiter_buf_index == s->limiter_buf_size) s->limiter_buf_index -= s->limiter_buf_size;
            s->prev_buf_index += 1;
            if (s->prev_buf_index == s->buf_size) s->prev_buf_index -= s->buf_size;
            s->buf_index += 1;
            if (s->buf_index == s->buf_size) s->buf_index -= s->buf_size;
        }
        subframe_length = frame_size(rate, 100) - nb;
        s->limiter_buf_index = s->limiter_buf_index + subframe_length < s->limiter_buf_size ? s->limiter_buf_index + subframe_length : s->limiter_buf_index + subframe_length - s->limiter_buf_size;
        true_peak_limiter(s, dst, nb);
        r128_add(&s->out, dst, nb);
        (void)r128_lra(&s->in);
        global = r128_global(&s->in); shortterm = r128_shortterm(&s->in); relative_threshold = r128_relative_threshold(&s->in);
        if (s->above_threshold == 0) {
            double shortterm_out;
            if (shortterm > s->measured_thresh) s->prev_delta *= 1.0058;
            shortterm_out = r128_shortterm(&s->out);
            if (shortterm_out >= s->target_i) s->above_threshold = 1;
        }
        if (shortterm < relative_threshold || shortterm <= -70. || s->above_threshold == 0) s->delta[s->index] = s->prev_delta;
        else {
            env_global = fabs(shortterm - global) < (s->target_lra / 2.) ? shortterm - global : (s->target_lra / 2.) * ((shortterm - global) < 0 ? -1 : 1);
            env_shortterm = s->target_i - shortterm;
            s->delta[s->index] = pow(10., (env_global + env_shortterm) / 20.);
        }
        s->prev_delta = s->delta[s->index];
        s->index++;
        if (s->index >= 30) s->index -= 30;
        s->prev_nb_samples = nb;
        break;
    case FINAL_FRAME:
        gain = gaussian_filter(s, s->index + 10 < 30 ? s->index + 10 : s->index + 10 - 30);
        s->limiter_buf_index = 0; src_index = 0;
        for (n = 0; n < s->limiter_buf_size; n++) {
            s->limiter_buf[s->limiter_buf_index] = src[src_index] * gain * s->offset;
            src_index += 1;
            s->limiter_buf_index += 1;
            if (s->limiter_buf_index == s->limiter_buf_size) s->limiter_buf_index -= s->limiter_buf_size;
        }
        subframe_length = frame_size(rate, 100);
        {
            double *d = dst;
            for (i = 0; i < nb / subframe_length; i++) {
                true_peak_limiter(s, d, subframe_length);
                for (n = 0; n < subframe_length; n++) {
                    if (src_index < nb) limiter_buf[s->limiter_buf_index] = src[src_index] * gain * s->offset;
                    else limiter_buf[s->limiter_buf_index] = 0.;
                    if (src_index < nb) src_index += 1;
                    s->limiter_buf_index += 1;
                    if (s->limiter_buf_index == s->limiter_buf_size) s->limiter_buf_index -= s->limiter_buf_size;
                }
                d += subframe_length;
            }
        }
        r128_add(&s->out, dst, nb);
        break;
    case LINEAR_MODE:
        for (n = 0; n < nb; n++) dst[n] = src[n] * s->offset;
        r128_add(&s->out, dst, nb);
        break;
    }
    return out_nb;
}

/* opt[8] = I, TP, LRA, measured_I, measured_TP, measured_LRA, measured_thresh, offset (the filter's options, dB / LU);
 * x is the filter's INPUT LINK: the caller resamples to 192 kHz when orc_loudnorm_is_linear() says dynamic, as
 * query_formats() makes libavfilter do.  frame_in: the size of the frames the link delivers in linear mode (any value:
 * linear mode is frame-size independent).  stats[10] = input_i, input_tp, input_lra, input_thresh, output_i, output_tp,
 * output_lra, output_thresh, normalization_type (0 linear / 1 dynamic, as printed), target_offset.
 * Returns the number of output samples (== n), or -1. */
int orc_loudnorm_is_linear(const double *opt, int linear)
{
    const double I = opt[0], TP = opt[1], LRA = opt[2], mI = opt[3], mTP = opt[4], mLRA = opt[5], mTh = opt[6];
    if (!linear) return 0;
    const double offset = I - mI, offset_tp = mTP + offset;
    if (mTP != 99 && mTh != -70 && mLRA != 0 && mI != 0) if (offset_tp <= TP && mLRA <= LRA) return 1;
    return 0;
}

int64_t orc_loudnorm(const double *x, int64_t n, int rate, const double *opt, int linear, int dual_mono,
                     double *y, double *stats)
{
    LN *s = (LN *)calloc(1, sizeof(LN));
    int64_t produced = 0;
    s->target_i = opt[0]; s->target_tp = opt[1]; s->target_lra = opt[2]; s->measured_i = opt[3]; s->measured_tp = opt[4];
    s->measured_lra = opt[5]; s->measured_thresh = opt[6]; s->offset = opt[7];
    /* init() */
    s->frame_type = FIRST_FRAME;
    if (orc_loudnorm_is_linear(opt, linear)) { s->frame_type = LINEAR_MODE; s->offset = s->target_i - s->measured_i; }
    /* config_input() */
    r128_init(&s->in, rate, dual_mono); r128_init(&s->out, rate, dual_mono);
    s->buf_size = frame_size(rate, 3000); s->buf = (double *)calloc((size_t)s->buf_size, sizeof(double));
    s->limiter_buf_size = frame_size(rate, 210); s->limiter_buf = (double *)calloc((size_t)s->buf_size, sizeof(double));
    init_gaussian_filter(s);
    s->index = 1; s->limiter_state = OUT;
    s->offset = pow(10., s->offset / 20.); s->target_tp = pow(10., s->target_tp / 20.);
    s->attack_length = frame_size(rate, 10); s->release_length = frame_size(rate, 100);
    /* activate(): FIRST_FRAME wants 3 s, later frames 100 ms (ff_inlink_consume_samples hands over a short frame at
     * EOF); LINEAR_MODE takes frames as they come */
    int64_t pos = 0;
    if (s->frame_type == LINEAR_MODE) {
        while (pos < n) { const int nb = (int)(n - pos < 4096 ? n - pos : 4096); produced += ln_filter_frame(s, rate, x + pos, nb, y + produced); pos += nb; }
    } else {
        while (pos < n) {
            const int want = s->frame_type == FIRST_FRAME ? frame_size(rate, 3000) : frame_size(rate, 100);
            const int nb = (int)(n - pos < want ? n - pos : want);
            produced += ln_filter_frame(s, rate, x + pos, nb, y + produced); pos += nb;
        }
        /* flush_frame() */
        if (s->frame_type == INNER_FRAME) {
            int nb_samples = s->buf_size - s->prev_nb_samples;
            nb_samples -= (frame_size(rate, 100) - s->prev_nb_samples);
            double *frame = (double *)calloc((size_t)nb_samples, sizeof(double));
            int offset = s->limiter_buf_size - s->prev_nb_samples;
            offset -= (frame_size(rate, 100) - s->prev_nb_samples);
            s->buf_index = s->buf_index - offset < 0 ? s->buf_index - offset + s->buf_size : s->buf_index - offset;
            for (int k = 0; k < nb_samples; k++) {
                frame[k] = s->buf[s->buf_index];
                s->buf_index += 1;
                if (s->buf_index >= s->buf_size) s->buf_index -= s->buf_size;
            }
            s->frame_type = FINAL_FRAME;
            produced += ln_filter_frame(s, rate, frame, nb_samples, y + produced);
            free(frame);
        }
    }
    /* uninit(): the JSON */
    stats[0] = r128_global(&s->in); stats[1] = 20. * log10(s->in.peak); stats[2] = r128_lra(&s->in); stats[3] = r128_relative_threshold(&s->in);
    stats[4] = r128_global(&s->out); stats[5] = 20. * log10(s->out.peak); stats[6] = r128_lra(&s->out); stats[7] = r128_relative_threshold(&s->out);
    stats[8] = s->frame_type == LINEAR_MODE ? 0.0 : 1.0;
    stats[9] = s->target_i - stats[4];
    r128_free(&s->in); r128_free(&s->out); free(s->buf); free(s->limiter_buf); free(s);
    return produced;
}
