/*
 * ORACLE (test infrastructure only) -- restatement of FFmpeg's afftdn (FFT spectral-subtraction
 * denoiser) as the reference instantiates it: "afftdn=nr=12:nt=w|custom:bn=<15>:tn=0|1[:nf=<dB>]"
 * (internal/processor/filters.go:830-861; adaptive.go:133-170).  Follows libavfilter/af_afftdn.c
 * (config_input, set_parameters, set_band_parameters, process_frame, filter_channel,
 * output_frame) from recollection of upstream -- parity unpinned (see orc.h).  The 15 profile
 * band centres are the list the reference verified against the 8.1 source
 * (internal/processor/analyser_noise_bands.go:15-17).  fltp path: float32 FFT data scaled by
 * 2^23, float64 statistics.  Output = input delayed by window - advance samples (no latency
 * compensation, no flush).
 */
#include "orc.h"
#include "orc_fft.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define C_ (M_LN10 * 0.1)
#define SOLVE_SIZE 5
#define NB_PROFILE_BANDS 15

static const int band_centre[NB_PROFILE_BANDS] = {80, 125, 195, 290, 440, 660, 1000, 1500, 2250, 3350, 5000, 7500, 11200, 16000, 24000};

typedef struct {
    double sample_rate; int sample_advance, window_length, fft_length, fft_length2, bin_count, number_of_bands;
    int *bin2band; double *window, *band_alpha, *band_beta;
    double floor, window_weight;
    double matrix_a[SOLVE_SIZE * SOLVE_SIZE], vector_b[SOLVE_SIZE], matrix_b[SOLVE_SIZE * NB_PROFILE_BANDS];
    /* channel */
    double band_noise[NB_PROFILE_BANDS];
    double *amt, *band_amt, *band_excit, *gain, *prior, *prior_band_excit, *clean_data, *noisy_data, *out_samples,
           *spread_function, *abs_var, *rel_var, *min_abs_var;
    double noise_reduction, last_noise_reduction, noise_floor, last_noise_floor, max_gain, max_var, gain_scale;
    double ratio, floor_offset_opt; int track_noise;
} Ctx;

static double freq2bark(double x) { double d = x / 7500.0; return 13.0 * atan(7.6E-4 * x) + 3.5 * atan(d * d); }

static void factor(double *array, int size)
{
    for (int i = 0; i < size - 1; i++)
        for (int j = i + 1; j < size; j++) {
            double d = array[j + i * size] / array[i + i * size];
            array[j + i * size] = d;
            for (int k = i + 1; k < size; k++) array[j + k * size] -= d * array[i + k * size];
        }
}
static void solve(double *matrix, double *vector, int size)
{
    for (int i = 0; i < size - 1; i++)
        for (int j = i + 1; j < size; j++) { double d = matrix[j + i * size]; vector[j] -= d * vector[i]; }
    vector[size - 1] /= matrix[size * size - 1];
    for (int i = size - 2; i >= 0; i--) {
        double d = vector[i];
        for (int j = i + 1; j < size; j++) d -= matrix[i + j * size] * vector[j];
        vector[i] = d / matrix[i + i * size];
    }
}
static double process_get_band_noise(Ctx *s, int band)
{
    double product, sum, f; int i = 0;
    if (band < NB_PROFILE_BANDS) return s->band_noise[band];
    for (int j = 0; j < SOLVE_SIZE; j++) {
        sum = 0.0;
        for (int k = 0; k < NB_PROFILE_BANDS; k++) sum += s->matrix_b[i++] * s->band_noise[k];
        s->vector_b[j] = sum;
    }
    solve(s->matrix_a, s->vector_b, SOLVE_SIZE);
    f = (0.5 * s->sample_rate) / band_centre[NB_PROFILE_BANDS - 1];
    f = 15.0 + log(f / 1.5) / log(1.5);
    sum = 0.0; product = 1.0;
    for (int j = 0; j < SOLVE_SIZE; j++) { sum += product * s->vector_b[j]; product *= f; }
    return sum;
}
static void set_band_parameters(Ctx *s)
{
    double band_noise, d2 = 1, d3, d4, d5 = 0.0; int i = 0, j = 0, k = 0;
    band_noise = process_get_band_noise(s, 0);
    for (int m = j; m < s->bin_count; m++) {
        if (m == j) {
            i = j; d5 = band_noise;
            if (k >= NB_PROFILE_BANDS) j = s->bin_count;
            else j = (int)(s->fft_length * band_centre[k] / s->sample_rate);
            d2 = j - i;
            band_noise = process_get_band_noise(s, k);
            k++;
        }
        d3 = (j - m) / d2; d4 = (m - i) / d2;
        s->rel_var[m] = exp((d5 * d3 + band_noise * d4) * C_);
    }
}
static void set_parameters(Ctx *s, int update_var)
{
    if (s->last_noise_floor != s->noise_floor) s->last_noise_floor = s->noise_floor;
    s->max_var = s->floor * exp((100.0 + s->last_noise_floor) * C_);
    if (s->noise_reduction != s->last_noise_reduction) {
        s->last_noise_reduction = s->noise_reduction;
        s->max_gain = exp(s->last_noise_reduction * (0.5 * C_));
    }
    s->gain_scale = 1.0 / (s->max_gain * s->max_gain);
    if (update_var) {
        set_band_parameters(s);
        for (int i = 0; i < s->bin_count; i++) {
            s->abs_var[i] = fmax(s->max_var * s->rel_var[i], 1.0);
            s->min_abs_var[i] = s->gain_scale * s->abs_var[i];
        }
    }
}
static double limit_gain(double a, double b)
{
    if (a > 1.0) return (b * a - 1.0) / (b + a - 2.0);
    if (a < 1.0) return (b * a - 2.0 * a + 1.0) / (b - a);
    return 1.0;
}

static void process_frame(Ctx *s, orc_cf *fft_data, int first_frame)
{
    const double ratio = first_frame ? 1.0 : s->ratio, rratio = 1. - ratio;
    for (int i = 0; i < s->bin_count; i++) {
        double mag = hypot(fft_data[i].re, fft_data[i].im), power, mav, nmav, g, sg;
        s->noisy_data[i] = mag;
        power = mag * mag;
        mav = power / s->abs_var[i];
        nmav = ratio * s->prior[i] + rratio * fmax(mav - 1.0, 0.0);
        g = nmav / (1.0 + nmav);
        sg = g * g;
        s->prior[i] = mav * sg;
        s->clean_data[i] = power * sg;
        s->gain[i] = g;
    }
    if (s->track_noise) {
        double num = 0., den = 0.; int size = 0;
        for (int n = 0; n < s->bin_count; n++) {
            const double v = s->noisy_data[n];
            if (v > s->floor) { num += log(v); den += v; size++; }
        }
        if (size < 1) size = 1;
        num /= size; den /= size; num = exp(num);
        if (num / den > 0.8) {
            double off = 0.0;
            for (int n = 0; n < s->bin_count; n++) off = fmax(off, fabs(s->noisy_data[n] - den));
            off = s->floor_offset_opt * (off / den);
            double nf = 10.0 * log10(den) - 100.0 + off;
            nf = nf < -90. ? -90. : nf > -20. ? -20. : nf;
            s->noise_floor = 0.1 * nf + s->noise_floor * 0.9;
            set_parameters(s, 1);
        }
    }
    for (int i = 0; i < s->number_of_bands; i++) { s->band_excit[i] = 0.0; s->band_amt[i] = 0.0; }
    for (int i = 0; i < s->bin_count; i++) s->band_excit[s->bin2band[i]] += s->clean_data[i];
    for (int i = 0; i < s->number_of_bands; i++) {
        s->band_excit[i] = fmax(s->band_excit[i], s->band_alpha[i] * s->band_excit[i] + s->band_beta[i] * s->prior_band_excit[i]);
        s->prior_band_excit[i] = s->band_excit[i];
    }
    for (int j = 0, i = 0; j < s->number_of_bands; j++)
        for (int k = 0; k < s->number_of_bands; k++) s->band_amt[j] += s->spread_function[i++] * s->band_excit[k];
    for (int i = 0; i < s->bin_count; i++) s->amt[i] = s->band_amt[s->bin2band[i]];
    for (int i = 0; i < s->bin_count; i++) {
        if (s->amt[i] > s->abs_var[i]) s->gain[i] = 1.0;
        else if (s->amt[i] > s->min_abs_var[i]) s->gain[i] = limit_gain(s->gain[i], sqrt(s->abs_var[i] / s->amt[i]));
        else s->gain[i] = limit_gain(s->gain[i], s->max_gain);
    }
    for (int i = 0; i < s->bin_count; i++) {
        const float ng = (float)s->gain[i];
        fft_data[i].re *= ng; fft_data[i].im *= ng;
    }
}

int orc_afftdn(const float *x, float *y, int64_t n, int rate, double nr, double nf, int noise_type,
               const double *bn15 /* or NULL */, int track_noise, double ratio, double floor_offset, double band_multiplier)
{
    Ctx S; Ctx *s = &S; memset(s, 0, sizeof(*s));
    s->sample_rate = (float)rate;
    s->sample_advance = (int)(s->sample_rate / 80);
    s->window_length = 3 * s->sample_advance;
    s->fft_length = 1; while (s->fft_length <= s->window_length) s->fft_length <<= 1;   /* 1 << (32 - clz(window_length)) */
    s->fft_length2 = s->fft_length / 2;
    s->bin_count = s->fft_length2 + 1;
    s->ratio = ratio; s->floor_offset_opt = floor_offset; s->track_noise = track_noise;
    for (int j = 0; j < SOLVE_SIZE; j++)
        for (int k = 0; k < SOLVE_SIZE; k++) {
            s->matrix_a[j + k * SOLVE_SIZE] = 0.0;
            for (int m = 0; m < NB_PROFILE_BANDS; m++) s->matrix_a[j + k * SOLVE_SIZE] += pow(m, j + k);
        }
    factor(s->matrix_a, SOLVE_SIZE);
    { int i = 0; for (int j = 0; j < SOLVE_SIZE; j++) for (int k = 0; k < NB_PROFILE_BANDS; k++) s->matrix_b[i++] = pow(k, j); }
    s->window = calloc(s->window_length, sizeof(double));
    s->bin2band = calloc(s->bin_count, sizeof(int));
    const double sdiv = band_multiplier;
    for (int i = 0; i < s->bin_count; i++) s->bin2band[i] = (int)lrint(sdiv * freq2bark((0.5 * i * s->sample_rate) / s->fft_length2));
    s->number_of_bands = s->bin2band[s->bin_count - 1] + 1;
    const int nb = s->number_of_bands, bc = s->bin_count;
    s->band_alpha = calloc(nb, sizeof(double)); s->band_beta = calloc(nb, sizeof(double));
    if (noise_type == 3 && bn15) for (int i = 0; i < NB_PROFILE_BANDS; i++) { double v = (float)bn15[i]; s->band_noise[i] = v < -24. ? -24. : v > 24. ? 24. : v; }
    else if (noise_type == 1 || noise_type == 2) {
        for (int i = 0; i < NB_PROFILE_BANDS; i++) {
            double a = noise_type == 1 ? 50.0 : 1.0, b = noise_type == 1 ? 500.5 : 500.0, c = noise_type == 1 ? 2125.0 : 1.0E10, d1, d2, d3;
            d1 = a / band_centre[i]; d1 = 10.0 * log(1.0 + d1 * d1) / M_LN10;
            d2 = b / band_centre[i]; d2 = 10.0 * log(1.0 + d2 * d2) / M_LN10;
            d3 = band_centre[i] / c; d3 = 10.0 * log(1.0 + d3 * d3) / M_LN10;
            s->band_noise[i] = -d1 + d2 - d3;
        }
    }
    { double mean = 0; for (int i = 0; i < NB_PROFILE_BANDS; i++) mean += s->band_noise[i]; mean /= NB_PROFILE_BANDS;
      for (int i = 0; i < NB_PROFILE_BANDS; i++) s->band_noise[i] -= mean; }
    s->amt = calloc(bc, 8); s->band_amt = calloc(nb, 8); s->band_excit = calloc(nb, 8); s->gain = calloc(bc, 8);
    s->prior = calloc(bc, 8); s->prior_band_excit = calloc(nb, 8); s->clean_data = calloc(bc, 8); s->noisy_data = calloc(bc, 8);
    s->out_samples = calloc(s->fft_length * 2, 8); s->spread_function = calloc((size_t)nb * nb, 8);
    s->abs_var = calloc(bc, 8); s->rel_var = calloc(bc, 8); s->min_abs_var = calloc(bc, 8);
    {
        double p1 = pow(0.1, 2.5 / sdiv), p2 = pow(0.1, 1.0 / sdiv), min, max; int j = 0;
        for (int m = 0; m < nb; m++)
            for (int nn = 0; nn < nb; nn++) {
                if (nn < m) s->spread_function[j++] = pow(p2, m - nn);
                else if (nn > m) s->spread_function[j++] = pow(p1, nn - m);
                else s->spread_function[j++] = 1.0;
            }
        for (int m = 0; m < bc; m++) s->band_excit[s->bin2band[m]] += 1.0;
        j = 0;
        for (int m = 0; m < nb; m++) for (int nn = 0; nn < nb; nn++) s->prior_band_excit[m] += s->spread_function[j++] * s->band_excit[nn];
        min = pow(0.1, 2.5); max = pow(0.1, 1.0);
        for (int i = 0; i < nb; i++) {
            if (i < lrint(12.0 * sdiv)) s->band_excit[i] = pow(0.1, 1.45 + 0.1 * i / sdiv);
            else s->band_excit[i] = pow(0.1, 2.5 - 0.2 * (i / sdiv - 14.0));
            s->band_excit[i] = s->band_excit[i] < min ? min : s->band_excit[i] > max ? max : s->band_excit[i];
        }
        j = 0;
        for (int i = 0; i < nb; i++) for (int k = 0; k < nb; k++) s->spread_function[j++] *= s->band_excit[i] / s->prior_band_excit[i];
        /* the recursion state starts from zero */
        for (int m = 0; m < nb; m++) { s->band_excit[m] = 0.0; s->prior_band_excit[m] = 0.0; }
    }
    {
        int j = 0; const double sar = s->sample_advance / s->sample_rate;
        for (int i = 0; i < bc; i++) {
            if ((i == s->fft_length2) || (s->bin2band[i] > j)) {
                double d6 = (i - 1) * s->sample_rate / s->fft_length;
                double d7 = fmin(0.008 + 2.2 / d6, 0.03);
                s->band_alpha[j] = exp(-sar / d7);
                s->band_beta[j] = 1.0 - s->band_alpha[j];
                j = s->bin2band[i];
            }
        }
    }
    {
        double wscale = sqrt(8.0 / (9.0 * s->fft_length)), sum = 0.0;
        for (int i = 0; i < s->window_length; i++) {
            double d10 = sin(i * M_PI / s->window_length);
            d10 *= wscale * d10; s->window[i] = d10; sum += d10 * d10;
        }
        s->window_weight = 0.5 * sum;
        s->floor = (double)(1LL << 48) * exp(-23.025558369790467) * s->window_weight;
    }
    s->noise_reduction = nr; s->noise_floor = nf; s->last_noise_reduction = -1e300; s->last_noise_floor = -1e300;
    set_parameters(s, 1);

    const int A = s->sample_advance, W = s->window_length, offset = W - A, FL = s->fft_length;
    float *win = calloc(W, sizeof(float));
    orc_cf *buf = malloc(sizeof(orc_cf) * FL);
    int first = 1;
    for (int64_t pos = 0; pos < n; pos += A) {
        const int nbs = (int)(n - pos < A ? n - pos : A);
        memmove(win, win + A, offset * sizeof(float));
        memcpy(win + offset, x + pos, nbs * sizeof(float));
        memset(win + offset + nbs, 0, (A - nbs) * sizeof(float));
        for (int m = 0; m < W; m++) { buf[m].re = (float)(s->window[m] * win[m] * (double)(1LL << 23)); buf[m].im = 0.f; }
        for (int m = W; m < FL; m++) { buf[m].re = 0.f; buf[m].im = 0.f; }
        orc_fft_f32(buf, FL, 0);
        process_frame(s, buf, first); first = 0;
        /* inverse real transform of the (Hermitian) half spectrum, unnormalised */
        for (int m = 1; m < s->fft_length2; m++) { buf[FL - m].re = buf[m].re; buf[FL - m].im = -buf[m].im; }
        buf[0].im = 0.f; buf[s->fft_length2].im = 0.f;
        orc_fft_f32(buf, FL, 1);
        for (int m = 0; m < W; m++) s->out_samples[m] += s->window[m] * buf[m].re / (double)(1LL << 23);
        for (int m = 0; m < nbs; m++) y[pos + m] = (float)s->out_samples[m];
        memmove(s->out_samples, s->out_samples + A, (size_t)(W - A) * sizeof(double));
        memset(s->out_samples + (W - A), 0, (size_t)A * sizeof(double));
    }
    free(win); free(buf); free(s->window); free(s->bin2band); free(s->band_alpha); free(s->band_beta);
    free(s->amt); free(s->band_amt); free(s->band_excit); free(s->gain); free(s->prior); free(s->prior_band_excit);
    free(s->clean_data); free(s->noisy_data); free(s->out_samples); free(s->spread_function); free(s->abs_var);
    free(s->rel_var); free(s->min_abs_var);
    return W - A;
}
