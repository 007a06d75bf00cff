/*
 * ORACLE (test infrastructure only) -- restatement of FFmpeg's aspectralstats filter for
 * one channel: "aspectralstats=win_size=2048:win_func=hann:measure=all"
 * (internal/processor/filters.go:625) and "aspectralstats=measure=all" (defaults,
 * analyser_output.go:18).  Follows libavfilter/af_aspectralstats.c filter_channel() and
 * the spectral_* helpers; formulas are tabulated in docs/Spectral-Metrics-Reference.md:9-33.
 * float32 arithmetic throughout, as upstream.  Parity unpinned (see orc.h).
 */
#include "orc.h"
#include "orc_fft.h"
#include <math.h>
#include <float.h>
#include <stdlib.h>
#include <string.h>

static float sqrf(float a) { return a * a; }
static float cbrf(float a) { return a * a * a; }

int64_t orc_aspectralstats(const float *x, int64_t n, int rate, int win_size, float *rows, int64_t hop_cap)
{
    const int hop = (int)(win_size * (1.f - 0.5f));
    const int size = win_size / 2;
    const int max_freq = rate / 2;
    const int offset = win_size - hop;
    float *lut = malloc(sizeof(float) * win_size);
    float *window = calloc(win_size, sizeof(float));
    float *mag = calloc(win_size, sizeof(float)), *prev = calloc(win_size, sizeof(float));
    orc_cf *buf = malloc(sizeof(orc_cf) * win_size);
    /* window_func.h WFUNC_HANNING: .5*(1-cos(2*M_PI*n/(N-1))) */
    for (int i = 0; i < win_size; i++) lut[i] = .5 * (1 - cos(2 * M_PI * i / (win_size - 1)));
    const float wscale = 1.f / win_size;
    int64_t hops = 0;

    for (int64_t pos = 0; pos < n; pos += hop) {
        int nb = (int)(n - pos < hop ? n - pos : hop);
        memmove(window, &window[hop], offset * sizeof(float));
        memcpy(&window[offset], x + pos, nb * sizeof(float));
        memset(&window[offset + nb], 0, (hop - nb) * sizeof(float));
        for (int i = 0; i < win_size; i++) { buf[i].re = window[i] * lut[i]; buf[i].im = 0; }
        orc_fft_f32(buf, win_size, 0);
        for (int i = 0; i < size; i++) { buf[i].re *= wscale; buf[i].im *= wscale; }
        for (int i = 0; i < size; i++) mag[i] = hypotf(buf[i].re, buf[i].im);

        float *r = hops < hop_cap ? rows + hops * ORC_NSPEC : NULL;
        if (r) {
            const float scale = max_freq / (float)size;
            float sum, num, den, mean, centroid, spread;
            /* mean */
            sum = 0.f; for (int i = 0; i < size; i++) sum += mag[i];
            mean = sum / size; r[0] = mean;
            /* variance */
            sum = 0.f; for (int i = 0; i < size; i++) sum += sqrf(mag[i] - mean);
            r[1] = sum / size;
            /* centroid */
            num = den = 0.f; for (int i = 0; i < size; i++) { num += mag[i] * i * scale; den += mag[i]; }
            centroid = den <= FLT_EPSILON ? 1.f : num / den; r[2] = centroid;
            /* spread */
            num = den = 0.f; for (int i = 0; i < size; i++) { num += mag[i] * sqrf(i * scale - centroid); den += mag[i]; }
            spread = den <= FLT_EPSILON ? 1.f : sqrtf(num / den); r[3] = spread;
            /* skewness */
            num = den = 0.f; for (int i = 0; i < size; i++) { num += mag[i] * cbrf(i * scale - centroid); den += mag[i]; }
            den *= cbrf(spread); r[4] = den <= FLT_EPSILON ? 1.f : num / den;
            /* kurtosis */
            num = den = 0.f; for (int i = 0; i < size; i++) { num += mag[i] * sqrf(sqrf(i * scale - centroid)); den += mag[i]; }
            den *= sqrf(sqrf(spread)); r[5] = den <= FLT_EPSILON ? 1.f : num / den;
            /* entropy */
            num = 0.f; for (int i = 0; i < size; i++) num += mag[i] * logf(mag[i] + FLT_EPSILON);
            den = logf(size); r[6] = den <= FLT_EPSILON ? 1.f : -num / den;
            /* flatness */
            num = den = 0.f; for (int i = 0; i < size; i++) { float v = FLT_EPSILON + mag[i]; num += logf(v); den += v; }
            num /= size; den /= size; num = expf(num); r[7] = den <= FLT_EPSILON ? 0.f : num / den;
            /* crest */
            { float mx = 0.f, mn = 0.f; for (int i = 0; i < size; i++) { mx = fmaxf(mx, mag[i]); mn += mag[i]; }
              mn /= size; r[8] = mn <= FLT_EPSILON ? 0.f : mx / mn; }
            /* flux */
            sum = 0.f; for (int i = 0; i < size; i++) sum += sqrf(mag[i] - prev[i]);
            r[9] = sqrtf(sum);
            /* slope */
            { const float mean_freq = size * 0.5f; float ms = 0.f; num = den = 0.f;
              for (int i = 0; i < size; i++) ms += mag[i];
              ms /= size;
              for (int i = 0; i < size; i++) { num += ((i - mean_freq) / mean_freq) * (mag[i] - ms); den += sqrf((i - mean_freq) / mean_freq); }
              r[10] = fabsf(den) <= FLT_EPSILON ? 0.f : num / den; }
            /* decrease */
            num = den = 0.f; for (int i = 1; i < size; i++) { num += (mag[i] - mag[0]) / i; den += mag[i]; }
            r[11] = den <= FLT_EPSILON ? 0.f : num / den;
            /* rolloff */
            { float norm = 0.f; int idx = 0; sum = 0.f;
              for (int i = 0; i < size; i++) norm += mag[i];
              norm *= 0.85f;
              for (int i = 0; i < size; i++) { sum += mag[i]; if (sum >= norm) { idx = i; break; } }
              r[12] = scale * idx; }
        }
        memcpy(prev, mag, win_size * sizeof(float));
        hops++;
    }
    free(lut); free(window); free(mag); free(prev); free(buf);
    return hops;
}
