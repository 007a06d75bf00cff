/*
 * ORACLE (test infrastructure only) -- restatement of FFmpeg's astats filter for one
 * channel, as instantiated by "astats=metadata=1:measure_perchannel=all" /
 * "measure_perchannel=0" (internal/processor/filters.go:624, analyser_bands.go:33,
 * analyser_output.go:18).  Follows libavfilter/af_astats.c update_stat()/set_metadata()
 * (length=0.05 s default, reset=0) -- parity unpinned (see orc.h).  Formulas are also
 * tabulated in docs/Spectral-Metrics-Reference.md:35-56.
 */
#include "orc.h"
#include <math.h>
#include <float.h>
#include <stdlib.h>
#include <string.h>

#define HISTOGRAM_SIZE 8192
#define HISTOGRAM_MAX  (HISTOGRAM_SIZE - 1)
#define LINEAR_TO_DB(x) (log10(x) * 20)
#define FFSIGN(a) ((a) > 0 ? 1 : -1)

int orc_astats(const void *xv, int fmt, int64_t n, int rate, orc_astats_out *o)
{
    const double time_constant = 0.05;
    const int tc_samples = (int)fmax(time_constant * rate + .5, 1);
    const double mult = exp((-1 / time_constant / rate));
    const int maxbitdepth = fmt == ORC_FMT_S16 ? 16 : (fmt == ORC_FMT_FLT || fmt == ORC_FMT_S32) ? 32 : 64;

    double min = DBL_MAX, max = -DBL_MAX, nmin = DBL_MAX, nmax = -DBL_MAX;
    double min_non_zero = DBL_MAX, min_diff = DBL_MAX, max_diff = 0;
    double sigma_x = 0, sigma_x2 = 0, avg_sigma_x2 = 0, min_sigma_x2 = DBL_MAX, max_sigma_x2 = -DBL_MAX;
    double diff1_sum = 0, diff1_sum_x2 = 0, last = NAN, last_non_zero = 0;
    double min_run = 0, max_run = 0, min_runs = 0, max_runs = 0;
    double noise_floor = NAN;
    uint64_t min_count = 0, max_count = 0, zero_runs = 0, nb_samples = 0, noise_floor_count = 0;
    uint64_t mask0 = 0;
    uint64_t *ehist = calloc(HISTOGRAM_SIZE, sizeof(uint64_t));
    /* sliding-window max of |nd| over tc_samples (af_astats.c calc_noise_floor keeps a
     * monotonic deque in sorted_samples); restated with an index deque */
    int64_t *dq = malloc(sizeof(int64_t) * (size_t)(tc_samples + 1));
    double *dqv = malloc(sizeof(double) * (size_t)(tc_samples + 1));
    int dq_head = 0, dq_len = 0;

    for (int64_t k = 0; k < n; k++) {
        double d, nd; int64_t iv;
        if (fmt == ORC_FMT_S16) { int16_t s = ((const int16_t *)xv)[k]; d = s; nd = s / (double)INT16_MAX; iv = s; }
        else if (fmt == ORC_FMT_S32) { int32_t s = ((const int32_t *)xv)[k]; d = s; nd = s / (double)INT32_MAX; iv = s; }
        else if (fmt == ORC_FMT_FLT) { float s = ((const float *)xv)[k]; d = s; nd = s; iv = llrint(s * (double)(UINT64_C(1) << 31)); }
        else { double s = ((const double *)xv)[k]; d = s; nd = s;
               double t = s * 9223372036854775808.0;
               iv = t >= 9223372036854775807.0 ? INT64_MAX : t <= -9223372036854775808.0 ? INT64_MIN : llrint(t); }

        if (d < min) { min = d; nmin = nd; min_run = 1; min_runs = 0; min_count = 1; }
        else if (d == min) { min_count++; min_run = d == last ? min_run + 1 : 1; }
        else if (last == min) { min_runs += min_run * min_run; }

        if (d != 0 && fabs(d) < min_non_zero) min_non_zero = fabs(d);

        if (d > max) { max = d; nmax = nd; max_run = 1; max_runs = 0; max_count = 1; }
        else if (d == max) { max_count++; max_run = d == last ? max_run + 1 : 1; }
        else if (last == max) { max_runs += max_run * max_run; }

        if (d != 0) { zero_runs += FFSIGN(d) != FFSIGN(last_non_zero); last_non_zero = d; }

        sigma_x += nd;
        sigma_x2 += nd * nd;
        avg_sigma_x2 = avg_sigma_x2 * mult + (1.0 - mult) * nd * nd;
        if (!isnan(last)) {
            double ad = fabs(d - last);
            if (ad < min_diff) min_diff = ad;
            if (ad > max_diff) max_diff = ad;
            diff1_sum += ad;
            diff1_sum_x2 += (d - last) * (d - last);
        }
        mask0 |= (uint64_t)(iv < 0 ? -iv : iv);
        last = d;

        {
            double a = fabs(nd); if (a > 1.0) a = 1.0;
            long idx = lrint(a * HISTOGRAM_MAX);
            if (idx < 0) idx = 0;
            if (idx > HISTOGRAM_MAX) idx = HISTOGRAM_MAX;
            ehist[idx]++;
        }
        if (nb_samples >= (uint64_t)tc_samples) {
            if (avg_sigma_x2 > max_sigma_x2) max_sigma_x2 = avg_sigma_x2;
            if (avg_sigma_x2 < min_sigma_x2) min_sigma_x2 = avg_sigma_x2;
        }
        nb_samples++;

        /* window max of |nd| over the last tc_samples samples (including this one) */
        {
            double a = fabs(nd);
            while (dq_len > 0 && dqv[(dq_head + dq_len - 1) % (tc_samples + 1)] <= a) dq_len--;
            int slot = (dq_head + dq_len) % (tc_samples + 1);
            dq[slot] = k; dqv[slot] = a; dq_len++;
            while (dq[dq_head] <= k - tc_samples) { dq_head = (dq_head + 1) % (tc_samples + 1); dq_len--; }
            double wmax = dqv[dq_head];
            if (nb_samples >= (uint64_t)tc_samples) {
                if (isnan(noise_floor)) { noise_floor = wmax; noise_floor_count = 1; }
                else if (wmax < noise_floor) { noise_floor = wmax; noise_floor_count = 1; }
                else if (wmax == noise_floor) noise_floor_count++;
            }
        }
    }

    memset(o, 0, sizeof(*o));
    o->nb_samples = (double)nb_samples;
    if (nb_samples == 0) { free(ehist); free(dq); free(dqv); return 0; }
    if (nb_samples < (uint64_t)tc_samples)
        min_sigma_x2 = max_sigma_x2 = sigma_x2 / nb_samples;

    o->DC_offset = sigma_x / nb_samples;
    o->Min_level = min;
    o->Max_level = max;
    o->Min_difference = min_diff;
    o->Max_difference = max_diff;
    o->Mean_difference = diff1_sum / (nb_samples - 1);
    o->RMS_difference = sqrt(diff1_sum_x2 / (nb_samples - 1));
    o->Peak_level = LINEAR_TO_DB(fmax(-nmin, nmax));
    o->RMS_level = LINEAR_TO_DB(sqrt(sigma_x2 / nb_samples));
    o->RMS_peak = LINEAR_TO_DB(sqrt(max_sigma_x2));
    o->RMS_trough = LINEAR_TO_DB(sqrt(min_sigma_x2));
    o->Crest_factor = sigma_x2 ? fmax(-min, max) / sqrt(sigma_x2 / nb_samples) : 1;
    o->Flat_factor = LINEAR_TO_DB((min_runs + max_runs) / (min_count + max_count));
    o->Peak_count = (float)(min_count + max_count);
    o->Noise_floor = LINEAR_TO_DB(noise_floor);
    o->Noise_floor_count = (double)noise_floor_count;
    {
        double e = 0;
        for (int i = 0; i < HISTOGRAM_SIZE; i++) {
            double entry = ehist[i] / ((double)nb_samples);
            if (entry > 1e-8) e += entry * log2(entry);
        }
        o->Entropy = -e / log2(HISTOGRAM_SIZE);
    }
    {
        int depth = 0;
        for (int i = 0; i < maxbitdepth; i++) depth += !!(mask0 & (1ULL << i));
        o->Bit_depth = depth;
    }
    o->Dynamic_range = LINEAR_TO_DB(2 * fmax(fabs(min), fabs(max)) / min_non_zero);
    o->Zero_crossings = (double)zero_runs;
    o->Zero_crossings_rate = zero_runs / (double)nb_samples;
    free(ehist); free(dq); free(dqv);
    return 0;
}
