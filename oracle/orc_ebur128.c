/*
 * ORACLE (test infrastructure only) -- restatement of FFmpeg's ebur128 filter as the
 * reference instantiates it: "ebur128=metadata=1:peak=sample+true:dualmono=true[:target=-16]"
 * (internal/processor/filters.go:626,684-689; analyser_output.go:18 without dualmono).
 * Follows libavfilter/f_ebur128.c (config_audio_input K-weighting, filter_frame,
 * gate_update, histogram gating, LRA) -- parity unpinned (see orc.h); true peak uses the
 * pinned swr restatement (orc_swr.c).  Mono only: every graph starts with
 * aformat=channel_layouts=mono.
 */
#include "orc.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define ABS_THRES    (-70)
#define ABS_UP_THRES 10
#define HIST_RES     100
#define HIST_SIZE    ((ABS_UP_THRES - ABS_THRES) * HIST_RES + 1)
#define HIST_POS(l)  (int)(((l) - ABS_THRES) * HIST_RES)
#define ENERGY(l)    (pow(10.0, ((l) + 0.691) / 10.))
#define LOUDNESS(e)  (-0.691 + 10 * log10(e))

typedef struct { double *cache; int cache_pos, cache_size, filled; double sum;
                 unsigned *hist_count; double sum_kept; uint64_t nb_kept; double rel_threshold; } integ;

static int clipi(int v, int lo, int hi) { return v < lo ? lo : v > hi ? hi : v; }

static int gate_update(integ *g, double power, double loudness, int gate_thres)
{
    int ip = clipi(HIST_POS(loudness), 0, HIST_SIZE - 1);
    g->hist_count[ip]++;
    g->sum_kept += power;
    g->nb_kept++;
    double rt = g->sum_kept / g->nb_kept;
    if (!rt) rt = 1e-12;
    g->rel_threshold = LOUDNESS(rt) + gate_thres;
    return clipi(HIST_POS(g->rel_threshold), 0, HIST_SIZE - 1);
}

int orc_ebur128(const double *x, int64_t n, int rate, int dualmono, int want_true_peak,
                double *M, double *S, double *sp_cum, double *tp_cum, int64_t tick_cap,
                orc_r128_summary *sum)
{
    /* K-weighting, recomputed for the link rate (f_ebur128.c config_audio_input) */
    double f0 = 1681.974450955533, G = 3.999843853973347, Q = 0.7071752369554196;
    double K = tan(M_PI * f0 / (double)rate);
    double Vh = pow(10.0, G / 20.0), Vb = pow(Vh, 0.4996667741545416);
    double a0 = 1.0 + K / Q + K * K;
    double pre_b[3], pre_a[3], rlb_b[3], rlb_a[3];
    pre_b[0] = (Vh + Vb * K / Q + K * K) / a0;
    pre_b[1] = 2.0 * (K * K - Vh) / a0;
    pre_b[2] = (Vh - Vb * K / Q + K * K) / a0;
    pre_a[1] = 2.0 * (K * K - 1.0) / a0;
    pre_a[2] = (1.0 - K / Q + K * K) / a0;
    f0 = 38.13547087602444; Q = 0.5003270373238773;
    K = tan(M_PI * f0 / (double)rate);
    rlb_b[0] = 1.0; rlb_b[1] = -2.0; rlb_b[2] = 1.0;
    rlb_a[1] = 2.0 * (K * K - 1.0) / (1.0 + K / Q + K * K);
    rlb_a[2] = (1.0 - K / Q + K * K) / (1.0 + K / Q + K * K);

    const int tick = rate / 10 > 0 ? rate / 10 : 1;
    integ i400 = {0}, i3000 = {0};
    i400.cache_size = rate * 4 / 10;
    i3000.cache_size = rate * 3;
    i400.cache = calloc(i400.cache_size, sizeof(double));
    i3000.cache = calloc(i3000.cache_size, sizeof(double));
    i400.hist_count = calloc(HIST_SIZE, sizeof(unsigned));
    i3000.hist_count = calloc(HIST_SIZE, sizeof(unsigned));
    double *hist_energy = malloc(sizeof(double) * HIST_SIZE), *hist_loud = malloc(sizeof(double) * HIST_SIZE);
    for (int i = 0; i < HIST_SIZE; i++) {
        hist_loud[i] = i / (double)HIST_RES + ABS_THRES;
        hist_energy[i] = ENERGY(hist_loud[i]);
    }
    const double pan_law = -3.01029995663978;

    /* true peak: whole-stream 192 kHz oversample; cumulative max per tick taken over
     * the outputs swr has produced once (k+1)*tick inputs were fed (no flush) */
    double *up = NULL; int64_t n_up = 0;
    if (want_true_peak && rate != 192000) {
        int64_t cap = orc_swr_out_count(n, rate, 192000);
        if (cap < 0) return -1;
        up = malloc(sizeof(double) * (size_t)(cap > 0 ? cap : 1));
        n_up = orc_swr_resample_f64(x, n, rate, 192000, 0, up, cap);
    }

    double xs[3] = {0}, ys[3] = {0}, zs[3] = {0};
    double sample_peak = 0, true_peak = 0;
    double integrated = ABS_THRES, lra = 0, lra_low = 0, lra_high = 0;
    int sample_count = 0;
    int64_t ticks = 0, up_done = 0;

    for (int64_t idx = 0; idx < n; idx++) {
        if (idx % tick == 0 && want_true_peak) {
            /* top of filter_frame: swr_convert of this (possibly short) frame */
            int64_t fed = idx + tick < n ? idx + tick : n;
            int64_t avail = up ? orc_swr_out_count(fed, rate, 192000) : 0;
            if (!up) { /* 192 kHz link: swr still runs as a pass-through copy */
                for (int64_t j = idx; j < fed; j++) if (fabs(x[j]) > true_peak) true_peak = fabs(x[j]);
            }
            if (avail > n_up) avail = n_up;
            for (; up_done < avail; up_done++) if (fabs(up[up_done]) > true_peak) true_peak = fabs(up[up_done]);
        }
        const int b400 = i400.cache_pos, b3000 = i3000.cache_pos;
        if (++i400.cache_pos == i400.cache_size) { i400.filled = 1; i400.cache_pos = 0; }
        if (++i3000.cache_pos == i3000.cache_size) { i3000.filled = 1; i3000.cache_pos = 0; }
        if (fabs(x[idx]) > sample_peak) sample_peak = fabs(x[idx]);
        xs[0] = x[idx];
        ys[2] = ys[1]; ys[1] = ys[0];
        ys[0] = xs[0] * pre_b[0] + xs[1] * pre_b[1] + xs[2] * pre_b[2] - ys[1] * pre_a[1] - ys[2] * pre_a[2];
        xs[2] = xs[1]; xs[1] = xs[0];
        zs[2] = zs[1]; zs[1] = zs[0];
        zs[0] = ys[0] * rlb_b[0] + ys[1] * rlb_b[1] + ys[2] * rlb_b[2] - zs[1] * rlb_a[1] - zs[2] * rlb_a[2];
        double bin = zs[0] * zs[0];
        i400.sum = i400.sum + bin - i400.cache[b400];
        i3000.sum = i3000.sum + bin - i3000.cache[b3000];
        i400.cache[b400] = bin;
        i3000.cache[b3000] = bin;

        if (++sample_count == tick) {
            double p400 = 1e-12, p3000 = 1e-12, l400, l3000;
            sample_count = 0;
            if (i400.filled) { p400 += 1.0 * i400.sum; p400 /= i400.cache_size; }
            l400 = LOUDNESS(p400);
            if (i3000.filled) { p3000 += 1.0 * i3000.sum; p3000 /= i3000.cache_size; }
            l3000 = LOUDNESS(p3000);

            if (l400 >= ABS_THRES) {
                double isum = 0.0; uint64_t nb = 0;
                int pos = gate_update(&i400, p400, l400, -10);
                for (int i = pos; i < HIST_SIZE; i++) {
                    unsigned c = i400.hist_count[i];
                    nb += c; isum += c * hist_energy[i];
                }
                if (nb) {
                    integrated = LOUDNESS(isum / nb);
                    if (dualmono) integrated -= pan_law;
                }
            }
            if (l3000 >= ABS_THRES) {
                uint64_t nb_powers = 0;
                int pos = gate_update(&i3000, p3000, l3000, -20);
                for (int i = pos; i < HIST_SIZE; i++) nb_powers += i3000.hist_count[i];
                if (nb_powers) {
                    uint64_t nn = 0, nb_pow = 10 * nb_powers * 0.01 + 0.5;
                    for (int i = pos; i < HIST_SIZE; i++) {
                        nn += i3000.hist_count[i];
                        if (nn >= nb_pow) { lra_low = hist_loud[i]; break; }
                    }
                    nn = nb_powers;
                    nb_pow = 95 * nb_powers * 0.01 + 0.5;
                    for (int i = HIST_SIZE - 1; i >= 0; i--) {
                        uint64_t c = i3000.hist_count[i];
                        nn -= (nn < c ? nn : c);
                        if (nn < nb_pow) { lra_high = hist_loud[i]; break; }
                    }
                    lra = lra_high - lra_low;
                }
            }
            if (dualmono) { l400 -= pan_law; l3000 -= pan_law; }
            if (ticks < tick_cap) {
                if (M) M[ticks] = l400;
                if (S) S[ticks] = l3000;
                if (sp_cum) sp_cum[ticks] = sample_peak;
                if (tp_cum) tp_cum[ticks] = true_peak;
            }
            sum->sample_peak = sample_peak;
            sum->true_peak = true_peak;
            ticks++;
        }
    }
    sum->n_ticks = ticks;
    sum->I = integrated; sum->LRA = lra; sum->LRA_low = lra_low; sum->LRA_high = lra_high;
    sum->rel_threshold_400 = i400.rel_threshold;
    if (!ticks) { sum->sample_peak = 0; sum->true_peak = 0; }
    free(i400.cache); free(i3000.cache); free(i400.hist_count); free(i3000.hist_count);
    free(hist_energy); free(hist_loud); free(up);
    return 0;
}
