/*
 * ORACLE (test infrastructure only) -- the BS.1770 meter FFmpeg's loudnorm uses for its
 * input_i / input_tp / input_lra / input_thresh JSON fields (internal/processor/normalise.go:64-75,
 * 328-343): libavfilter/ebur128.c (libebur128 port): 4th-order direct-form-II K-weighting,
 * 400 ms gating blocks every 100 ms kept as an energy LIST (no histogram), relative gate -10 LU,
 * LRA from 3 s blocks every 1 s, sample peak.  Mono, FF_EBUR128_DUAL_MONO doubles the energy.
 * Parity unpinned (see orc.h); cross-checked against torchaudio in tests/test_oracle_filters.py.
 */
#include "orc.h"
#include <math.h>
#include <float.h>
#include <stdlib.h>

static int cmp_d(const void *a, const void *b) { double x = *(const double *)a, y = *(const double *)b; return (x > y) - (x < y); }

int orc_loudnorm_meter(const double *x, int64_t n, int rate, int dual_mono, double *out4 /* I, LRA, thresh, sample_peak */)
{
    double f0 = 1681.974450955533, G = 3.999843853973347, Q = 0.7071752369554196;
    double K = tan(M_PI * f0 / (double)rate), Vh = pow(10.0, G / 20.0), Vb = pow(Vh, 0.4996667741545416);
    double pb[3] = {0.0, 0.0, 0.0}, pa[3] = {1.0, 0.0, 0.0}, rb[3] = {1.0, -2.0, 1.0}, ra[3] = {1.0, 0.0, 0.0};
    double a0 = 1.0 + K / Q + K * K, b[5], a[5], v[5] = {0};
    pb[0] = (Vh + Vb * K / Q + K * K) / a0; pb[1] = 2.0 * (K * K - Vh) / a0; pb[2] = (Vh - Vb * K / Q + K * K) / a0;
    pa[1] = 2.0 * (K * K - 1.0) / a0; pa[2] = (1.0 - K / Q + K * K) / a0;
    f0 = 38.13547087602444; Q = 0.5003270373238773; K = tan(M_PI * f0 / (double)rate);
    ra[1] = 2.0 * (K * K - 1.0) / (1.0 + K / Q + K * K); ra[2] = (1.0 - K / Q + K * K) / (1.0 + K / Q + K * K);
    b[0] = pb[0] * rb[0]; b[1] = pb[0] * rb[1] + pb[1] * rb[0]; b[2] = pb[0] * rb[2] + pb[1] * rb[1] + pb[2] * rb[0];
    b[3] = pb[1] * rb[2] + pb[2] * rb[1]; b[4] = pb[2] * rb[2];
    a[0] = pa[0] * ra[0]; a[1] = pa[0] * ra[1] + pa[1] * ra[0]; a[2] = pa[0] * ra[2] + pa[1] * ra[1] + pa[2] * ra[0];
    a[3] = pa[1] * ra[2] + pa[2] * ra[1]; a[4] = pa[2] * ra[2];

    const int s100 = (rate + 5) / 10;
    const int64_t nfull = n / s100;
    double *p100 = calloc(nfull > 0 ? nfull : 1, sizeof(double));
    double peak = 0;
    for (int64_t i = 0; i < n; i++) {
        if (fabs(x[i]) > peak) peak = fabs(x[i]);
        v[0] = x[i] - a[1] * v[1] - a[2] * v[2] - a[3] * v[3] - a[4] * v[4];
        double o = b[0] * v[0] + b[1] * v[1] + b[2] * v[2] + b[3] * v[3] + b[4] * v[4];
        v[4] = v[3]; v[3] = v[2]; v[2] = v[1]; v[1] = v[0];
        if (i / s100 < nfull) p100[i / s100] += o * o;
    }
    const double wgt = dual_mono ? 2.0 : 1.0, abs_thr = pow(10.0, (-70.0 + 0.691) / 10.0);
    double *blocks = malloc(sizeof(double) * (nfull > 0 ? nfull : 1)); int64_t nb = 0; double sum = 0;
    for (int64_t k = 3; k < nfull; k++) {
        double e = wgt * (p100[k - 3] + p100[k - 2] + p100[k - 1] + p100[k]) / (4.0 * s100);
        if (e >= abs_thr) { blocks[nb++] = e; sum += e; }
    }
    double I = -HUGE_VAL, thresh = -70.0, lra = 0;
    if (nb) {
        double rel = sum / nb * 0.1, g = 0; int64_t gc = 0;
        thresh = 10 * log10(rel) - 0.691;
        for (int64_t k = 0; k < nb; k++) if (blocks[k] >= rel) { g += blocks[k]; gc++; }
        if (gc) I = 10 * log10(g / gc) - 0.691;
    }
    double *st = malloc(sizeof(double) * (nfull / 10 + 2)); int64_t ns = 0;
    for (int64_t k = 29; k < nfull; k += 10) {
        double s = 0; for (int j = 29; j >= 0; j--) s += p100[k - j];
        double e = wgt * s / (30.0 * s100);
        if (e >= abs_thr) st[ns++] = e;
    }
    if (ns) {
        qsort(st, ns, sizeof(double), cmp_d);
        double p = 0; for (int64_t k = 0; k < ns; k++) p += st[k]; p /= ns;
        double integ = 0.01 * p; int64_t first = 0;
        while (first < ns && st[first] < integ) first++;
        int64_t sz = ns - first;
        if (sz) {
            double h = st[first + (int64_t)((sz - 1) * 0.95 + 0.5)], l = st[first + (int64_t)((sz - 1) * 0.1 + 0.5)];
            lra = (10 * log10(h) - 0.691) - (10 * log10(l) - 0.691);
        }
    }
    out4[0] = I; out4[1] = lra; out4[2] = thresh; out4[3] = peak;
    free(p100); free(blocks); free(st);
    return 0;
}
