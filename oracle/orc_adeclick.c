/*
 * ORACLE (test infrastructure only) -- restatement of FFmpeg's adeclick as the reference
 * instantiates it: "adeclick=t=1.7:w=55:o=50:m=s" (internal/processor/filters.go:947-962,
 * 513-521).  Follows libavfilter/af_adeclick.c: config_input(), autoregression() (Levinson-
 * Durbin), detect_clicks(), interpolation() (LDL^T solve), filter_channel(), the overlap-save
 * ("m=s") fifo cadence of activate()/filter_frame() incl. the EOF tail.  f64, mono; every
 * window is independent given the input.  Parity unpinned (see orc.h).
 */
#include "orc.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

static void autocorrelation(const double *input, int order, int size, double *output, double scale)
{
    for (int i = 0; i <= order; i++) {
        double value = 0.;
        for (int j = i; j < size; j++) value += input[j] * input[j - i];
        output[i] = value * scale;
    }
}

static double autoregression(const double *samples, int ar_order, int nb_samples, double *k, double *r, double *a)
{
    double alpha;
    memset(a, 0, ar_order * sizeof(*a));
    autocorrelation(samples, ar_order, nb_samples, r, 1. / nb_samples);
    k[0] = a[0] = -r[1] / r[0];
    alpha = r[0] * (1. - k[0] * k[0]);
    for (int i = 1; i < ar_order; i++) {
        double epsilon = 0.;
        for (int j = 0; j < i; j++) epsilon += a[j] * r[i - j];
        epsilon += r[i + 1];
        k[i] = -epsilon / alpha;
        alpha *= (1. - k[i] * k[i]);
        for (int j = i - 1; j >= 0; j--) k[j] = a[j] + k[i] * a[i - j - 1];
        for (int j = 0; j <= i; j++) a[j] = k[j];
    }
    k[0] = 1.;
    for (int i = 1; i <= ar_order; i++) k[i] = a[i - 1];
    return sqrt(alpha);
}

static int isfinite_array(const double *samples, int nb_samples)
{
    for (int i = 0; i < nb_samples; i++) if (!isfinite(samples[i])) return 0;
    return 1;
}

static int find_index(const int *index, int value, int size)
{
    int i, start, end;
    if ((value < index[0]) || (value > index[size - 1])) return 1;
    i = start = 0; end = size - 1;
    while (start <= end) {
        i = (end + start) / 2;
        if (index[i] == value) return 0;
        if (value < index[i]) end = i - 1;
        if (value > index[i]) start = i + 1;
    }
    return 1;
}

static int factorization(double *matrix, int n)
{
    for (int i = 0; i < n; i++) {
        const int in = i * n;
        double value = matrix[in + i];
        for (int j = 0; j < i; j++) value -= matrix[j * n + j] * matrix[in + j] * matrix[in + j];
        if (value == 0.) return -1;
        matrix[in + i] = value;
        for (int j = i + 1; j < n; j++) {
            const int jn = j * n;
            double x = matrix[jn + i];
            for (int k = 0; k < i; k++) x -= matrix[k * n + k] * matrix[in + k] * matrix[jn + k];
            matrix[jn + i] = x / matrix[in + i];
        }
    }
    return 0;
}

static int do_interpolation(double *matrix, double *vector, int n, double *out)
{
    if (factorization(matrix, n) < 0) return -1;
    double *y = malloc(sizeof(double) * n);
    for (int i = 0; i < n; i++) {
        const int in = i * n;
        double value = vector[i];
        for (int j = 0; j < i; j++) value -= matrix[in + j] * y[j];
        y[i] = value;
    }
    for (int i = n - 1; i >= 0; i--) {
        out[i] = y[i] / matrix[i * n + i];
        for (int j = i + 1; j < n; j++) out[i] -= matrix[j * n + i] * out[j];
    }
    free(y);
    return 0;
}

int64_t orc_adeclick(const double *x, double *y, int64_t n, int rate, double w_ms, double overlap_pct, double ar_pct,
                     double threshold, double burst_pct, int method_save, int64_t *detected)
{
    const int window_size = (int)(rate * w_ms / 1000.);
    if (window_size < 100) return -1;
    int ar_order = (int)(window_size * ar_pct / 100.); if (ar_order < 1) ar_order = 1;
    const int nb_burst_samples = (int)(window_size * burst_pct / 1000.);
    const int hop_size = (int)(window_size * (1. - (overlap_pct / 100.)));
    if (hop_size < 1) return -1;
    if (!method_save) return -2;                                   /* overlap-add path not restated */
    const int overlap_skip = (window_size - hop_size) / 2;

    double *in = calloc(window_size, sizeof(double)), *dst = calloc(window_size, sizeof(double));
    double *detection = calloc(window_size, sizeof(double)), *interpolated = calloc(window_size, sizeof(double));
    double *acoef = calloc(ar_order + 1, sizeof(double)), *acorr = calloc(ar_order + 1, sizeof(double));
    double *tmp = calloc(ar_order + 1, sizeof(double)), *aux = calloc(ar_order + 1, sizeof(double));
    unsigned char *click = calloc(window_size, 1);
    int *index = calloc(window_size, sizeof(int));
    int64_t det = 0, written = 0;

    /* fifo = overlap_skip zeros followed by the stream; window w peeks fifo[w*hop, w*hop+window) */
    const int64_t fifo_len = overlap_skip + n;
    for (int64_t w = 0;; w++) {
        const int64_t fpos = w * (int64_t)hop_size;
        const int64_t avail = fifo_len - fpos;                     /* samples the peek can deliver */
        if (avail < window_size) {
            /* EOF: samples_left = fifo_size - overlap_skip at the moment EOF is seen; keep going while > 0 */
            if (written >= n) break;
        }
        for (int j = 0; j < window_size; j++) {
            const int64_t s = fpos + j - overlap_skip;
            if (fpos + j < fifo_len) in[j] = s < 0 ? 0.0 : x[s];
            /* else: av_audio_fifo_peek() delivers fewer samples; the buffer keeps its previous content */
        }
        double sigmae = autoregression(in, ar_order, window_size, acoef, acorr, tmp);
        if (isfinite_array(acoef, ar_order + 1)) {
            int nb_clicks = 0, prev = -1;
            memset(detection, 0, window_size * sizeof(double));
            for (int i = ar_order; i < window_size; i++)
                for (int j = 0; j <= ar_order; j++) detection[i] += acoef[j] * in[i - j];
            for (int i = 0; i < window_size; i++) { click[i] = fabs(detection[i]) > sigmae * threshold; dst[i] = in[i]; }
            for (int i = 0; i < window_size; i++) {
                if (!click[i]) continue;
                if (prev >= 0 && (i > prev + 1) && (i <= nb_burst_samples + prev))
                    for (int j = prev + 1; j < i; j++) click[j] = 1;
                prev = i;
            }
            memset(click, 0, ar_order);
            memset(click + (window_size - ar_order), 0, ar_order);
            for (int i = ar_order; i < window_size - ar_order; i++) if (click[i]) index[nb_clicks++] = i;
            if (nb_clicks > 0) {
                double *matrix = malloc(sizeof(double) * (size_t)nb_clicks * nb_clicks), *vector = malloc(sizeof(double) * nb_clicks);
                autocorrelation(acoef, ar_order, ar_order + 1, aux, 1.);
                for (int i = 0; i < nb_clicks; i++) {
                    const int im = i * nb_clicks;
                    for (int j = i; j < nb_clicks; j++) {
                        if (abs(index[j] - index[i]) <= ar_order) matrix[j * nb_clicks + i] = matrix[im + j] = aux[abs(index[j] - index[i])];
                        else matrix[j * nb_clicks + i] = matrix[im + j] = 0;
                    }
                }
                for (int i = 0; i < nb_clicks; i++) {
                    double value = 0.;
                    for (int j = -ar_order; j <= ar_order; j++)
                        if (find_index(index, index[i] - j, nb_clicks)) value -= in[index[i] - j] * aux[abs(j)];
                    vector[i] = value;
                }
                if (do_interpolation(matrix, vector, nb_clicks, interpolated) == 0)
                    for (int j = 0; j < nb_clicks; j++) {
                        dst[index[j]] = interpolated[j];
                        if (index[j] >= overlap_skip && index[j] < overlap_skip + hop_size) det++;
                    }
                free(matrix); free(vector);
            }
        } else memcpy(dst, in, window_size * sizeof(double));
        for (int j = 0; j < hop_size && written < n; j++) y[written++] = dst[overlap_skip + j];
        if (written >= n) break;
    }
    if (detected) *detected = det;
    free(in); free(dst); free(detection); free(interpolated); free(acoef); free(acorr); free(tmp); free(aux); free(click); free(index);
    return written;
}
