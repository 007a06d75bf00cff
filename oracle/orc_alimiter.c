/*
 * ORACLE (test infrastructure only) -- restatement of FFmpeg's alimiter (Calf lookahead
 * limiter port) as the reference instantiates it:
 *   "alimiter=limit=%.6f:attack=5:release=100:level_in=1:level_out=1:level=0:latency=1:asc=1:asc_level=0.8"
 *   "alimiter=limit=0.803526:attack=1:release=50:..."     (internal/processor/normalise.go:446-480)
 * Follows libavfilter/af_alimiter.c init(), config_input(), get_rdelta(), filter_frame() incl. the
 * latency=1 trimming (output aligned with input, same length).  f64, mono, sequential.
 * Parity unpinned (see orc.h).
 */
#include "orc.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    double limit, attack, release, att, level_in, level_out; int auto_release, auto_level;
    double asc; int asc_c, asc_pos; double asc_coeff;
    double *buffer; int buffer_size, pos; int *nextpos; double *nextdelta; double delta; int nextiter, nextlen, asc_changed;
} Lim;

static double get_rdelta(Lim *s, double release, int sample_rate, double peak, double limit, double patt, int asc)
{
    double rdelta = (1.0 - patt) / (sample_rate * release);
    if (asc && s->auto_release && s->asc_c > 0) {
        double a_att = limit / (s->asc_coeff * s->asc) * (double)s->asc_c;
        if (a_att > patt) {
            double delta = fmax((a_att - patt) / (sample_rate * release), rdelta / 10);
            if (delta < rdelta) rdelta = delta;
        }
    }
    return rdelta;
}

int orc_alimiter(const double *x, double *y, int64_t n, int rate, double limit, double attack_ms, double release_ms,
                 double level_in, double level_out, int auto_level, int asc, double asc_level)
{
    Lim S, *s = &S; memset(s, 0, sizeof(*s));
    const int channels = 1;
    s->limit = limit; s->attack = attack_ms / 1000.; s->release = release_ms / 1000.; s->att = 1.; s->asc_pos = -1;
    s->level_in = level_in; s->level_out = level_out; s->auto_release = asc; s->auto_level = auto_level;
    s->asc_coeff = pow(0.5, asc_level - 0.5) * 2 * -1;
    const int obuffer_size = (int)(rate * channels * 100 / 1000. + channels);
    s->buffer = calloc(obuffer_size, sizeof(double));
    s->nextdelta = calloc(obuffer_size, sizeof(double));
    s->nextpos = malloc(obuffer_size * sizeof(int));
    memset(s->nextpos, -1, obuffer_size * sizeof(int));
    s->buffer_size = (int)(rate * s->attack * channels);
    s->buffer_size -= s->buffer_size % channels;
    if (s->buffer_size <= 0) { free(s->buffer); free(s->nextdelta); free(s->nextpos); return -1; }
    const int latency = s->buffer_size / channels - 1;       /* in_trim = out_pad */
    const int buffer_size = s->buffer_size;
    double *buffer = s->buffer, *nextdelta = s->nextdelta; int *nextpos = s->nextpos;
    const double release = s->release, level = s->auto_level ? 1 / limit : 1;

    /* the filter sees the n input samples followed by out_pad zero samples (request_frame at EOF);
     * the first in_trim outputs are dropped */
    for (int64_t t = 0; t < n + latency; t++) {
        double peak = 0, sample = (t < n ? x[t] : 0.0) * level_in, out;
        int i;
        buffer[s->pos] = sample;
        peak = fmax(peak, fabs(sample));
        if (s->auto_release && peak > limit) { s->asc += peak; s->asc_c++; }
        if (peak > limit) {
            double patt = fmin(limit / peak, 1.);
            double rdelta = get_rdelta(s, release, rate, peak, limit, patt, 0);
            double delta = (limit / peak - s->att) / buffer_size * channels;
            int found = 0;
            if (delta < s->delta) {
                s->delta = delta;
                nextpos[0] = s->pos; nextpos[1] = -1; nextdelta[0] = rdelta;
                s->nextlen = 1; s->nextiter = 0;
            } else {
                for (i = s->nextiter; i < s->nextiter + s->nextlen; i++) {
                    int j = i % buffer_size;
                    double ppeak = fabs(buffer[nextpos[j]]), pdelta;
                    pdelta = (limit / peak - limit / ppeak) / (((buffer_size - nextpos[j] + s->pos) % buffer_size) / channels);
                    if (pdelta < nextdelta[j]) { nextdelta[j] = pdelta; found = 1; break; }
                }
                if (found) {
                    s->nextlen = i - s->nextiter + 1;
                    nextpos[(s->nextiter + s->nextlen) % buffer_size] = s->pos;
                    nextdelta[(s->nextiter + s->nextlen) % buffer_size] = rdelta;
                    nextpos[(s->nextiter + s->nextlen + 1) % buffer_size] = -1;
                    s->nextlen++;
                }
            }
        }
        {
            const double *buf = &buffer[(s->pos + channels) % buffer_size];
            peak = fabs(buf[0]);
            if (s->pos == s->asc_pos && !s->asc_changed) s->asc_pos = -1;
            if (s->auto_release && s->asc_pos == -1 && peak > limit) { s->asc -= peak; s->asc_c--; }
            s->att += s->delta;
            out = buf[0] * s->att;
        }
        if ((s->pos + channels) % buffer_size == nextpos[s->nextiter]) {
            if (s->auto_release) {
                s->delta = get_rdelta(s, release, rate, peak, limit, s->att, 1);
                if (s->nextlen > 1) {
                    int pnextpos = nextpos[(s->nextiter + 1) % buffer_size];
                    double ppeak = fabs(buffer[pnextpos]), pdelta;
                    pdelta = (limit / ppeak - s->att) / (((buffer_size + pnextpos - ((s->pos + channels) % buffer_size)) % buffer_size) / channels);
                    if (pdelta < s->delta) s->delta = pdelta;
                }
            } else {
                s->delta = nextdelta[s->nextiter];
                s->att = limit / peak;
            }
            s->nextlen -= 1;
            nextpos[s->nextiter] = -1;
            s->nextiter = (s->nextiter + 1) % buffer_size;
        }
        if (s->att > 1.) { s->att = 1.; s->delta = 0.; s->nextiter = 0; s->nextlen = 0; nextpos[0] = -1; }
        if (s->att <= 0.) { s->att = 0.0000000000001; s->delta = (1.0 - s->att) / (rate * release); }
        if (s->att != 1. && (1. - s->att) < 0.0000000000001) s->att = 1.;
        if (s->delta != 0. && fabs(s->delta) < 0.00000000000001) s->delta = 0.;
        out = (out < -limit ? -limit : out > limit ? limit : out) * level * level_out;
        s->pos = (s->pos + channels) % buffer_size;
        if (t >= latency) y[t - latency] = out;
    }
    free(s->buffer); free(s->nextdelta); free(s->nextpos);
    return latency;
}
