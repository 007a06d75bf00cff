/*
 * ORACLE (test infrastructure only) -- restatement of FFmpeg's highpass / lowpass filters
 * (libavfilter/af_biquads.c config_filter(), biquad_<fmt>() and biquad_tdii_<fmt>()) as the
 * reference instantiates them: "highpass=f=80:poles=2:width_type=q:width=0.707:normalize=1:a=tdii",
 * "lowpass=f=20500:..." (internal/processor/filters.go:740-769) and "highpass=f=%f:p=2,
 * lowpass=f=%f:p=2" (analyser_bands.go:33).  Sequential, whole stream, state carried across
 * the stream exactly as libavfilter carries it across frames.  Parity unpinned (see orc.h);
 * coefficients cross-checked against scipy in tests/test_oracle_filters.py.
 */
#include "orc.h"
#include <math.h>
#include <stdint.h>
#include <string.h>

void orc_biquad_design(int highpass, double freq, double q, int rate, int normalize, double *c /* b0 b1 b2 a1 a2 */)
{
    double w0 = 2 * M_PI * freq / rate;
    /* config_filter(): s->bypass when w0 > pi, w0 <= 0 or width <= 0 -- frames pass through untouched */
    if (w0 > M_PI || w0 <= 0.0 || q <= 0.0) { c[0] = 1; c[1] = c[2] = c[3] = c[4] = 0; return; }
    double alpha = sin(w0) / (2 * q);
    double a0 = 1 + alpha, a1 = -2 * cos(w0), a2 = 1 - alpha, b0, b1, b2;
    if (highpass) { b0 = (1 + cos(w0)) / 2; b1 = -(1 + cos(w0)); b2 = (1 + cos(w0)) / 2; }
    else { b0 = (1 - cos(w0)) / 2; b1 = 1 - cos(w0); b2 = (1 - cos(w0)) / 2; }
    a1 /= a0; a2 /= a0; b0 /= a0; b1 /= a0; b2 /= a0; a0 = 1;
    if (normalize && fabs(b0 + b1 + b2) > 1e-6) {
        double factor = (a0 + a1 + a2) / (b0 + b1 + b2);
        b0 *= factor; b1 *= factor; b2 *= factor;
    }
    c[0] = b0; c[1] = b1; c[2] = b2; c[3] = a1; c[4] = a2;
}

#define GEN(NAME, T, F, CLIP, LO, HI)                                                              \
void NAME(const T *in, T *out, int64_t n, const double *c, int tdii, double mix)           \
{                                                                                          \
    F b0 = (F)c[0], b1 = (F)c[1], b2 = (F)c[2], a1 = (F)-c[3], a2 = (F)-c[4];              \
    F wet = (F)mix, dry = (F)(1. - wet);                                                   \
    F w1 = 0, w2 = 0, i1 = 0, i2 = 0, o1 = 0, o2 = 0;                                      \
    for (int64_t i = 0; i < n; i++) {                                                      \
        F x = in[i], o;                                                                    \
        if (tdii) {                                                                        \
            o = b0 * x + w1;                                                               \
            w1 = b1 * x + w2 + a1 * o;                                                     \
            w2 = b2 * x + a2 * o;                                                          \
        } else {                                                                           \
            o = i2 * b2 + i1 * b1 + x * b0 + o2 * a2 + o1 * a1;                            \
            i2 = i1; i1 = x; o2 = o1; o1 = o;                                              \
        }                                                                                  \
        o = o * wet + x * dry;                                                             \
        if (CLIP) { if (o < (F)(LO)) o = (F)(LO); else if (o > (F)(HI)) o = (F)(HI); }     \
        out[i] = (T)o;                                                                     \
    }                                                                                      \
}
GEN(orc_biquad_f32, float, float, 0, 0, 0)
GEN(orc_biquad_f64, double, double, 0, 0, 0)
GEN(orc_biquad_s16, int16_t, float, 1, INT16_MIN, INT16_MAX)
GEN(orc_biquad_s32, int32_t, double, 1, INT32_MIN, INT32_MAX)      /* BIQUAD_FILTER(s32, int32_t, double, INT32_MIN, INT32_MAX, 1) */
