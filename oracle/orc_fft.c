/* ORACLE (test infrastructure only): radix-2 decimation-in-time FFTs standing in for
 * libavutil/tx (AV_TX_FLOAT_FFT / AV_TX_DOUBLE_FFT / RDFT), which the reference's
 * aspectralstats and afftdn filters call.  Twiddles are computed in double and rounded
 * to the working type; results differ from av_tx at the rounding level only. */
#include "orc_fft.h"
#include <math.h>
#include <stdlib.h>
#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

#define GEN_FFT(NAME, C, T)                                                    \
void NAME(C *x, int n, int inverse)                                            \
{                                                                              \
    for (int i = 1, j = 0; i < n; i++) {                                       \
        int bit = n >> 1;                                                      \
        for (; j & bit; bit >>= 1) j ^= bit;                                   \
        j ^= bit;                                                              \
        if (i < j) { C t = x[i]; x[i] = x[j]; x[j] = t; }                      \
    }                                                                          \
    for (int len = 2; len <= n; len <<= 1) {                                   \
        int half = len >> 1;                                                   \
        for (int k = 0; k < half; k++) {                                       \
            double ang = (inverse ? 2.0 : -2.0) * M_PI * k / len;              \
            T wr = (T)cos(ang), wi = (T)sin(ang);                              \
            for (int i = k; i < n; i += len) {                                 \
                C a = x[i], b = x[i + half];                                   \
                T tr = b.re * wr - b.im * wi, ti = b.re * wi + b.im * wr;      \
                x[i].re = a.re + tr; x[i].im = a.im + ti;                      \
                x[i + half].re = a.re - tr; x[i + half].im = a.im - ti;        \
            }                                                                  \
        }                                                                      \
    }                                                                          \
}
GEN_FFT(orc_fft_f32, orc_cf, float)
GEN_FFT(orc_fft_f64, orc_cd, double)
