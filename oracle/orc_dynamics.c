/*
 * ORACLE (test infrastructure only) -- restatement of FFmpeg's agate, acompressor, deesser
 * and volume as the reference instantiates them (internal/processor/filters.go:869-932,
 * normalise.go:446-465).  Follows libavfilter/af_agate.c gate()/output_gain(),
 * af_sidechaincompress.c compressor()/output_gain(), hermite.h, af_deesser.c filter_frame(),
 * af_volume.c (precision=float).  f64, sequential, mono.  Parity unpinned (see orc.h).
 */
#include "orc.h"
#include <math.h>

#define FAKE_INFINITY (65536.0 * 65536.0)
#define IS_FAKE_INFINITY(v) (fabs((v) - FAKE_INFINITY) < 1.0)

static double hermite_interpolation(double x, double x0, double x1, double p0, double p1, double m0, double m1)
{
    double width = x1 - x0, t = (x - x0) / width, t2, t3, ct0, ct1, ct2, ct3;
    m0 *= width; m1 *= width;
    t2 = t * t; t3 = t2 * t;
    ct0 = p0; ct1 = m0;
    ct2 = -3 * p0 - 2 * m0 + 3 * p1 - m1;
    ct3 = 2 * p0 + m0 - 2 * p1 + m1;
    return ct3 * t3 + ct2 * t2 + ct1 * t + ct0;
}

void orc_agate(const double *src, double *dst, int64_t n, int rate, double threshold, double ratio, double attack,
               double release, double range, double knee, double makeup, int detection_rms)
{
    double lin_threshold = threshold, lin_knee_sqrt = sqrt(knee);
    if (detection_rms) lin_threshold *= lin_threshold;
    const double attack_coeff = fmin(1., 1. / (attack * rate / 4000.));
    const double release_coeff = fmin(1., 1. / (release * rate / 4000.));
    const double lin_knee_stop = lin_threshold * lin_knee_sqrt, lin_knee_start = lin_threshold / lin_knee_sqrt;
    const double thres = log(lin_threshold), knee_start = log(lin_knee_start), knee_stop = log(lin_knee_stop);
    double lin_slope = 0.0;
    for (int64_t i = 0; i < n; i++) {
        double abs_sample = fabs(src[i]), gain = 1.0;
        if (detection_rms) abs_sample *= abs_sample;
        lin_slope += (abs_sample - lin_slope) * (abs_sample > lin_slope ? attack_coeff : release_coeff);
        if (lin_slope > 0.0 && lin_slope < lin_knee_stop) {
            double slope = log(lin_slope), tratio = ratio, g, delta;
            if (IS_FAKE_INFINITY(ratio)) tratio = 1000.;
            g = (slope - thres) * tratio + thres;
            delta = tratio;
            if (knee > 1. && slope > knee_start)
                g = hermite_interpolation(slope, knee_start, knee_stop, ((knee_start - thres) * tratio + thres), knee_stop, delta, 1.);
            gain = fmax(range, exp(g - slope));
        }
        dst[i] = src[i] * (1.0 * gain * makeup);
    }
}

void orc_acompressor(const double *src, double *dst, int64_t n, int rate, double threshold, double ratio, double attack,
                     double release, double makeup, double knee, double mix, int detection_rms)
{
    const double thres = log(threshold);
    const double lin_knee_start = threshold / sqrt(knee), lin_knee_stop = threshold * sqrt(knee);
    const double adj_knee_start = lin_knee_start * lin_knee_start;
    const double knee_start = log(lin_knee_start), knee_stop = log(lin_knee_stop);
    const double compressed_knee_stop = (knee_stop - thres) / ratio + thres;
    const double attack_coeff = fmin(1., 1. / (attack * rate / 4000.));
    const double release_coeff = fmin(1., 1. / (release * rate / 4000.));
    double lin_slope = 0.0;
    for (int64_t i = 0; i < n; i++) {
        double abs_sample = fabs(src[i]), gain = 1.0;
        if (detection_rms) abs_sample *= abs_sample;
        lin_slope += (abs_sample - lin_slope) * (abs_sample > lin_slope ? attack_coeff : release_coeff);
        const double detector = detection_rms ? adj_knee_start : lin_knee_start;
        if (lin_slope > 0.0 && lin_slope > detector) {
            double slope = log(lin_slope), g, delta;
            if (detection_rms) slope *= 0.5;
            if (IS_FAKE_INFINITY(ratio)) { g = thres; delta = 0.0; }
            else { g = (slope - thres) / ratio + thres; delta = 1.0 / ratio; }
            if (knee > 1.0 && slope < knee_stop)
                g = hermite_interpolation(slope, knee_start, knee_stop, knee_start, compressed_knee_stop, 1.0, delta);
            gain = exp(g - slope);
        }
        dst[i] = src[i] * 1.0 * (gain * makeup * mix + (1. - mix));
    }
}

void orc_deesser(const double *src, double *dst, int64_t n, int rate, double intensity_opt, double max_opt, double freq_opt)
{
    double s1 = 0, s2 = 0, s3 = 0, m1, m2, ratioA = 1.0, ratioB = 1.0, iirSampleA = 0, iirSampleB = 0;
    int flip = 0;
    const double overallscale = rate < 44100 ? 44100.0 / rate : rate / 44100.0;
    const double intensity = pow(intensity_opt, 5) * (8192 / overallscale);
    const double maxdess = 1.0 / pow(10.0, ((max_opt - 1.0) * 48.0) / 20);
    const double iirAmount = pow(freq_opt, 2) / overallscale;
    for (int64_t i = 0; i < n; i++) {
        double sample = src[i], sense, attackspeed, recovery, offset;
        s3 = s2; s2 = s1; s1 = sample;
        m1 = (s1 - s2) * ((s1 - s2) / 1.3);
        m2 = (s2 - s3) * ((s1 - s2) / 1.3);
        sense = (m1 - m2) * ((m1 - m2) / 1.3);
        attackspeed = 7.0 + sense * 1024;
        sense = 1.0 + intensity * intensity * sense;
        sense = fmin(sense, intensity);
        recovery = 1.0 + (0.01 / sense);
        offset = 1.0 - fabs(sample);
        if (flip) {
            iirSampleA = (iirSampleA * (1.0 - (offset * iirAmount))) + (sample * (offset * iirAmount));
            if (ratioA < sense) ratioA = ((ratioA * attackspeed) + sense) / (attackspeed + 1.0);
            else ratioA = 1.0 + ((ratioA - 1.0) / recovery);
            ratioA = fmin(ratioA, maxdess);
            sample = iirSampleA + ((sample - iirSampleA) / ratioA);
        } else {
            iirSampleB = (iirSampleB * (1.0 - (offset * iirAmount))) + (sample * (offset * iirAmount));
            if (ratioB < sense) ratioB = ((ratioB * attackspeed) + sense) / (attackspeed + 1.0);
            else ratioB = 1.0 + ((ratioB - 1.0) / recovery);
            ratioB = fmin(ratioB, maxdess);
            sample = iirSampleB + ((sample - iirSampleB) / ratioB);
        }
        flip = !flip;
        dst[i] = sample;
    }
}

void orc_volume_f32(const float *src, float *dst, int64_t n, double volume)
{
    const float v = (float)volume;
    for (int64_t i = 0; i < n; i++) dst[i] = src[i] * v;
}
