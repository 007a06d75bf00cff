/*
 * jivetalking-b200 ORACLE -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C, single-threaded, sequential CPU restatement of the DSP the reference
 * (linuxmatters/jivetalking) delegates to FFmpeg 8.1.1 filters through its spec
 * strings (internal/processor/filters.go:968-989, normalise.go:257-264,1231-1334,
 * analyser_bands.go:33, analyser_output.go:18).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may link or call anything in this directory.  The product
 * (libjtdsp.so) never does.
 *
 * PARITY STATUS: the arithmetic lives in a third-party dependency that is NOT in
 * /root/reference (FFmpeg 8.1.1 libavfilter 11.14.101 / libswresample 6.3.101,
 * fetched as libffmpeg.a by third_party/ffmpeg-statigo/lib/fetch.go:21-23).
 *   - libswresample parts (resampler, true-peak oversampler, sample-format
 *     conversion, downmix): PINNED against the real FFmpeg 8.0.1 libswresample
 *     6.1.100 found in this image (oracle/ref_swr.py, tests/golden/swr_*.npz).
 *   - libavfilter parts (ebur128, astats, aspectralstats, biquads, anlmdn, afftdn,
 *     agate, acompressor, deesser, alimiter, adeclick, loudnorm, volume):
 *     "parity unpinned" -- restated from the published algorithms and recollection
 *     of upstream source; the reference's own tests hold only loose sanity ranges
 *     (analyser_test.go:183-205), which tests/test_oracle_reference_ranges.py checks.
 */
#ifndef JT_ORC_H
#define JT_ORC_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---------------- swr (libswresample/resample.c, resample_template.c) ------------- */
/* Whole-stream polyphase resample with swr's defaults (filter_size 32, phase_shift 10,
 * kaiser beta 9, cutoff 0.97, exact_rational).  `flush` != 0 appends swr's end
 * reflection and drains (aresample at EOF); 0 stops where swr_convert would
 * (ebur128's true-peak path never flushes).  Returns output count (<= cap). */
int64_t orc_swr_resample_f64(const double *in, int64_t n, int in_rate, int out_rate,
                             int flush, double *out, int64_t cap);
int64_t orc_swr_resample_f32(const float *in, int64_t n, int in_rate, int out_rate,
                             int flush, float *out, int64_t cap);
/* number of output samples swr has produced after being fed n_in samples (no flush) */
int64_t orc_swr_out_count(int64_t n_in, int in_rate, int out_rate);
/* total with flush */
int64_t orc_swr_out_count_flush(int64_t n_in, int in_rate, int out_rate);
/* filter bank access for tests: returns phase_count, writes filter_length */
int orc_swr_filter_bank(int in_rate, int out_rate, int *filter_length, double *bank, int cap);

/* sample format conversions (libswresample/audioconvert.c) */
void orc_conv_s16_to_f64(const int16_t *in, int64_t n, double *out);
void orc_conv_s16_to_f32(const int16_t *in, int64_t n, float *out);
void orc_conv_f64_to_s16(const double *in, int64_t n, int16_t *out);
void orc_conv_f32_to_s16(const float *in, int64_t n, int16_t *out);
/* stereo->mono rematrix, float path (libswresample/rematrix.c: 1/sqrt2 each, no normalise) */
void orc_downmix_stereo_f32(const float *in_interleaved, int64_t n_frames, float *out);
/* integer path (normalised 0.5/0.5) */
void orc_downmix_stereo_s16(const int16_t *in_interleaved, int64_t n_frames, int16_t *out);


/* ---------------- ebur128 (libavfilter/f_ebur128.c), mono ------------------------- */
/* x: mono f64 at `rate`.  One tick per rate/10 samples (the 100 ms frames the filter
 * forces with metadata=1).  Per-tick outputs (arrays of tick_cap): M, S (LUFS, after the
 * dual-mono correction, NOT yet "%.3f"-quantised), cumulative sample peak and true peak
 * (linear).  I/LRA are the values the filter would export on the last tick. */
typedef struct orc_r128_summary {
    int64_t n_ticks;
    double I, LRA, LRA_low, LRA_high;
    double sample_peak, true_peak;      /* linear, at the last tick */
    double rel_threshold_400;
} orc_r128_summary;
int orc_ebur128(const double *x, int64_t n, int rate, int dualmono, int want_true_peak,
                double *M, double *S, double *sp_cum, double *tp_cum, int64_t tick_cap,
                orc_r128_summary *sum);

/* ---------------- astats (libavfilter/af_astats.c), one channel -------------------- */
enum { ORC_FMT_S16 = 1, ORC_FMT_S32 = 2, ORC_FMT_FLT = 3, ORC_FMT_DBL = 4 };   /* AVSampleFormat values */
typedef struct orc_astats_out {
    double nb_samples;
    double DC_offset, Min_level, Max_level, Min_difference, Max_difference, Mean_difference,
           RMS_difference, Peak_level, RMS_level, RMS_peak, RMS_trough, Crest_factor,
           Flat_factor, Peak_count, Noise_floor, Noise_floor_count, Entropy, Bit_depth,
           Dynamic_range, Zero_crossings, Zero_crossings_rate;
} orc_astats_out;
int orc_astats(const void *x, int fmt, int64_t n, int rate, orc_astats_out *out);

/* ---------------- aspectralstats (libavfilter/af_aspectralstats.c) ------------------ */
/* x: mono f32.  win_size 2048 (or the filter default 2048), hann, overlap 0.5 -> hop
 * win/2.  One row of 13 floats per hop, order: mean variance centroid spread skewness
 * kurtosis entropy flatness crest flux slope decrease rolloff.  Returns n_hops. */
#define ORC_NSPEC 13
int64_t orc_aspectralstats(const float *x, int64_t n, int rate, int win_size,
                           float *rows, int64_t hop_cap);

#ifdef __cplusplus
}
#endif
#endif
