/*
 * libjtdsp -- B200-native (sm_100a CUDA) replacement for the DSP that jivetalking
 * delegates to embedded FFmpeg filter graphs.  C ABI: plain pointers and sizes only.
 *
 * The boundary is the reference's seam S4:
 *     setupFilterGraph(decCtx, filterSpec)          internal/processor/frame_processor.go:164-216
 *     runFilterGraph(ctx, reader, src, sink, cfg)   internal/processor/frame_processor.go:64-159
 *     per-frame metadata dictionary                 internal/processor/analyser_metrics.go:432-483
 *     loudnorm stats_file JSON                      internal/processor/normalise.go:64-75,143-165
 * exposed as a whole-buffer batch call instead of 4096-sample frame pumping.
 *
 * Conventions (mirroring third_party/ffmpeg-statigo/functions.gen.go:5246-5257 WrapErr):
 *   - every entry point returns 0 on success, a negative JT_ERR_* otherwise; a failing
 *     call fails as a whole (never partial output);
 *   - the caller owns every buffer; the library keeps no pointer past the call;
 *   - a jt_ctx is single-threaded (one CUDA stream); distinct contexts may be used
 *     concurrently from distinct threads (one per worker, cf. CloneForWorker,
 *     internal/processor/filters.go:368-373);
 *   - there is no CPU fallback: without a CUDA device jt_create fails with JT_ERR_CUDA.
 */
#ifndef JTDSP_H
#define JTDSP_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define JT_OK                 0
#define JT_ERR_INVALID_ARG   (-1)
#define JT_ERR_CUDA          (-2)
#define JT_ERR_NOMEM         (-3)
#define JT_ERR_SPEC          (-4)   /* filter spec string could not be parsed           */
#define JT_ERR_UNSUPPORTED   (-5)   /* valid FFmpeg, but outside the hot path built here */
#define JT_ERR_CANCELLED     (-6)   /* jt_cancel() seen (ctx.Err(), frame_processor.go:116) */
#define JT_ERR_BUFFER        (-7)   /* caller buffer too small                           */

/* AVSampleFormat values (third_party/ffmpeg-statigo/include/libavutil/samplefmt.h) */
#define JT_FMT_S16 1
#define JT_FMT_S32 2
#define JT_FMT_FLT 3
#define JT_FMT_DBL 4

typedef struct jt_ctx jt_ctx;

int         jt_version(void);
int         jt_create(int device, jt_ctx **out);
void        jt_destroy(jt_ctx *ctx);
const char *jt_strerror(int code);
const char *jt_last_error(const jt_ctx *ctx);      /* detail of the last failure on this ctx */
void        jt_cancel(jt_ctx *ctx);                 /* async-signal-safe flag, polled between launches */

/* ---- the a3 wire (analyser_metrics.go:432-483) as doubles, NaN = key absent ---------- */
enum {  /* lavfi.astats.1.<key>, order of analyser_metrics.go:448-469 */
    JT_AS_Dynamic_range = 0, JT_AS_RMS_level, JT_AS_Peak_level, JT_AS_RMS_trough, JT_AS_RMS_peak,
    JT_AS_DC_offset, JT_AS_Flat_factor, JT_AS_Crest_factor, JT_AS_Zero_crossings_rate,
    JT_AS_Zero_crossings, JT_AS_Max_difference, JT_AS_Min_difference, JT_AS_Mean_difference,
    JT_AS_RMS_difference, JT_AS_Entropy, JT_AS_Min_level, JT_AS_Max_level, JT_AS_Noise_floor,
    JT_AS_Noise_floor_count, JT_AS_Bit_depth, JT_AS_Number_of_samples,
    JT_AS_COUNT
};
enum {  /* lavfi.aspectralstats.1.<key>, order of analyser_metrics.go:433-445 */
    JT_SP_mean = 0, JT_SP_variance, JT_SP_centroid, JT_SP_spread, JT_SP_skewness, JT_SP_kurtosis,
    JT_SP_entropy, JT_SP_flatness, JT_SP_crest, JT_SP_flux, JT_SP_slope, JT_SP_decrease, JT_SP_rolloff,
    JT_SP_COUNT
};

/* One record per frame the reference would pull from the buffersink.  Values are what
 * strconv.ParseFloat would read back from the metadata strings, i.e. already rounded the
 * way FFmpeg prints them ("%.3f" for lavfi.r128.*, "%f" for astats, "%g" for
 * aspectralstats).  astats values are cumulative (reset=0): the reference keeps only the
 * latest (analyser_metrics.go:485-487), so they are materialised on the LAST record only. */
typedef struct jt_frame_meta {
    int64_t first_sample;          /* index of the frame's first sample on the sink link */
    int32_t nb_samples;
    int32_t reserved;
    double  r128_M, r128_S, r128_I, r128_LRA, r128_LRA_low, r128_LRA_high;
    double  r128_true_peak, r128_sample_peak;     /* linear, max over channels            */
    double  astats[JT_AS_COUNT];                   /* lavfi.astats.1.*                      */
    double  astats_overall_RMS_level, astats_overall_Peak_level;  /* lavfi.astats.Overall.* */
    double  spectral[JT_SP_COUNT];                 /* lavfi.aspectralstats.1.*              */
} jt_frame_meta;

/* loudnorm print_format=json fields (LoudnormStats, normalise.go:64-75) as numbers;
 * jt_loudnorm_stats_json() renders the exact text FFmpeg writes to stats_file. */
typedef struct jt_loudnorm_stats {
    double input_i, input_tp, input_lra, input_thresh;
    double output_i, output_tp, output_lra, output_thresh;
    double target_offset;
    int32_t normalization_type;    /* 0 = "linear", 1 = "dynamic" */
    int32_t valid;                 /* 0 when the spec held no loudnorm filter */
} jt_loudnorm_stats;
int jt_loudnorm_stats_json(const jt_loudnorm_stats *st, char *buf, size_t cap);

/* ---- S4: run one filter graph over one whole (decoded) stream --------------------------
 * filter_spec : the exact string the reference builds (BuildFilterSpec filters.go:968-989,
 *               measureWithLoudnorm normalise.go:257-264, buildLoudnormFilterSpec
 *               normalise.go:1231-1334, analyser_bands.go:33, analyser_output.go:18).
 * pcm_in      : interleaved samples, n_frames * channels, sample_fmt = JT_FMT_*.
 * frame_size  : decoder frame size to emulate for metadata cadence (4096 for WAV/FLAC,
 *               processor.go:272-275); <= 0 selects 4096.
 * pcm_out     : receives the sink stream (may be NULL with cap 0 for measure-only graphs);
 *               *out_fmt / *out_rate describe it.
 * meta        : receives up to meta_cap sink-frame records; *n_meta = records the graph
 *               produced (call fails with JT_ERR_BUFFER if meta != NULL and cap is short).
 * ln_stats    : loudnorm JSON fields when the spec contains loudnorm, else valid = 0.
 * The *_dev variant takes DEVICE pointers for pcm_in / pcm_out (already resident in HBM).
 */
int jt_run_graph(jt_ctx *ctx, const char *filter_spec,
                 const void *pcm_in, int64_t n_frames, int sample_rate, int channels, int sample_fmt,
                 int frame_size,
                 void *pcm_out, int64_t pcm_out_cap_frames, int64_t *n_out, int *out_rate, int *out_fmt,
                 jt_frame_meta *meta, int64_t meta_cap, int64_t *n_meta,
                 jt_loudnorm_stats *ln_stats);
int jt_run_graph_dev(jt_ctx *ctx, const char *filter_spec,
                 const void *d_pcm_in, int64_t n_frames, int sample_rate, int channels, int sample_fmt,
                 int frame_size,
                 void *d_pcm_out, int64_t pcm_out_cap_frames, int64_t *n_out, int *out_rate, int *out_fmt,
                 jt_frame_meta *meta, int64_t meta_cap, int64_t *n_meta,
                 jt_loudnorm_stats *ln_stats);
/* upper bound on sink frames / records for sizing caller buffers */
int64_t jt_graph_max_out_frames(const char *filter_spec, int64_t n_frames, int sample_rate);
int64_t jt_graph_max_meta(const char *filter_spec, int64_t n_frames, int sample_rate, int frame_size);

/* ---- accumulated Pass-1 result (metadataAccumulators + []IntervalSample,
 *      analyser_metrics.go:17-32,488-621; collectAnalysisFrames analyser.go:538-650) ----- */
typedef struct jt_interval {
    double timestamp_s;
    double rms_level, peak_level;                 /* dBFS from raw input frames (a2)       */
    double spectral[JT_SP_COUNT];
    int32_t spectral_found, frame_count;
    double momentary_lufs, short_term_lufs, true_peak, sample_peak;   /* dB */
} jt_interval;

typedef struct jt_measurements {
    double input_i, input_tp, input_sp, input_lra, last_m, last_s;    /* LUFS / dB */
    double astats[JT_AS_COUNT];
    double spectral_mean[JT_SP_COUNT];            /* mean over sink frames with keys found */
    int64_t spectral_frames, sink_frames;
    double duration_s;
} jt_measurements;

/* Pass 1 (collectAnalysisFrames): spec = Pass1FilterOrder (filters.go:42-45). */
int jt_analyse(jt_ctx *ctx, const void *pcm_in, int64_t n_frames, int sample_rate, int channels,
               int sample_fmt, int frame_size,
               jt_measurements *out, jt_interval *intervals, int64_t interval_cap, int64_t *n_intervals);

/* ---- one long stream over several GPUs (BASELINE.json configs[3]) -------------------------------
 * Each rank analyses one contiguous chunk of the stream: frames [own_first, own_first + owned), handed over
 * inside a local buffer [local_first, local_first + n_local) that adds context on both sides (one
 * jt_analyse_chunk_unit of left context in mid-stream, >= 4096 frames of right context before the stream's
 * end; own_first, local_first and every chunk length but the last are multiples of the unit).  The result is
 * an opaque, position-independent blob of MERGEABLE values (per-tick K-weighted energy / sample peak / true
 * peak, per-decoder-frame raw sums, the spectral rows its sink frames show, partial astats).  The ranks
 * exchange blobs with one all-gather (ncclAllGather / MPI_Allgather on bytes) and any of them calls
 * jt_analyse_merge (host only) to obtain what jt_analyse returns for the whole stream: windows, gating, LRA
 * and intervals are evaluated on the merged values, so they do not depend on the chunking. */
int64_t jt_analyse_chunk_unit(int sample_rate);
int64_t jt_analyse_chunk_bytes(int64_t owned_frames, int sample_rate);        /* upper bound of a blob */
int jt_analyse_chunk(jt_ctx *ctx, const void *pcm_local, int64_t n_local, int sample_rate, int channels, int sample_fmt,
                     int64_t local_first, int64_t own_first, int64_t owned, int64_t total_frames,
                     void *blob, int64_t blob_cap, int64_t *blob_bytes);
int jt_analyse_merge(int n_chunks, const void *const *blobs,
                     jt_measurements *out, jt_interval *intervals, int64_t interval_cap, int64_t *n_intervals);

/* ---- Passes 2, 3 and 4 of one long stream over several GPUs (configs[3]) ---------------------------
 * jt_graph_chunk runs ANY spec jt_run_graph accepts (BuildFilterSpec filters.go:968-989, measureWithLoudnorm
 * normalise.go:257-264, buildLoudnormFilterSpec normalise.go:1231-1334) on a window [local_first, +n_local) of the
 * stream and returns (a) the OWNED part of the sink audio -- sink samples [*out_first, *out_first + *n_out), which
 * tile the sink stream across chunks, the end-of-stream flush and the asetnsamples padding included in the last one --
 * and (b) a blob of mergeable measurement values of the owned part (per-tick ebur128 energy / peaks, the spectral
 * rows the stream's sink frames show, partial astats, loudnorm's per-100 ms meter values).  jt_graph_merge (host
 * only) turns the blobs of all chunks into the sink-frame records / loudnorm JSON / accumulated measurements
 * jt_run_graph gives for the whole stream: windows, gating, percentiles and the sink-frame cadence are evaluated on
 * the merged values, so they do not depend on the chunking.
 *
 * Boundaries: own_first, local_first and every owned length but the last are multiples of
 * jt_graph_chunk_unit(spec, rate) (whole 100 ms ticks on every measuring link, whole afftdn / adeclick hops, whole
 * resampler periods).  Context: jt_graph_chunk_context() recommends the left / right context (input frames, multiples
 * of the unit) that lets every contractive state of the chain (envelope followers, limiter, biquads, gain smoothing)
 * forget the cut; a mid-stream chunk must bring at least 2 s on the left and 0.25 s on the right.
 *
 * The one state with unbounded memory, afftdn's tracked noise floor (tn=1), crosses chunk boundaries exactly: each
 * chunk reduces its owned hops to an affine carry and the chunks exchange these 32-byte records through the
 * callback set with jt_set_exchange -- an all-gather in rank order (ncclAllGather / MPI_Allgather) that EVERY rank
 * must enter once per call of jt_graph_chunk whose spec makes jt_graph_exchanges() return 1 (ranks without a chunk
 * send `bytes` zero bytes).  This is the only collective inside the data path. */
typedef int (*jt_exchange_fn)(void *user, const void *send, int64_t bytes, void *recv_all /* n_ranks * bytes */);
void    jt_set_exchange(jt_ctx *ctx, jt_exchange_fn fn, void *user, int n_ranks);
int     jt_graph_exchanges(const char *filter_spec);
int64_t jt_graph_chunk_unit(const char *filter_spec, int sample_rate);
int     jt_graph_chunk_context(const char *filter_spec, int sample_rate, int64_t *left_frames, int64_t *right_frames);
int64_t jt_graph_chunk_bytes(const char *filter_spec, int64_t owned_frames, int sample_rate);   /* upper bound of a blob */
int jt_graph_chunk(jt_ctx *ctx, const char *filter_spec,
                   const void *pcm_local, int64_t n_local, int sample_rate, int channels, int sample_fmt,
                   int64_t local_first, int64_t own_first, int64_t owned, int64_t total_frames, int frame_size,
                   void *pcm_out, int64_t pcm_out_cap_frames, int64_t *out_first, int64_t *n_out, int *out_rate, int *out_fmt,
                   void *blob, int64_t blob_cap, int64_t *blob_bytes);
int jt_graph_merge(const char *filter_spec, int64_t total_frames, int sample_rate, int channels, int sample_fmt,
                   int frame_size, int n_chunks, const void *const *blobs,
                   jt_frame_meta *meta, int64_t meta_cap, int64_t *n_meta,
                   jt_loudnorm_stats *ln_stats, struct jt_measurements *accumulated);

/* 17-band region RMS (measureSpeechBandRMS analyser_bands.go:33-104): n_bands (lo,hi) pairs
 * over region [start_s, start_s+duration_s); rms_db[i] = lavfi.astats.Overall.RMS_level,
 * found[i] = 0 when the reference would have seen no metadata. */
int jt_band_rms(jt_ctx *ctx, const void *pcm_in, int64_t n_frames, int sample_rate, int channels,
                int sample_fmt, double start_s, double duration_s,
                const double *lo_hz, const double *hi_hz, int n_bands, double *rms_db, int32_t *found);

/* Output of the full four-pass chain (ProcessAudio processor.go:78-216). */
typedef struct jt_process_result {
    jt_measurements input;            /* Pass 1                                            */
    jt_measurements filtered;         /* Pass 2 output analysis                            */
    jt_measurements final;            /* Pass 4 output analysis                            */
    jt_loudnorm_stats pass3, pass4;   /* loudnorm JSON of Pass 3 and Pass 4                */
    double limiter_ceiling_db, limiter_pregain_db, gain_db, effective_target_i;
    int32_t limiter_needed, limiter_clamped, linear_possible, reserved;
    int64_t n_out;                    /* s16 mono 44.1 kHz samples written                 */
} jt_process_result;

/* pass2_spec: Pass-2 spec string from the reference's AdaptConfig+BuildFilterSpec
 * (adaptive.go:13-40); NULL selects DefaultFilterConfig()'s spec (filters.go:353-355).
 * pcm_out: int16 mono 44.1 kHz, 4096-sample frames (processor.go:379-384). */
int jt_process_audio(jt_ctx *ctx, const void *pcm_in, int64_t n_frames, int sample_rate, int channels,
                     int sample_fmt, const char *pass2_spec,
                     int16_t *pcm_out, int64_t pcm_out_cap, jt_process_result *res);
int jt_process_audio_dev(jt_ctx *ctx, const void *d_pcm_in, int64_t n_frames, int sample_rate, int channels,
                     int sample_fmt, const char *pass2_spec,
                     int16_t *d_pcm_out, int64_t pcm_out_cap, jt_process_result *res);

/* Host-side planners of Pass 3/4 (a9: normalise.go:373-392,407-425,539-561,583-585,611-632,
 * 1198-1203) and the spec builders, exported so callers and tests see the same strings. */
int jt_build_pass3_spec(double output_i, double output_tp, double target_i, double target_tp, double target_lra,
                        char *buf, size_t cap, jt_process_result *plan_out);
int jt_build_pass4_spec(const jt_process_result *plan, const jt_loudnorm_stats *pass3,
                        double target_i, double target_tp, double target_lra, int source_rate,
                        char *buf, size_t cap, double *effective_target_i, double *offset_db);
int jt_default_pass2_spec(char *buf, size_t cap);
int jt_pass1_spec(char *buf, size_t cap);

/* ---- Pass 1 -> Pass 2 host logic: voice-activity detector, region election, AdaptConfig, BuildFilterSpec --------
 *   Pure host functions over the Pass-1 interval records; no device work, no jt_ctx.
 * Durations are Go time.Duration values (int64 nanoseconds) so region arithmetic is exact. */
typedef struct jt_region { int64_t start_ns, end_ns, duration_ns; } jt_region;   /* SpeechRegion / RoomToneRegion analyser.go:26-30,106-112 */

typedef struct jt_region_sample {             /* RegionSample analyser.go:88-104 */
    double rms_level, peak_level, crest_factor;
    double spectral[JT_SP_COUNT];
    double momentary_lufs, short_term_lufs, true_peak, sample_peak;
} jt_region_sample;

typedef struct jt_speech_candidate {          /* SpeechCandidateMetrics analyser.go:114-142 */
    jt_region region;
    jt_region_sample sample;
    double voicing_density, body_band_rms, sib_band_rms, score;
    int64_t original_start_ns, original_duration_ns;
    int32_t bands_measured, was_refined;
} jt_speech_candidate;

#define JT_AFFTDN_BANDS 15                    /* afftdnBandCentresHz analyser_noise_bands.go:15-17 */
typedef struct jt_noise_profile {             /* NoiseProfile analyser.go:49-81 */
    int64_t start_ns, duration_ns;
    double measured_noise_floor, peak_level, crest_factor, entropy;
    double spectral[JT_SP_COUNT];
    double band_noise[JT_AFFTDN_BANDS];
    int32_t bands_measured, n_band_noise;     /* len(BandNoise): 0 until measured, then 15 */
    int32_t warning;                          /* 0 none, 1 short (< 8 s), 2 long (> 18 s): analyser_vad.go:590-594 */
    int32_t reserved;
} jt_noise_profile;

enum { JT_FLOOR_ASTATS = 0, JT_FLOOR_RMS_ESTIMATE, JT_FLOOR_EBUR128_ESTIMATE, JT_FLOOR_VAD_PERCENTILE };

/* What buildInputMeasurements + detectVoiceActivity + assignInputMeasurementSuggestions add to the Pass-1 result
 * (NoiseMetrics / RegionMetrics, analyser.go:188-234; analyser.go:374-415; analyser_vad.go:728-783). */
typedef struct jt_voice_activity {
    double floor, floor_prescan, floor_astats, room_tone_detect_level, floored_fraction, reduction_headroom;
    double split, margin;                     /* clamped Otsu split and hysteresis margin (the VAD log line) */
    double voiced_low_percentile, noise_high_percentile, gate_separation_db;
    int32_t floor_source, voice_activated, gap_tolerance, reserved;
    int32_t has_noise_profile, has_room_tone_sample, has_speech_profile, speech_profile_index /* into candidates */;
    int64_t n_speech_regions, n_candidates;
    jt_region noise_region;                   /* elected (refined) low-cluster region */
    jt_noise_profile noise_profile;
    jt_region_sample room_tone_sample;        /* ElectedRoomToneSample */
    jt_speech_candidate speech_profile;       /* copy of the elected candidate */
} jt_voice_activity;

/* buildInputMeasurements' noise seed (analyser.go:374-415, analyser_noise_seed.go:150-223) followed by
 * detectVoiceActivity (analyser_vad.go:728-783) and assignInputMeasurementSuggestions (analyser.go:515-531).
 * target_i is config.Loudnorm.TargetI (unused by the detector; kept for the reference's signature).
 * speech_regions / candidates may be NULL (then only the counts are returned); JT_ERR_BUFFER when a cap is short. */
int jt_detect_voice_activity(const jt_measurements *m, const jt_interval *intervals, int64_t n_intervals,
                             jt_voice_activity *out,
                             jt_region *speech_regions, int64_t regions_cap,
                             jt_speech_candidate *candidates, int64_t candidates_cap);

/* measureSpeechBands / measureNoiseBands results (analyser_bands.go:106-167, analyser_noise_bands.go:62-119) written
 * onto the elected profiles.  rms[0..1] = body (1-3 kHz) and sibilant (6-9 kHz) band over the speech region,
 * rms[2..16] = the 15 afftdn bands over the room-tone region; found[i] as jt_band_rms reports it.  Either half may
 * be skipped by passing NULL. */
int jt_apply_band_rms(jt_voice_activity *va, const double *speech_rms, const int32_t *speech_found,
                      const double *noise_rms, const int32_t *noise_found);
/* the 17 band edges in that order (speechBandPlan analyser_bands.go:98-103, afftdnBandEdgesHz analyser_noise_bands.go:34-52) */
void jt_band_plan(double lo_hz[17], double hi_hz[17]);

/* ---- EffectiveFilterConfig (filters.go:111-255,349) ---- */
enum { JT_FILTER_DOWNMIX = 1, JT_FILTER_ANALYSIS, JT_FILTER_RESAMPLE, JT_FILTER_RUMBLE_HIGHPASS, JT_FILTER_BANDLIMIT_LOWPASS,
       JT_FILTER_SPEECH_GATE, JT_FILTER_NOISE_REDUCTION, JT_FILTER_LEVELLING_COMPRESSOR, JT_FILTER_DEESSER };
typedef struct jt_biquad_config { int32_t enabled, poles; double frequency, width, mix; char transform[8]; } jt_biquad_config;
typedef struct jt_filter_config {
    int32_t downmix_enabled, analysis_enabled;
    int32_t resample_enabled, resample_rate, resample_frame_size, reserved0;
    char    resample_format[8];
    jt_biquad_config rumble_highpass, bandlimit_lowpass;
    struct { int32_t enabled, afftdn_enabled, afftdn_track_noise, reserved;
             double strength, patch_s, research_s, smooth, afftdn_noise_reduction, afftdn_noise_floor;
             char afftdn_noise_type[8]; char afftdn_band_noise[136]; } noise_reduction;
    struct { int32_t enabled, reserved; double threshold /* linear */, ratio, attack, release, range /* linear */, knee, makeup;
             char detection[8]; } speech_gate;
    struct { int32_t enabled, reserved; double threshold /* dB */, ratio, attack, release, makeup /* dB */, knee, mix; } levelling_compressor;
    struct { int32_t enabled, reserved; double intensity, amount, frequency; } deesser;
    struct { int32_t enabled, reserved; double threshold, window, overlap; char method[8]; } adeclick;
    struct { int32_t enabled, dual_mono, linear, reserved; double target_i, target_tp, target_lra; } loudnorm;
    int32_t n_filter_order;                   /* 0 = Pass2FilterOrder (filters.go:58-68) */
    int32_t filter_order[12];                 /* JT_FILTER_* */
    int32_t reserved1;
} jt_filter_config;

typedef struct jt_adapt_diagnostics {         /* AdaptiveDiagnostics filters.go:275-313 */
    char   bandlimit_lp_reason[40];
    double speech_gate_dynamic_range, speech_gate_quiet_speech_estimate, speech_gate_speech_separation,
           speech_gate_speech_headroom, speech_gate_threshold_unclamped, speech_gate_depth_db;
    char   speech_gate_clamp_reason[16];
    int32_t speech_gate_narrow_gap, afftdn_enabled;
    double afftdn_noise_floor_db;
    char   afftdn_disable_reason[24];
    char   afftdn_noise_type[8];
} jt_adapt_diagnostics;

void jt_default_filter_config(jt_filter_config *cfg);                 /* DefaultFilterConfig filters.go:353-355, 421-532 */
/* AdaptConfig (adaptive.go:13-40): base may be NULL (defaults); diag may be NULL. */
int jt_adapt_config(const jt_filter_config *base, const jt_measurements *m, const jt_voice_activity *va,
                    jt_filter_config *out, jt_adapt_diagnostics *diag);
void jt_sanitize_config(jt_filter_config *cfg);                       /* sanitizeConfig adaptive.go:175-232 */
/* BuildFilterSpec (filters.go:968-989) and the single-filter builders it calls (filters.go:607-960);
 * jt_build_filter(cfg, JT_FILTER_*, ...) renders one of them ("" when the filter is disabled). */
int jt_build_filter_spec(const jt_filter_config *cfg, char *buf, size_t cap);
int jt_build_filter(const jt_filter_config *cfg, int filter_id, char *buf, size_t cap);
int jt_build_adeclick_filter(const jt_filter_config *cfg, char *buf, size_t cap);     /* buildAdeclickFilter filters.go:940-956 */
/* fmt.Sprintf("%g", v) as Go renders it (shortest round-trip digits) -- afftdn nr / nf in the spec (filters.go:806-826) */
int jt_go_format_g(double v, char *buf, size_t cap);

/* The detector's unit-tested stages (analyser_vad.go, analyser_candidates_*.go, analyser_noise_seed.go), exported so the
 * parity tests can replay the reference's own table-driven cases.  axis: 0 = momentary LUFS, 1 = RMS. */
int     jt_vad_detect(const jt_interval *intervals, int64_t n_intervals, double noise_floor_seed, jt_voice_activity *out,
                      jt_region *speech_regions, int64_t regions_cap, jt_speech_candidate *candidates, int64_t candidates_cap);  /* detectVoiceActivity with a given seed */
int64_t jt_vad_intervals_for_duration(int64_t d_ns, int64_t hop_ns);
int     jt_vad_histogram(const jt_interval *iv, int64_t n, int axis, double bin_width,
                         int32_t *bins, int64_t bins_cap, int64_t *n_bins, double *min_level, double *max_level, int64_t *count);
double  jt_vad_otsu_split(const int32_t *bins, int64_t n_bins, double bin_width, double min_level, double max_level);
double  jt_vad_hysteresis_margin(const int32_t *bins, int64_t n_bins, double bin_width, double min_level, double split);
double  jt_vad_percentile_of_sorted(const double *sorted, int64_t n, double pct);
double  jt_vad_percentile_floor(const double *sorted_levels, int64_t n, double seed);
double  jt_vad_clamp_split(double split, double noise_floor, double p75);
double  jt_vad_floored_fraction(const jt_interval *iv, int64_t n, int axis);
int     jt_vad_is_speech_interval(const jt_interval *iv, double split, int axis);
int     jt_vad_gap_tolerance(const uint8_t *flags, int64_t n, int64_t hop_ns);
int64_t jt_vad_build_speech_runs(const jt_interval *iv, int64_t n, double split, double margin, int tol, int axis, int64_t hop_ns,
                                 jt_region *runs, int64_t runs_cap);
int     jt_vad_pick_low_cluster_region(const jt_interval *iv, int64_t n, double split, int axis, int64_t hop_ns, jt_region *out);
int     jt_vad_gate_statistics(const jt_interval *iv, int64_t n, double split, int axis, const jt_region *speech_region /* or NULL */,
                               double *voiced_low, double *noise_high, double *separation);
int     jt_vad_estimate_noise_floor(const jt_interval *iv, int64_t n, double *noise_floor, double *silence_threshold);   /* 1 = ok */
int     jt_vad_noise_profile(const jt_interval *iv, int64_t n, const jt_region *region, jt_noise_profile *out);          /* 1 = profile */
int64_t jt_vad_intervals_in_range(const jt_interval *iv, int64_t n, int64_t start_ns, int64_t end_ns, int64_t *first);
double  jt_vad_score_interval_window(const jt_interval *iv, int64_t n);
double  jt_vad_score_speech_window(const jt_interval *iv, int64_t n);
double  jt_vad_level_variance(const jt_interval *iv, int64_t n, int axis);
double  jt_vad_score_candidate_grounded(const jt_speech_candidate *c, double noise_floor_db, double level_var);
int     jt_vad_measure_candidate(const jt_interval *iv, int64_t n, const jt_region *region, jt_speech_candidate *out);    /* 1 = measured */
int     jt_vad_refine_speech_region(const jt_interval *iv, int64_t n, const jt_region *candidate, jt_region *out);
/* findBestSpeechRegion (analyser_candidates_speech.go:222-326): noise_floor_db = NoiseProfile.MeasuredNoiseFloor, or
 * -INFINITY when there is no profile; returns 1 and *best when a region is elected */
int     jt_vad_find_best_speech_region(const jt_interval *iv, int64_t n, const jt_region *regions, int64_t n_regions,
                                       double noise_floor_db, jt_region *best, jt_speech_candidate *candidates, int64_t cap, int64_t *n_candidates);
double  jt_adapt_gate_threshold(double voiced_low_percentile, double separation, int *narrow_gap);    /* calculateSpeechGateThreshold */
double  jt_adapt_gate_threshold_no_profile(double floor, double room_tone_peak, double room_tone_crest, double ratio, double lufs_gap);
int     jt_adapt_band_noise(const double *bands, int n, char *buf, size_t cap);                      /* buildAfftdnBandNoise */

/* Region re-measure of a pass's output (measureOutputRegionFromReader analyser_output.go:95-233): the graph of
 * analyser_output.go:18 over [start, start + duration) and the Go-side reduction to a RegionSample -- astats Overall RMS / peak
 * (crest = peak - RMS), spectral means over frames that carry them, last M / S, true / sample peak in dB.
 * *frames_processed = sink frames seen; JT_ERR_INVALID_ARG when the region is invalid or holds no frame. */
int jt_measure_output_region(jt_ctx *ctx, const void *pcm, int64_t n_frames, int sample_rate, int channels, int sample_fmt,
                             int64_t start_ns, int64_t duration_ns, jt_region_sample *out, int64_t *frames_processed);
int jt_measure_output_region_dev(jt_ctx *ctx, const void *d_pcm, int64_t n_frames, int sample_rate, int channels, int sample_fmt,
                             int64_t start_ns, int64_t duration_ns, jt_region_sample *out, int64_t *frames_processed);

/* ---- S2 / S1 with the adaptive path on the library side ------------------------------------------------------------
 * jt_analyse_adaptive = AnalyseAudio + AdaptConfig (analyser.go:325-372, processor.go:37-69): Pass 1, the detector, the
 * 17 band RMS graphs over the elected regions (one launch), AdaptConfig and BuildFilterSpec.
 * jt_process_audio_adaptive = ProcessAudio (processor.go:78-216) with that analysis feeding Pass 2. */
typedef struct jt_output_regions {           /* OutputMeasurements.RoomToneSample / SpeechSample analyser.go:318-326 */
    jt_region_sample room_tone, speech;
    int32_t has_room_tone, has_speech;
} jt_output_regions;
typedef struct jt_analysis {
    jt_measurements measurements;
    jt_voice_activity voice_activity;
    jt_filter_config config;
    jt_adapt_diagnostics diagnostics;
    char pass2_spec[2048];
    jt_output_regions filtered_regions, final_regions;    /* jt_process_audio_adaptive only: MeasureOutputRegions on the Pass-2 / Pass-4 output */
} jt_analysis;
int jt_analyse_adaptive(jt_ctx *ctx, const void *pcm_in, int64_t n_frames, int sample_rate, int channels, int sample_fmt,
                        int frame_size, const jt_filter_config *base,
                        jt_analysis *out, jt_interval *intervals, int64_t interval_cap, int64_t *n_intervals);
int jt_process_audio_adaptive(jt_ctx *ctx, const void *pcm_in, int64_t n_frames, int sample_rate, int channels, int sample_fmt,
                        const jt_filter_config *base, int16_t *pcm_out, int64_t pcm_out_cap,
                        jt_process_result *res, jt_analysis *analysis);
int jt_process_audio_adaptive_dev(jt_ctx *ctx, const void *d_pcm_in, int64_t n_frames, int sample_rate, int channels, int sample_fmt,
                        const jt_filter_config *base, int16_t *d_pcm_out, int64_t pcm_out_cap,
                        jt_process_result *res, jt_analysis *analysis);

/* ---- the whole ProcessAudio of ONE stream over several GPUs behind one call per rank (configs[3]) ---------------
 * jt_sharded_plan tells rank `rank` of `n_ranks` which frames of the stream it owns and which window [local_first, +n_local)
 * it must decode (8 s of context on the left, 1 s on the right, all boundaries on `unit`).  Every rank then calls
 * jt_process_audio_sharded with that window and an all-gather callback installed by jt_set_exchange(ctx, fn, user, n_ranks):
 * Pass 1 -> merged measurements -> detector, band graphs, AdaptConfig (identical on every rank) -> Pass 2 (the owned part of
 * the 44.1 kHz output stays on the GPU that made it) -> Pass 3 -> Pass 4, each pass one chunk per rank.  Only measurement
 * blobs, the elected regions' samples and the HALO a neighbour lacks for Pass 3 / 4 cross ranks.  Each rank gets back the
 * part of the result it owns: samples [*out_first, *out_first + *n_out) of the res->n_out samples of the whole output.
 * adaptive = 0 runs DefaultFilterConfig's Pass-2 spec.  loudnorm's dynamic fall-back is not available chunked
 * (JT_ERR_UNSUPPORTED).  The collective must be entered by every rank the same number of times: all ranks must pass the
 * same stream description. */
typedef struct jt_shard_plan { int64_t unit, own_first, owned, local_first, n_local; } jt_shard_plan;
typedef struct jt_shard_timing {          /* host seconds per phase on this rank (device synchronised at each lap) */
    double upload, pass1_chunk, pass1_merge, adapt, pass2_chunk, pass2_merge, regions, halo, pass3_chunk, pass3_merge,
           pass4_chunk, pass4_merge, download, exchange /* inside the callback, all phases */;
    int64_t halo_bytes; int32_t exchange_calls, reserved;
} jt_shard_timing;
int jt_sharded_plan(int64_t total_frames, int sample_rate, int n_ranks, int rank, jt_shard_plan *out);
int jt_process_audio_sharded(jt_ctx *ctx, const void *pcm_local, int64_t n_local, int sample_rate, int channels, int sample_fmt,
                             int64_t total_frames, int n_ranks, int rank, const jt_filter_config *base, int adaptive,
                             int16_t *pcm_out, int64_t pcm_out_cap, int64_t *out_first, int64_t *n_out,
                             jt_process_result *res, jt_analysis *analysis, jt_shard_timing *timing);
int jt_process_audio_sharded_dev(jt_ctx *ctx, const void *d_pcm_local, int64_t n_local, int sample_rate, int channels, int sample_fmt,
                             int64_t total_frames, int n_ranks, int rank, const jt_filter_config *base, int adaptive,
                             int16_t *d_pcm_out, int64_t pcm_out_cap, int64_t *out_first, int64_t *n_out,
                             jt_process_result *res, jt_analysis *analysis, jt_shard_timing *timing);

/* ---- the numbers of the per-file run record ---------------------------------------------------------------------
 * The `schema_version`, `run`, `loudness`, `dynamics`, `spectral` and `noise` blocks of the reference's RunRecord
 * (internal/processor/runrecord.go:24-100; tags analyser.go:140-199, 262-264, analyser_metrics.go:696-710) as the JSON text
 * MarshalRunRecord writes (runrecord.go:425-433): keys sorted within every object, two-space indent, encoding/json float form,
 * non-finite values as null -- so loudness.stages.{input,filtered,final}.{integrated_lufs,true_peak_dbtp,lra_lu} can be
 * compared byte for byte with the reference's `.json`.  res = NULL renders an analysis-only record (input stages only, from
 * analysis->measurements); analysis = NULL drops `noise`; run = NULL drops `run`.  `regions`, `filters`, `normalisation` and
 * `interval_summary` serialise Go structs the caller already holds (jt_analysis / jt_process_result carry their values) and
 * stay with the Go side.  *needed (may be NULL) receives the size the text takes, terminator included; JT_ERR_BUFFER when
 * cap is short.  Host-only. */
typedef struct jt_run_info {                 /* RunProvenance runrecord.go:53-61 */
    const char *input_file, *version, *executable, *processed_at;
    double duration_s; int32_t sample_rate_hz, channels;
} jt_run_info;
int jt_run_record_json(const jt_process_result *res, const jt_analysis *analysis, const jt_run_info *run, double target_i,
                       char *buf, size_t cap, size_t *needed);
int jt_go_json_float(double v, char *buf, size_t cap);        /* encoding/json's rendering of a float64 */

/* RIFF / WAVE input (the reference's fixtures are s16 WAVs, testutil_test.go:140-190; it decodes through libavformat,
 * internal/audio/reader.go): locates the PCM of a file image in memory.  *sample_fmt is a JT_FMT_* value, the samples are
 * interleaved at bytes + *data_offset.  JT_ERR_UNSUPPORTED for 8 / 24 bit or compressed data and for RF64 / BW64 files.  A data
 * chunk of declared size 0 or 0xFFFFFFFF (streamed files) runs to the end of the buffer, as libavformat reads it.
 * Host-only, no jt_ctx. */
int jt_wav_parse(const void *bytes, int64_t n_bytes, int *sample_fmt, int *sample_rate, int *channels,
                 int64_t *data_offset, int64_t *n_frames);

/* ---- FLAC container of the chain's output ------------------------------------------------------------------
 * The reference hands its s16 mono 44.1 kHz result to libavcodec's FLAC encoder in 4096-sample frames
 * (internal/processor/encoder.go:92-101, processor.go:379-384).  jt_flac_encode writes a complete FLAC stream (RFC 9639:
 * "fLaC", STREAMINFO, fixed-block-size frames with CONSTANT / FIXED-predictor / LPC (orders 1..8, 12-bit coefficients, Welch
 * window) + partitioned Rice / VERBATIM subframes, CRC-8 / CRC-16) of n mono s16 samples: bit-exact audio for any FLAC decoder,
 * not the byte sequence libavcodec would emit (its order estimate is replaced by the measured residual of every order; streams
 * come out 0.1 % smaller than its compression_level 5 on the synthetic recipes).  block_size 16..4096 (the reference uses 4096);
 * MD5 is left "unknown" (zero): jt_flac_set_md5. */
int64_t jt_flac_max_bytes(int64_t n_samples, int block_size);
int jt_flac_encode(jt_ctx *ctx, const int16_t *pcm, int64_t n_samples, int sample_rate, int block_size,
                   void *out, int64_t out_cap, int64_t *n_bytes);
int jt_flac_encode_dev(jt_ctx *ctx, const int16_t *d_pcm, int64_t n_samples, int sample_rate, int block_size,
                   void *d_out, int64_t out_cap, int64_t *n_bytes);

/* STREAMINFO's MD5 of the unencoded samples (libavcodec's encoder fills it, encoder.go:92-101).  MD5 is a serial chain -- host
 * work, ~0.5 s per hour of audio on one core -- so it is separate from the encode: jt_flac_encode leaves the field zero ("not
 * known", legal), jt_flac_set_md5 fills it from the samples the stream was made of.  Host-only, no jt_ctx. */
int jt_md5(const void *data, int64_t n_bytes, unsigned char out[16]);
int jt_flac_set_md5(void *stream, int64_t n_bytes, const int16_t *pcm, int64_t n_samples);

/* ---- input-side containers: what the reference's audio.Reader (internal/audio/reader.go:29-188: libavformat demux +
 * libavcodec decode) hands the passes, from a FILE IMAGE in memory ------------------------------------------------------------
 * FLAC: interleaved s16 for streams of <= 16 bits per sample, interleaved s32 with the samples shifted up to fill 32 bits above
 * (libavcodec flacdec.c); 4..24 bits, 1..8 channels, fixed or variable block size, every subframe type of RFC 9639 (CONSTANT,
 * VERBATIM, FIXED, LPC up to order 32, 4 / 5 bit Rice parameters, escape partitions, wasted bits, left/side, side/right, mid/side).
 * Frames are found by a parallel header scan and verified by their CRC-16 before they are decoded; a stream whose frames do
 * not chain from the first one to the end of the buffer (truncated file, trailing tag) is JT_ERR_INVALID_ARG.
 * jt_flac_stream_info is host-only (STREAMINFO: *n_frames is 0 when the encoder did not know the length; *sample_fmt as above).
 * The decode entries write up to cap_frames frames to pcm_out (JT_ERR_BUFFER when the stream has more; *n_frames says how many). */
int jt_flac_stream_info(const void *bytes, int64_t n_bytes, int *sample_fmt, int *sample_rate, int *channels, int *bits_per_sample,
                        int64_t *n_frames, int64_t *audio_offset);
int jt_flac_decode(jt_ctx *ctx, const void *bytes, int64_t n_bytes, void *pcm_out, int64_t cap_frames, int64_t *n_frames,
                   int *sample_fmt, int *sample_rate, int *channels);
int jt_flac_decode_dev(jt_ctx *ctx, const void *d_bytes, int64_t n_bytes, void *d_pcm_out, int64_t cap_frames, int64_t *n_frames,
                   int *sample_fmt, int *sample_rate, int *channels);
/* WAV (RIFF / WAVE, PCM 16 / 24 / 32 bit, IEEE float 32 / 64 bit): the samples as libavcodec's pcm_* decoders give them -- s16, s32
 * (24-bit samples shifted up by 8), flt, dbl -- interleaved.  Same header rules as jt_wav_parse. */
int jt_wav_decode(jt_ctx *ctx, const void *bytes, int64_t n_bytes, void *pcm_out, int64_t cap_frames, int64_t *n_frames,
                  int *sample_fmt, int *sample_rate, int *channels);
int jt_wav_decode_dev(jt_ctx *ctx, const void *d_bytes, int64_t n_bytes, void *d_pcm_out, int64_t cap_frames, int64_t *n_frames,
                  int *sample_fmt, int *sample_rate, int *channels);

/* Double-buffered input for a worker that processes file after file: starts the host -> device copy of a LATER call's input on a
 * separate stream and returns at once.  The next host-buffer entry of this context that is handed the same pointer and the same
 * size (jt_analyse*, jt_run_graph, jt_process_audio*) uses the resident copy instead of uploading.  Issue it right before the call
 * that processes the current file, so the copy runs under that call's kernels: prefetch(file k+1); process(file k); ...
 * The host buffer must stay unchanged until the call that consumes it has returned, and should be pinned (cudaHostAlloc /
 * cudaHostRegister): pageable memory is legal but copies synchronously.  Two copies may be outstanding per context. */
int jt_prefetch_input(jt_ctx *ctx, const void *pcm, int64_t n_frames, int channels, int sample_fmt);

/* Host threads one call may use for its per-frame metadata work (default: up to 8; 0 restores the default).  Process-wide.  A
 * process that shares its node with other workers -- the ranks of jt_process_audio_sharded do this themselves -- should divide the
 * cores among them.  Also settable as JT_HOST_THREADS. */
void jt_set_host_threads(int n);

/* the cudaStream_t (as void *) every kernel and copy of this context is issued on, so a caller can order its own
 * device work against it or bracket calls with CUDA events */
void   *jt_cuda_stream(const jt_ctx *ctx);
/* kernel launches since jt_create / since the last reset (bench.py's gpu_launches) */
int64_t jt_launch_count(const jt_ctx *ctx);
void    jt_reset_launch_count(jt_ctx *ctx);
/* device time (ms) of the dominant kernel group accumulated by CUDA events when
 * jt_enable_kernel_timing(ctx, 1) is set; name of slot i or NULL past the end */
void        jt_enable_kernel_timing(jt_ctx *ctx, int on);
const char *jt_kernel_timing(const jt_ctx *ctx, int slot, double *ms, int64_t *launches);

#ifdef __cplusplus
}
#endif
#endif
