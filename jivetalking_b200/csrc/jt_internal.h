// libjtdsp internals: context, device-signal descriptor, kernel launcher prototypes.
// Product code: nothing here may reference oracle/.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <atomic>
#include <string>
#include <vector>
#include <map>
#include <cmath>
#include <cstring>
#include "../../include/jtdsp.h"

#define JT_NSM_DEFAULT 148

struct JtTimingSlot { std::string name; double ms = 0; int64_t launches = 0; };

struct jt_ctx {
    int device = 0;
    int num_sms = JT_NSM_DEFAULT;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;   // device->host copy of a finished result while later analysis kernels still run
    cudaStream_t side_stream = nullptr;   // small input-only work (the 17 band graphs) that must not queue behind Pass 2's head (highest priority)
    cudaStream_t low_stream = nullptr;    // lowest priority: work whose result is needed late (Pass 1's astats on the adaptive path) fills the
                                          // SMs the main stream leaves free and the gaps in which the main stream waits for the host
    // JT_TRACE=1: host timestamps (and an event on the main stream) at the driver's milestones, printed when the call returns
    struct TracePt { std::string label; double host; cudaEvent_t ev; };
    std::vector<TracePt> trace_pts; int trace = -1;
    bool meter_after_output = false;      // set around Pass 4's enqueue: ebur128's kernels are queued behind the output stage, so the
                                          // result's download starts ~7 ms earlier and hides under the meter and the analysis tail
    bool defer_astats = false;            // set by the adaptive driver around Pass 1's enqueue: a pre-launched astats goes to low_stream
    std::string last_error;
    std::atomic<int> cancel{0};
    int64_t launches = 0;
    bool timing = false;
    std::vector<JtTimingSlot> slots;
    struct Pending { cudaEvent_t a, b; int slot; };
    std::vector<Pending> pending;
    std::vector<void *> allocs;        // freed by jt_release_all at the end of each API call
    // Device arena: slabs from cudaMalloc, first-fit free lists.  Everything runs on ONE stream, so a block may be
    // handed out again as soon as it is released on the host -- later work is ordered behind its last user -- and
    // steady-state steps never enter the driver's allocator (whose pool growth showed up as 60-450 ms stalls).
    struct Slab { char *base = nullptr; size_t size = 0; std::map<size_t, size_t> free_list; };   // offset -> length
    std::vector<Slab> slabs;
    std::map<void *, size_t> live;     // pointer -> bytes
    size_t live_bytes = 0, peak_bytes = 0;
    std::vector<void *> host_allocs;   // pinned staging
    void *pin_in = nullptr; size_t pin_in_bytes = 0;
    void *pin_out = nullptr; size_t pin_out_bytes = 0;
    // immutable device tables (filter banks, twiddles, windows), content-addressed: steady-state calls upload nothing
    std::map<std::string, void *> dev_tables;
    // pinned host arena for small device->host results (tick energies, statistics rows); reset per API call
    std::vector<std::pair<char *, size_t>> pin_blocks; size_t pin_block = 0, pin_used = 0;
    std::vector<cudaEvent_t> event_pool; size_t events_used = 0;
    // jt_prefetch_input: the NEXT call's input on its way to the device while the current call computes (two slots, ping-pong)
    struct Prefetch { void *dev = nullptr; size_t cap = 0; const void *host = nullptr; size_t bytes = 0; cudaEvent_t ev = nullptr; bool valid = false; };
    Prefetch prefetch[2]; int prefetch_next = 0;
    cudaStream_t upload_stream = nullptr;
    // cross-chunk carries of a stream sharded over several contexts / GPUs (jt_set_exchange)
    jt_exchange_fn exchange = nullptr; void *exchange_user = nullptr; int exchange_ranks = 1;
};

struct JtError { int code; std::string msg; };

#define JT_THROW(code, ...) do { char _b[512]; snprintf(_b, sizeof(_b), __VA_ARGS__); throw JtError{code, _b}; } while (0)
#define JT_CUDA(x) do { cudaError_t _e = (x); if (_e != cudaSuccess) \
    JT_THROW(JT_ERR_CUDA, "%s:%d %s: %s", __FILE__, __LINE__, #x, cudaGetErrorString(_e)); } while (0)

// Opt a kernel in to `bytes` of dynamic shared memory.  The attribute belongs to the function, not to a context: it is only
// ever RAISED (process-wide, per device, under a lock), so a launch that needs less is always legal and two worker
// threads launching the same kernel with different sizes cannot lower it under each other (SURVEY 8b threading contract).
void jt_smem_optin(const void *kernel, size_t bytes);
void *jt_dalloc_bytes(jt_ctx *c, size_t bytes);
template <class T> static inline T *jt_dalloc(jt_ctx *c, size_t n) { return (T *)jt_dalloc_bytes(c, (n ? n : 1) * sizeof(T)); }
void jt_release_all(jt_ctx *c);
// device copy of an immutable host table, cached for the life of the context (keyed by tag + content hash)
const void *jt_dev_table(jt_ctx *c, const char *tag, const void *host, size_t bytes);
template <class T> static inline const T *jt_dev_table(jt_ctx *c, const char *tag, const std::vector<T> &v) { return (const T *)jt_dev_table(c, tag, v.data(), v.size() * sizeof(T)); }
// pinned host scratch valid until the end of the current API call
void *jt_pinned_bytes(jt_ctx *c, size_t bytes);
template <class T> static inline T *jt_pinned(jt_ctx *c, size_t n) { return (T *)jt_pinned_bytes(c, (n ? n : 1) * sizeof(T)); }
// Small transfers between device memory and PINNED host memory (jt_pinned) as a copy kernel on the context's stream instead of
// the copy engines: a copy engine is held by one transfer at a time, so the per-tick values / statistics rows of the compute
// stream queued behind a file-sized upload (jt_prefetch_input) or the download of the result for up to 12 ms.  Either pointer may
// be the pinned one (cudaHostAlloc memory is device-accessible under unified addressing).
void jt_copy_small(jt_ctx *c, void *dst, const void *src, size_t bytes);
// Device copy of a small PER-CALL parameter table (values that differ from file to file -- jt_dev_table would keep one device copy per
// distinct content for the life of the context and waits for its upload): an arena block filled from pinned staging by the copy
// kernel, in stream order, released with the pass's other buffers
template <class T> static inline const T *jt_upload_params(jt_ctx *c, const std::vector<T> &v)
{
    T *d = jt_dalloc<T>(c, v.size());
    if (v.empty()) return d;
    T *h = jt_pinned<T>(c, v.size());
    memcpy(h, v.data(), v.size() * sizeof(T));
    jt_copy_small(c, d, h, v.size() * sizeof(T));
    return d;
}
// an event (timing disabled) recorded on the context's stream now; owned by the context, recycled per API call
cudaEvent_t jt_record_event(jt_ctx *c);
void jt_release_since(jt_ctx *c, size_t mark, const void *keep);   // free allocations made after `mark`, except the one holding `keep`
void jt_release_range(jt_ctx *c, size_t from, size_t to);          // free allocations [from, to) of the call's list, keep the rest
void jt_check_cancel(jt_ctx *c);
void jt_trace(jt_ctx *c, const char *label);     // no-op unless JT_TRACE is set
void jt_trace_dump(jt_ctx *c);
extern "C" void jt_vad_assign_astats(const jt_measurements *m, jt_voice_activity *va);   // jt_adapt.cu
void jt_flush_timing(jt_ctx *c);

// RAII launch bookkeeping: counts one launch of `name`, optionally brackets it with events.
struct JtLaunch {
    jt_ctx *c; int slot = -1; cudaEvent_t a = nullptr, b = nullptr;
    JtLaunch(jt_ctx *ctx, const char *name, int n_launches = 1);
    ~JtLaunch();
};

// Host-side wall time of a section (only when kernel timing is enabled): shows up as a
// "host:<name>" slot next to the kernel groups so bench.py can account for the non-kernel time.
struct JtHost {
    jt_ctx *c; int slot = -1; double t0 = 0;
    JtHost(jt_ctx *ctx, const char *name);
    ~JtHost();
};

// ---- device-resident mono signal --------------------------------------------------------
struct Sig {
    int fmt = 0;        // JT_FMT_S16 / S32 / FLT / DBL
    int rate = 0;
    int64_t n = 0;      // samples (mono)
    void *d = nullptr;  // device pointer
};
static inline size_t jt_fmt_bytes(int fmt) { return fmt == JT_FMT_S16 ? 2 : fmt == JT_FMT_DBL ? 8 : 4; }
static inline bool jt_valid_fmt(int fmt) { return fmt == JT_FMT_S16 || fmt == JT_FMT_S32 || fmt == JT_FMT_FLT || fmt == JT_FMT_DBL; }

// ---- jt_util.cu ---------------------------------------------------------------------------
Sig  jt_convert(jt_ctx *c, const Sig &in, int out_fmt);                   // audioconvert.c semantics
Sig  jt_downmix(jt_ctx *c, const void *d_in, int64_t n_frames, int channels, int fmt, int rate);
void jt_raw_frame_stats(jt_ctx *c, const void *d_in, int64_t n_frames, int channels, int fmt, int frame_size,
                        double *d_sumsq, double *d_peak, int64_t n_src_frames);   // a2, per decoder frame
Sig  jt_volume(jt_ctx *c, const Sig &in, double volume_linear);           // af_volume.c
Sig  jt_gain_f64(jt_ctx *c, const Sig &in, double gain);                  // loudnorm linear mode
Sig  jt_slice(const Sig &in, int64_t start, int64_t count);               // atrim (view, no copy)
Sig  jt_pad_zero(jt_ctx *c, const Sig &in, int64_t n_total);              // asetnsamples pad=1

// ---- k_swr.cu -----------------------------------------------------------------------------
struct SwrPlan {
    int in_rate = 0, out_rate = 0, phase_count = 0, filter_length = 0, div = 0;
    // index advance per output in 1/phase_count input samples = inc_num / inc_den (swr's dst_incr / src_incr, reduced).  Exact
    // ratios (reduced phase count <= 1024): inc_den == 1, inc_num == div.  Otherwise 1024 phases, a fractional advance and
    // swr's linear interpolation between neighbouring phases (linear_interp defaults to on): `linear`, bank has phase_count + 1 rows
    int64_t inc_num = 0, inc_den = 1;
    bool linear = false;
    std::vector<double> bank;   // phase_count (+ 1 when linear) x filter_length
    bool identity = false;
    int64_t out_count(int64_t n_in) const;        // outputs produced after n_in inputs, no flush
    int64_t out_count_flush(int64_t n_in) const;  // total with end reflection
    int64_t first_tap(int64_t m) const;           // input index of output m's first tap
};
SwrPlan jt_swr_plan(int in_rate, int out_rate);
// resample whole stream; work_fmt = JT_FMT_FLT or JT_FMT_DBL (swr's internal format); returns work_fmt signal
// fuse_out_fmt = JT_FMT_S16 lets the f32 path store s16 directly (the result's fmt says what was produced)
Sig  jt_swr_resample(jt_ctx *c, const Sig &in, const SwrPlan &p, int work_fmt, bool flush, int fuse_out_fmt = 0);
// per-tick max |oversampled| (ebur128 true peak): d_tick_tp[k] = max over outputs first available at tick k
void jt_swr_tick_absmax(jt_ctx *c, const Sig &in, const SwrPlan &p, int tick, int64_t n_ticks, double *d_tick_tp);

// ---- k_r128.cu ----------------------------------------------------------------------------
struct R128Result {      // host copies
    std::vector<double> M, S, sp_cum, tp_cum;   // per tick
    double I = -70, LRA = 0, LRA_low = 0, LRA_high = 0;
    int64_t n_ticks = 0;
};
void jt_ebur128(jt_ctx *c, const Sig &in, bool dualmono, bool true_peak, R128Result &out);
// launch / finish halves: the kernels and the device->host copies of the per-tick values are enqueued by
// *_launch; *_finish waits for them (only them) and runs the host part.  Lets the caller keep the GPU busy.
struct R128Pending { int64_t nt = 0; int tick = 0; bool dualmono = false, true_peak = false; double *hp = nullptr, *hk = nullptr, *ht = nullptr; cudaEvent_t ev = nullptr; };
void jt_ebur128_launch(jt_ctx *c, const Sig &in, bool dualmono, bool true_peak, R128Pending &pd);
void jt_ebur128_finish(jt_ctx *c, R128Pending &pd, R128Result &out);
void jt_ebur128_host_finalize(jt_ctx *c, const double *tick_pow, const double *tick_peak, const double *tick_tp_or_null,
                              int64_t n_ticks, int tick, bool dualmono, R128Result &out);
struct LoudnormMeter { double I, LRA, thresh, sample_peak; };
// the state of libavfilter/ebur128.c replayed over per-100 ms K-weighted energies hp[] (see k_r128.cu)
struct LnMeterState {
    int s100; double wgt;
    std::vector<uint32_t> h400, h3000; double sum400 = 0; uint64_t cnt400 = 0;
    LnMeterState(int s100, bool dual_mono);
    void add_tick(const double *hp, int64_t k);                    // tick k (0-based) has completed
    double shortterm_energy(const double *hp, int64_t k) const;    // 3 s ending with tick k (zeros before the stream)
    double shortterm(const double *hp, int64_t k) const;
    double relative_threshold() const, global() const, lra() const;
};
void jt_loudnorm_meter(jt_ctx *c, const Sig &in, bool dual_mono, LoudnormMeter &out);   // libavfilter/ebur128.c
struct LoudnormPending { int64_t nt = 0, nfull = 0; int s100 = 0; bool dual_mono = false; double *hp = nullptr, *hk = nullptr; cudaEvent_t ev = nullptr; };
void jt_loudnorm_meter_launch(jt_ctx *c, const Sig &in, bool dual_mono, LoudnormPending &pd);
void jt_loudnorm_meter_finish(jt_ctx *c, LoudnormPending &pd, LoudnormMeter &out);
// host part over per-100 ms values (K-weighted energy, sample peak): nt ticks, the first nfull of them complete
void jt_loudnorm_meter_host_finalize(const double *hp, const double *hk, int64_t nt, int64_t nfull, int s100, bool dual_mono, LoudnormMeter &out);

// ---- k_loudnorm.cu: af_loudnorm.c dynamic mode ------------------------------------------------
struct jt_loudnorm_opts { double I, TP, LRA, measured_I, measured_TP, measured_LRA, measured_thresh, offset; int dual_mono; };
Sig  jt_loudnorm_dynamic(jt_ctx *c, const Sig &x192, const jt_loudnorm_opts &o, LoudnormPending &pd_in, int *normalization_type);
void jt_loudnorm_tail_launch(jt_ctx *c, const Sig &x_end, int64_t n_stream, bool dual_mono, LoudnormPending &tail, int64_t *first_tick);

// ---- k_astats.cu --------------------------------------------------------------------------
struct AstatsResult { double v[JT_AS_COUNT]; double overall_rms, overall_peak; double nb_samples; };
void jt_astats(jt_ctx *c, const Sig &in, int64_t n_upto, AstatsResult &out);
struct AstatsPending { int64_t n = 0; int fmt = 0, rate = 0, tc = 0; void *host = nullptr; cudaEvent_t ev = nullptr; };
void jt_astats_launch(jt_ctx *c, const Sig &in, int64_t n_upto, AstatsPending &pd);
void jt_astats_finish(jt_ctx *c, AstatsPending &pd, AstatsResult &out);
// one chunk of a longer stream: samples [own0, own0 + own_n) of `in` are stream samples [global_first, ...); the
// partial statistics (pd.host, jt_astats_host_bytes() bytes) merge across chunks and finalise on any host
void jt_astats_chunk_launch(jt_ctx *c, const Sig &in, int64_t own0, int64_t own_n, int64_t global_first, AstatsPending &pd);
size_t jt_astats_host_bytes();
void jt_astats_host_merge(void *dst, const void *src);
void jt_astats_host_finalize(const void *host, int64_t n, int fmt, int tc, AstatsResult &out);

// ---- k_spectral.cu ------------------------------------------------------------------------
// rows: n_hops x JT_SP_COUNT floats on the host
// wanted != nullptr: only those hops (and their predecessors, for flux) are computed; other rows stay 0
void jt_aspectralstats(jt_ctx *c, const Sig &in_flt, int win_size, std::vector<float> &rows, int64_t &n_hops,
                       const std::vector<int64_t> *wanted = nullptr);

struct SpectralPending { int64_t n_hops = 0, n_items = 0; bool sparse = false; std::vector<int64_t> items; float *h_rows = nullptr; cudaEvent_t ev = nullptr; };
void jt_aspectralstats_launch(jt_ctx *c, const Sig &in_flt, int win_size, const std::vector<int64_t> *wanted, SpectralPending &pd);
void jt_aspectralstats_finish(jt_ctx *c, SpectralPending &pd, std::vector<float> &rows, int64_t &n_hops);

// ---- k_biquad.cu --------------------------------------------------------------------------
struct BiquadCoef { double b0, b1, b2, a1, a2; };
BiquadCoef jt_biquad_design(bool highpass, double freq, double q, int rate, bool normalize);
Sig  jt_biquad(jt_ctx *c, const Sig &in, const BiquadCoef &k, bool tdii, double mix);
// 17-band batch (K20): per band highpass(lo) -> lowpass(hi) (direct form I) -> sum of squares / peak
void jt_band_rms_batch(jt_ctx *c, const Sig &in, const double *lo, const double *hi, int n_bands,
                       double *rms_db, int32_t *found);
#define JT_BAND_WARM_MAX (1 << 16)      // samples a band filter lane runs ahead of its segment at most
void jt_band_sumsq(jt_ctx *c, const Sig &in, int64_t acc_from, const double *lo, const double *hi, int n_bands, double *sumsq_host);

// ---- k_flac.cu ------------------------------------------------------------------------------
int64_t jt_flac_bound(int64_t n, int block_size);
void *jt_flac_encode_device(jt_ctx *c, const int16_t *d_pcm, int64_t n, int rate, int block_size, int64_t *n_bytes);
// ---- k_flac_dec.cu: the input side (audio.Reader) ---------------------------------------------
void jt_flac_info_host(const void *bytes, int64_t n_bytes, int *fmt, int *rate, int *channels, int *bps, int64_t *total, int64_t *audio_off);
void *jt_flac_decode_device(jt_ctx *c, const uint8_t *d_bytes, int64_t n_bytes, int *fmt, int *rate, int *channels, int64_t *n_frames);
// packed little-endian 24-bit PCM -> s32 (<< 8), as libavcodec's pcm_s24le decoder
void jt_unpack_s24(jt_ctx *c, const uint8_t *d_bytes, int64_t n_samples, int32_t *d_out);

// ---- k_anlmdn.cu / k_afftdn.cu ------------------------------------------------------------
Sig  jt_anlmdn(jt_ctx *c, const Sig &in_flt, double strength, double patch_s, double research_s, double smooth);
struct AfftdnParams {
    double nr = 12, nf = -50, rf = -38, ad = 0.5, fo = 1.0, bm = 1.25; int nt = 0; int tn = 0; int gs = 0;
    bool has_bn = false; double bn[15] = {0};
};
// afftdn over a window of a longer stream: hops [hop0, hop1) of the window are owned; the tracked noise floor entering
// hop0 is obtained by composing the affine carries of the chunks before this one (key = stream position), exchanged
// through `fn` (an all-gather of fixed-size records; NULL = this chunk starts the stream / single chunk)
struct AfftdnCarry { int64_t hop0 = 0, hop1 = 0, key = 0; jt_exchange_fn fn = nullptr; void *user = nullptr; int n_ranks = 1; };
// The forward transforms of afftdn do not depend on its noise parameters (only on the rate, and on tn / fo for the floor
// candidates): jt_afftdn_forward runs them ahead, while the host still derives those parameters, and jt_afftdn takes them over
// when the stash was made from the very signal it is given.
struct AfftdnFwd { void *d_spec = nullptr; double *d_cand = nullptr; const void *src = nullptr; int64_t n = 0, n_hops = 0; int rate = 0, tn = 0; double fo = 0; };
void jt_afftdn_forward(jt_ctx *c, const Sig &in_flt, const AfftdnParams &p, AfftdnFwd &out);
Sig  jt_afftdn(jt_ctx *c, const Sig &in_flt, const AfftdnParams &p, const AfftdnCarry *carry = nullptr, const AfftdnFwd *fwd = nullptr);

// ---- k_dynamics.cu ------------------------------------------------------------------------
struct GateParams { double threshold, ratio, attack, release, range, knee, makeup; int detection_rms; };
struct CompParams { double threshold, ratio, attack, release, makeup, knee, mix; int detection_rms; };
Sig  jt_agate(jt_ctx *c, const Sig &in_f64, const GateParams &p);
Sig  jt_acompressor(jt_ctx *c, const Sig &in_f64, const CompParams &p);
Sig  jt_deesser(jt_ctx *c, const Sig &in_f64, double intensity, double max_amount, double frequency);

// ---- k_alimiter.cu / k_adeclick.cu --------------------------------------------------------
struct LimiterParams { double limit, attack_ms, release_ms, level_in, level_out; int auto_level, asc, latency; double asc_level; };
Sig  jt_alimiter(jt_ctx *c, const Sig &in_f64, const LimiterParams &p);
Sig  jt_adeclick(jt_ctx *c, const Sig &in_f64, double window_ms, double overlap_pct, double ar_pct,
                 double threshold, double burst_pct, int method_save);

// ---- small helpers ------------------------------------------------------------------------
double jt_wire(const char *fmt, double v);    // value as Go parses it back from FFmpeg printf
static inline int jt_grid_for(int64_t work_items, int block, int num_sms, int max_waves = 64) {
    int64_t g = (work_items + block - 1) / block;
    int64_t cap = (int64_t)num_sms * max_waves;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}
