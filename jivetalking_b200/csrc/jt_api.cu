// extern "C" entry points (include/jtdsp.h) + the host-side accumulation the reference does
// in Go around its frame pump (collectAnalysisFrames analyser.go:538-650,
// extractFrameMetadata / extractOutputFrameMetadata analyser_metrics.go:784-924,
// ApplyNormalisation normalise.go:722-905) restated in C++ above the kernels.
#include "jt_graph.h"
#include <thread>
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <climits>
#include <memory>
#include <ctime>

// ---------------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------------
extern "C" int jt_version(void) { return 100; }

extern "C" const char *jt_strerror(int code)
{
    switch (code) {
    case JT_OK: return "success";
    case JT_ERR_INVALID_ARG: return "invalid argument";
    case JT_ERR_CUDA: return "CUDA error (no usable sm_100a device, or a kernel failed)";
    case JT_ERR_NOMEM: return "out of device memory";
    case JT_ERR_SPEC: return "filter spec could not be parsed";
    case JT_ERR_UNSUPPORTED: return "filter or option outside the implemented hot path";
    case JT_ERR_CANCELLED: return "cancelled";
    case JT_ERR_BUFFER: return "caller buffer too small";
    }
    return "unknown error";
}

extern "C" int jt_create(int device, jt_ctx **out)
{
    if (!out) return JT_ERR_INVALID_ARG;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return JT_ERR_CUDA;     // no CPU fallback
    if (device < 0 || device >= ndev) return JT_ERR_INVALID_ARG;
    if (cudaSetDevice(device) != cudaSuccess) return JT_ERR_CUDA;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return JT_ERR_CUDA;
    if (prop.major < 10) return JT_ERR_CUDA;                                           // built for sm_100a only
    jt_ctx *c = new jt_ctx();
    c->device = device; c->num_sms = prop.multiProcessorCount;
    // three priority levels: the side stream (band graphs the host is waiting for) above the main stream above the low stream
    int pr_least = 0, pr_greatest = 0;
    if (cudaDeviceGetStreamPriorityRange(&pr_least, &pr_greatest) != cudaSuccess) { pr_least = pr_greatest = 0; }
    const int pr_main = pr_greatest < pr_least ? std::min(pr_greatest + 1, pr_least) : pr_least;
    if (cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, pr_main) != cudaSuccess) { delete c; return JT_ERR_CUDA; }
    if (cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking) != cudaSuccess) { cudaStreamDestroy(c->stream); delete c; return JT_ERR_CUDA; }
    if (cudaStreamCreateWithPriority(&c->side_stream, cudaStreamNonBlocking, pr_greatest) != cudaSuccess) { cudaStreamDestroy(c->copy_stream); cudaStreamDestroy(c->stream); delete c; return JT_ERR_CUDA; }
    if (pr_main < pr_least && cudaStreamCreateWithPriority(&c->low_stream, cudaStreamNonBlocking, pr_least) != cudaSuccess) c->low_stream = nullptr;
    *out = c;
    return JT_OK;
}

extern "C" void jt_destroy(jt_ctx *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    jt_release_all(c);
    jt_flush_timing(c);
    if (c->pin_in) cudaFreeHost(c->pin_in);
    if (c->pin_out) cudaFreeHost(c->pin_out);
    if (c->upload_stream) { cudaStreamSynchronize(c->upload_stream); cudaStreamDestroy(c->upload_stream); }
    for (auto &p : c->prefetch) { if (p.dev) cudaFree(p.dev); if (p.ev) cudaEventDestroy(p.ev); }
    if (c->stream) cudaStreamSynchronize(c->stream);
    for (auto &kv : c->dev_tables) cudaFree(kv.second);
    for (auto &sl : c->slabs) cudaFree(sl.base);
    for (auto &b : c->pin_blocks) cudaFreeHost(b.first);
    for (cudaEvent_t e : c->event_pool) cudaEventDestroy(e);
    if (c->copy_stream) { cudaStreamSynchronize(c->copy_stream); cudaStreamDestroy(c->copy_stream); }
    if (c->side_stream) { cudaStreamSynchronize(c->side_stream); cudaStreamDestroy(c->side_stream); }
    if (c->low_stream) { cudaStreamSynchronize(c->low_stream); cudaStreamDestroy(c->low_stream); }
    if (c->stream) { cudaStreamSynchronize(c->stream); cudaStreamDestroy(c->stream); }
    delete c;
}

// A long transfer owns its copy engine for its whole duration, and the small copies of the compute stream (per-tick values,
// statistics rows) that land on the same engine queue behind it: a 691 MB upload in one piece delayed the concurrent call by 8 ms
// (scripts/debug/e2e_parts.py).  In pieces of 4 MB (~75 us each) the engine interleaves the other streams' copies between them.
static cudaError_t copy_in_pieces(void *dst, const void *src, size_t bytes, cudaMemcpyKind kind, cudaStream_t stream)
{
    const size_t piece = 4u << 20;
    for (size_t off = 0; off < bytes; off += piece) {
        const cudaError_t e = cudaMemcpyAsync((char *)dst + off, (const char *)src + off, std::min(piece, bytes - off), kind, stream);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

// Double-buffered input: starts the host -> device copy of the input of a LATER call on the context's upload stream and returns
// at once; the next host-buffer entry (jt_analyse*, jt_run_graph, jt_process_audio*) that is handed the same pointer and size finds
// its input resident instead of copying it.  Call it right before processing the current file: the copy then runs under that
// call's kernels (pinned host memory; pageable memory is copied synchronously here, which is legal but hides nothing).
extern "C" int jt_prefetch_input(jt_ctx *c, const void *pcm, int64_t n_frames, int channels, int fmt)
{
    if (!c || !pcm || n_frames <= 0 || channels <= 0 || !jt_valid_fmt(fmt)) return JT_ERR_INVALID_ARG;
    cudaSetDevice(c->device);
    const size_t bytes = (size_t)n_frames * channels * jt_fmt_bytes(fmt);
    if (!c->upload_stream && cudaStreamCreateWithFlags(&c->upload_stream, cudaStreamNonBlocking) != cudaSuccess) { c->last_error = "cudaStreamCreate (upload)"; return JT_ERR_CUDA; }
    jt_ctx::Prefetch &p = c->prefetch[c->prefetch_next];
    c->prefetch_next ^= 1;
    p.valid = false;
    if (p.cap < bytes) {
        if (p.dev) { cudaStreamSynchronize(c->upload_stream); cudaFree(p.dev); p.dev = nullptr; p.cap = 0; }
        if (cudaMalloc(&p.dev, bytes + 256) != cudaSuccess) { cudaGetLastError(); c->last_error = "cudaMalloc (prefetch)"; return JT_ERR_NOMEM; }
        p.cap = bytes;
    }
    if (!p.ev && cudaEventCreateWithFlags(&p.ev, cudaEventDisableTiming) != cudaSuccess) { c->last_error = "cudaEventCreate"; return JT_ERR_CUDA; }
    if (copy_in_pieces(p.dev, pcm, bytes, cudaMemcpyHostToDevice, c->upload_stream) != cudaSuccess || cudaEventRecord(p.ev, c->upload_stream) != cudaSuccess) {
        c->last_error = cudaGetErrorString(cudaGetLastError()); return JT_ERR_CUDA;
    }
    p.host = pcm; p.bytes = bytes; p.valid = true;
    return JT_OK;
}

extern "C" const char *jt_last_error(const jt_ctx *c) { return c ? c->last_error.c_str() : ""; }
extern "C" void jt_cancel(jt_ctx *c) { if (c) c->cancel.store(1); }
extern "C" void *jt_cuda_stream(const jt_ctx *c) { return c ? (void *)c->stream : nullptr; }
extern "C" int64_t jt_launch_count(const jt_ctx *c) { return c ? c->launches : 0; }
extern "C" void jt_reset_launch_count(jt_ctx *c) { if (c) { c->launches = 0; jt_flush_timing(c); for (auto &s : c->slots) { s.ms = 0; s.launches = 0; } } }
extern "C" void jt_enable_kernel_timing(jt_ctx *c, int on) { if (c) c->timing = on != 0; }
extern "C" const char *jt_kernel_timing(const jt_ctx *cc, int slot, double *ms, int64_t *launches)
{
    jt_ctx *c = const_cast<jt_ctx *>(cc);
    if (!c) return nullptr;
    jt_flush_timing(c);
    if (slot < 0 || slot >= (int)c->slots.size()) return nullptr;
    if (ms) *ms = c->slots[slot].ms;
    if (launches) *launches = c->slots[slot].launches;
    return c->slots[slot].name.c_str();
}

// run `body`, translate exceptions into codes, always release per-call device memory
template <class F> static int guarded(jt_ctx *c, F body)
{
    if (!c) return JT_ERR_INVALID_ARG;
    int rc = JT_OK;
    c->cancel.store(0);
    cudaSetDevice(c->device);
    try { body(); }
    catch (const JtError &e) { rc = e.code; c->last_error = e.msg; }
    catch (const std::bad_alloc &) { rc = JT_ERR_NOMEM; c->last_error = "host allocation failed"; }
    if (c->low_stream) cudaStreamSynchronize(c->low_stream);      // (idle unless the call failed half-way)
    if (c->side_stream) cudaStreamSynchronize(c->side_stream);
    jt_release_all(c);
    cudaError_t e = cudaStreamSynchronize(c->stream);
    if (rc == JT_OK && e != cudaSuccess) { rc = JT_ERR_CUDA; c->last_error = cudaGetErrorString(e); }
    jt_trace_dump(c);
    e = cudaStreamSynchronize(c->copy_stream);
    if (rc == JT_OK && e != cudaSuccess) { rc = JT_ERR_CUDA; c->last_error = cudaGetErrorString(e); }
    cudaError_t e2 = cudaGetLastError();
    if (rc == JT_OK && e2 != cudaSuccess) { rc = JT_ERR_CUDA; c->last_error = cudaGetErrorString(e2); }
    return rc;
}

static void *pinned(jt_ctx *c, void **slot, size_t *cap, size_t bytes)
{
    if (*cap < bytes) {
        if (*slot) cudaFreeHost(*slot);
        *slot = nullptr; *cap = 0;
        if (cudaHostAlloc(slot, bytes, cudaHostAllocDefault) != cudaSuccess) JT_THROW(JT_ERR_NOMEM, "cudaHostAlloc(%zu)", bytes);
        *cap = bytes;
    }
    return *slot;
}

// host -> device upload of the caller's PCM.  Pageable caller memory is staged through a
// ctx-owned pinned buffer in chunks so the copy engine runs at full PCIe rate.
static void *upload(jt_ctx *c, const void *h, size_t bytes)
{
    // already on its way (jt_prefetch_input with this very buffer): order the compute stream behind the copy and use it in place
    for (auto &p : c->prefetch)
        if (p.valid && p.host == h && p.bytes == bytes && bytes) {
            p.valid = false;
            JT_CUDA(cudaStreamWaitEvent(c->stream, p.ev, 0));
            return p.dev;
        }
    void *d = jt_dalloc_bytes(c, bytes);
    if (!bytes) return d;
    cudaPointerAttributes at;
    bool is_pinned = cudaPointerGetAttributes(&at, h) == cudaSuccess && at.type == cudaMemoryTypeHost;
    cudaGetLastError();
    if (is_pinned) { JT_CUDA(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, c->stream)); return d; }
    const size_t chunk = 32u << 20;
    char *stage = (char *)pinned(c, &c->pin_in, &c->pin_in_bytes, 2 * chunk);
    cudaEvent_t ev[2]; cudaEventCreate(&ev[0]); cudaEventCreate(&ev[1]);
    size_t off = 0; int k = 0;
    while (off < bytes) {
        const size_t m = std::min(chunk, bytes - off);
        cudaEventSynchronize(ev[k]);
        memcpy(stage + k * chunk, (const char *)h + off, m);
        cudaMemcpyAsync((char *)d + off, stage + k * chunk, m, cudaMemcpyHostToDevice, c->stream);
        cudaEventRecord(ev[k], c->stream);
        off += m; k ^= 1;
    }
    cudaEventSynchronize(ev[0]); cudaEventSynchronize(ev[1]);
    cudaEventDestroy(ev[0]); cudaEventDestroy(ev[1]);
    JT_CUDA(cudaGetLastError());
    return d;
}

static void download(jt_ctx *c, void *h, const void *d, size_t bytes)
{
    if (!bytes) return;
    JT_CUDA(cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, c->stream));
    JT_CUDA(cudaStreamSynchronize(c->stream));
}

extern "C" int jt_loudnorm_stats_json(const jt_loudnorm_stats *s, char *buf, size_t cap)
{
    if (!s || !buf) return JT_ERR_INVALID_ARG;
    int n = snprintf(buf, cap,
        "\n{\n\t\"input_i\" : \"%.2f\",\n\t\"input_tp\" : \"%.2f\",\n\t\"input_lra\" : \"%.2f\",\n\t\"input_thresh\" : \"%.2f\",\n"
        "\t\"output_i\" : \"%.2f\",\n\t\"output_tp\" : \"%+.2f\",\n\t\"output_lra\" : \"%.2f\",\n\t\"output_thresh\" : \"%.2f\",\n"
        "\t\"normalization_type\" : \"%s\",\n\t\"target_offset\" : \"%.2f\"\n}\n",
        s->input_i, s->input_tp, s->input_lra, s->input_thresh, s->output_i, s->output_tp, s->output_lra, s->output_thresh,
        s->normalization_type == 0 ? "linear" : "dynamic", s->target_offset);
    return (n < 0 || (size_t)n >= cap) ? JT_ERR_BUFFER : JT_OK;
}

// ---------------------------------------------------------------------------------------
// S4
// ---------------------------------------------------------------------------------------
static int run_graph_common(jt_ctx *c, const char *spec, const void *pcm_in, bool in_on_device, int64_t n_frames, int rate,
                            int channels, int fmt, int frame_size, void *pcm_out, bool out_on_device, int64_t cap,
                            int64_t *n_out, int *out_rate, int *out_fmt, jt_frame_meta *meta, int64_t meta_cap,
                            int64_t *n_meta, jt_loudnorm_stats *ln)
{
    return guarded(c, [&]() {
        if (!spec || (!pcm_in && n_frames > 0)) JT_THROW(JT_ERR_INVALID_ARG, "null spec or input");
        if (!jt_valid_fmt(fmt)) JT_THROW(JT_ERR_UNSUPPORTED, "input sample format %d", fmt);
        const size_t in_bytes = (size_t)n_frames * channels * jt_fmt_bytes(fmt);
        const void *d_in = in_on_device ? pcm_in : upload(c, pcm_in, in_bytes);
        GraphResult g;
        jt_graph_run(c, spec, d_in, n_frames, rate, channels, fmt, frame_size, pcm_out != nullptr, meta != nullptr || n_meta != nullptr, g);
        if (n_out) *n_out = g.out.n;
        if (out_rate) *out_rate = g.out.rate;
        if (out_fmt) *out_fmt = g.out.fmt;
        if (pcm_out) {
            if (g.out.n > cap) JT_THROW(JT_ERR_BUFFER, "pcm_out holds %lld frames, graph produced %lld", (long long)cap, (long long)g.out.n);
            const size_t ob = (size_t)g.out.n * jt_fmt_bytes(g.out.fmt);
            if (out_on_device) { if (ob) JT_CUDA(cudaMemcpyAsync(pcm_out, g.out.d, ob, cudaMemcpyDeviceToDevice, c->stream)); }
            else download(c, pcm_out, g.out.d, ob);
        }
        if (n_meta) *n_meta = (int64_t)g.meta.size();
        if (meta) {
            if ((int64_t)g.meta.size() > meta_cap) JT_THROW(JT_ERR_BUFFER, "meta holds %lld records, graph produced %zu", (long long)meta_cap, g.meta.size());
            if (!g.meta.empty()) memcpy(meta, g.meta.data(), g.meta.size() * sizeof(jt_frame_meta));
        }
        if (ln) *ln = g.ln;
    });
}

extern "C" int jt_run_graph(jt_ctx *c, const char *spec, const void *pcm_in, int64_t n_frames, int rate, int channels, int fmt,
                            int frame_size, void *pcm_out, int64_t cap, int64_t *n_out, int *out_rate, int *out_fmt,
                            jt_frame_meta *meta, int64_t meta_cap, int64_t *n_meta, jt_loudnorm_stats *ln)
{
    return run_graph_common(c, spec, pcm_in, false, n_frames, rate, channels, fmt, frame_size, pcm_out, false, cap, n_out, out_rate, out_fmt, meta, meta_cap, n_meta, ln);
}
extern "C" int jt_run_graph_dev(jt_ctx *c, const char *spec, const void *d_in, int64_t n_frames, int rate, int channels, int fmt,
                            int frame_size, void *d_out, int64_t cap, int64_t *n_out, int *out_rate, int *out_fmt,
                            jt_frame_meta *meta, int64_t meta_cap, int64_t *n_meta, jt_loudnorm_stats *ln)
{
    return run_graph_common(c, spec, d_in, true, n_frames, rate, channels, fmt, frame_size, d_out, true, cap, n_out, out_rate, out_fmt, meta, meta_cap, n_meta, ln);
}

static int spec_out_rate(const char *spec, int rate)
{
    try {
        for (const FilterNode &f : jt_parse_spec(spec ? spec : "")) {
            if (f.name == "aformat") rate = (int)f.num("sample_rates", "r", rate);
            else if (f.name == "aresample") { if (const std::string *p = f.get("")) rate = atoi(p->c_str()); }
            else if (f.name == "loudnorm" && !jt_loudnorm_linear_mode(f)) rate = 192000;
        }
    } catch (const JtError &) {}
    return rate;
}
extern "C" int64_t jt_graph_max_out_frames(const char *spec, int64_t n_frames, int rate)
{
    const int orate = spec_out_rate(spec, rate);
    return (int64_t)((double)n_frames * orate / rate) + 2 * 4096 + 1024;
}
extern "C" int64_t jt_graph_max_meta(const char *spec, int64_t n_frames, int rate, int frame_size)
{
    (void)spec;
    if (frame_size <= 0) frame_size = 4096;
    const int64_t by_tick = n_frames / std::max(rate / 10, 1) + 4;
    const int64_t by_frame = n_frames / std::min(frame_size, 512) + 4;     // finest framing any filter of the path emits (afftdn rate/80, anlmdn 2K+1)
    return std::max(by_tick, by_frame);
}

// ---------------------------------------------------------------------------------------
// measurement accumulation (latest-wins / averages), shared by Pass 1, 2 and 4
// ---------------------------------------------------------------------------------------
static double ratio_db(double r) { return r <= 0 ? -120.0 : 20 * log10(r); }

struct MeasAcc {
    jt_measurements m; double spec_sum[JT_SP_COUNT];
    MeasAcc() { memset(&m, 0, sizeof(m)); for (double &s : spec_sum) s = 0; for (double &a : m.astats) a = NAN; }
    void add(const jt_frame_meta &f) {
        bool sf = false;
        for (int k = 0; k < JT_SP_COUNT; k++) if (!std::isnan(f.spectral[k])) sf = true;
        if (sf) { for (int k = 0; k < JT_SP_COUNT; k++) spec_sum[k] += std::isnan(f.spectral[k]) ? 0 : f.spectral[k]; m.spectral_frames++; }
        for (int k = 0; k < JT_AS_COUNT; k++) if (!std::isnan(f.astats[k])) m.astats[k] = f.astats[k];
        if (!std::isnan(f.r128_I)) m.input_i = f.r128_I;
        if (!std::isnan(f.r128_M)) m.last_m = f.r128_M;
        if (!std::isnan(f.r128_S)) m.last_s = f.r128_S;
        if (!std::isnan(f.r128_true_peak)) m.input_tp = ratio_db(f.r128_true_peak);
        if (!std::isnan(f.r128_sample_peak)) m.input_sp = ratio_db(f.r128_sample_peak);
        if (!std::isnan(f.r128_LRA)) m.input_lra = f.r128_LRA;
        m.sink_frames++;
    }
    void finish(double duration_s) {
        for (int k = 0; k < JT_SP_COUNT; k++) m.spectral_mean[k] = m.spectral_frames ? spec_sum[k] / m.spectral_frames : 0.0;
        m.duration_s = duration_s;
    }
};

static const char *PASS1_SPEC =
    "aformat=channel_layouts=mono,astats=metadata=1:measure_perchannel=all,"
    "aspectralstats=win_size=2048:win_func=hann:measure=all,"
    "ebur128=metadata=1:peak=sample+true:dualmono=true:target=-16";
static const char *PASS2_DEFAULT_SPEC =
    "aformat=channel_layouts=mono,"
    "highpass=f=80:poles=2:width_type=q:width=0.707:normalize=1:a=tdii,"
    "lowpass=f=20500:poles=2:width_type=q:width=0.707:normalize=1:a=tdii,"
    "anlmdn=s=0.00001:p=0.0060:r=0.0020:m=3,afftdn=nr=12:nt=w:tn=1,"
    "agate=threshold=0.010000:ratio=2.0:attack=5.00:release=200:range=0.1995:knee=3.0:detection=rms:makeup=1.0,"
    "acompressor=threshold=0.125893:ratio=3.0:attack=10:release=200:makeup=1.00:knee=4.0:detection=rms:mix=1.00,"
    "astats=metadata=1:measure_perchannel=all,aspectralstats=win_size=2048:win_func=hann:measure=all,"
    "ebur128=metadata=1:peak=sample+true:dualmono=true:target=-16,"
    "aformat=sample_rates=44100:channel_layouts=mono:sample_fmts=s16,asetnsamples=n=4096";

static int copy_str(const char *s, char *buf, size_t cap)
{
    if (!buf || strlen(s) + 1 > cap) return JT_ERR_BUFFER;
    strcpy(buf, s); return JT_OK;
}
extern "C" int jt_pass1_spec(char *buf, size_t cap) { return copy_str(PASS1_SPEC, buf, cap); }
extern "C" int jt_default_pass2_spec(char *buf, size_t cap) { return copy_str(PASS2_DEFAULT_SPEC, buf, cap); }

// ---------------------------------------------------------------------------------------
// Pass 1 with the interval accumulation of collectAnalysisFrames (analyser.go:571-638)
// ---------------------------------------------------------------------------------------
struct AnalysePending { GraphRun g; double *h_ss = nullptr, *h_pk = nullptr; int64_t nsrc = 0, n_frames = 0; int rate = 0, channels = 0, F = 4096; cudaEvent_t ev = nullptr; };

// Pass 1, device part: raw per-frame statistics (a2) and the Pass-1 graph; nothing here waits for the GPU
static void analyse_enqueue(jt_ctx *c, const void *d_in, int64_t n_frames, int rate, int channels, int fmt, int F, AnalysePending &ap)
{
    if (F <= 0) F = 4096;
    ap.F = F; ap.n_frames = n_frames; ap.rate = rate; ap.channels = channels;
    const int64_t nsrc = (n_frames + F - 1) / F;
    ap.nsrc = nsrc;
    double *d_ss = jt_dalloc<double>(c, nsrc), *d_pk = jt_dalloc<double>(c, nsrc);
    jt_raw_frame_stats(c, d_in, n_frames, channels, fmt, F, d_ss, d_pk, nsrc);
    ap.h_ss = jt_pinned<double>(c, nsrc); ap.h_pk = jt_pinned<double>(c, nsrc);
    if (nsrc) {
        jt_copy_small(c, ap.h_ss, d_ss, sizeof(double) * nsrc);
        jt_copy_small(c, ap.h_pk, d_pk, sizeof(double) * nsrc);
    }
    ap.ev = jt_record_event(c);
    jt_graph_enqueue(c, PASS1_SPEC, d_in, n_frames, rate, channels, fmt, F, false, true, ap.g);
}

// The Go-side accumulation of Pass 1 (collectAnalysisFrames, analyser.go:571-638) over the sink-frame records `g`
// and the raw per-source-frame statistics (a2): whole-file accumulators + 250 ms IntervalSamples.
static void analyse_accumulate(jt_ctx *c, const GraphResult &g, const double *ss, const double *pk, int64_t nsrc, int64_t n_frames,
                               int rate, int channels, int F, jt_measurements *out, jt_interval *iv, int64_t iv_cap, int64_t *n_iv)
{
    MeasAcc acc;
    JtHost hacc(c, "interval_accumulation");
    struct IvAcc { int frameCount = 0; double rawSS = 0; int64_t rawN = 0; double rawPeak = 0; double spec[JT_SP_COUNT] = {0}; bool specFound = false;
                   double mSum = 0, sSum = 0, tpMax = 0, spMax = 0; } ia;
    auto reset = [&]() { ia = IvAcc(); ia.tpMax = -120.0; ia.spMax = -120.0; };
    int64_t n_int = 0;
    auto finalize = [&](int64_t start_ns) {
        if (iv && n_int >= iv_cap) JT_THROW(JT_ERR_BUFFER, "interval buffer holds %lld records", (long long)iv_cap);
        jt_interval r; memset(&r, 0, sizeof(r));
        r.timestamp_s = start_ns * 1e-9;
        r.peak_level = ia.rawPeak > 0 ? 20.0 * log10(ia.rawPeak) : -120.0;
        r.true_peak = ia.tpMax; r.sample_peak = ia.spMax;
        if (ia.rawN > 0) { double rms = sqrt(ia.rawSS / (double)ia.rawN); r.rms_level = rms < 0.00001 ? -120.0 : 20.0 * log10(rms); }
        else r.rms_level = -120.0;
        r.frame_count = ia.frameCount;
        if (ia.frameCount > 0) {
            const double n = ia.frameCount;
            for (int k = 0; k < JT_SP_COUNT; k++) r.spectral[k] = ia.spec[k] / n;
            r.spectral_found = ia.specFound;
            r.momentary_lufs = ia.mSum / n; r.short_term_lufs = ia.sSum / n;
        }
        if (iv) iv[n_int] = r;
        n_int++;
    };
    auto add_sink = [&](const jt_frame_meta &f) {
        acc.add(f);
        const double tp = std::isnan(f.r128_true_peak) ? 0.0 : ratio_db(f.r128_true_peak);
        const double sp = std::isnan(f.r128_sample_peak) ? 0.0 : ratio_db(f.r128_sample_peak);
        if (ia.frameCount == 0 || tp > ia.tpMax) ia.tpMax = tp;
        if (ia.frameCount == 0 || sp > ia.spMax) ia.spMax = sp;
        for (int k = 0; k < JT_SP_COUNT; k++) if (!std::isnan(f.spectral[k])) { ia.spec[k] += f.spectral[k]; ia.specFound = true; }
        ia.mSum += std::isnan(f.r128_M) ? 0.0 : f.r128_M;
        ia.sSum += std::isnan(f.r128_S) ? 0.0 : f.r128_S;
        ia.frameCount++;
    };
    const int64_t hop_ns = 250000000;
    int64_t start_ns = 0, pushed = 0; size_t sink = 0;
    for (int64_t f = 0; f < nsrc; f++) {
        const int64_t t_ns = (int64_t)((double)pushed / (double)rate * 1e9);
        const int64_t nb = std::min<int64_t>(F, n_frames - pushed);
        pushed += nb;
        ia.rawSS += ss[f]; ia.rawN += nb * channels; if (pk[f] > ia.rawPeak) ia.rawPeak = pk[f];
        if (t_ns - start_ns >= hop_ns) { finalize(start_ns); start_ns = t_ns; reset(); }
        while (sink < g.meta.size() && g.meta_ready[sink] <= pushed) add_sink(g.meta[sink++]);
    }
    while (sink < g.meta.size()) add_sink(g.meta[sink++]);       // EOF flush
    if (ia.rawN > 0) finalize(start_ns);
    acc.finish((double)n_frames / rate);
    if (out) *out = acc.m;
    if (n_iv) *n_iv = n_int;
}


// Pass 1, host part: metadata assembly, then the accumulation
static void analyse_finish(jt_ctx *c, AnalysePending &ap, jt_measurements *out, jt_interval *iv, int64_t iv_cap, int64_t *n_iv)
{
    GraphResult g;
    jt_graph_finish(c, ap.g, g);
    jt_trace(c, "pass1 records done");
    JT_CUDA(cudaEventSynchronize(ap.ev));
    analyse_accumulate(c, g, ap.h_ss, ap.h_pk, ap.nsrc, ap.n_frames, ap.rate, ap.channels, ap.F, out, iv, iv_cap, n_iv);
}

static void analyse_device(jt_ctx *c, const void *d_in, int64_t n_frames, int rate, int channels, int fmt, int F,
                           jt_measurements *out, jt_interval *iv, int64_t iv_cap, int64_t *n_iv)
{
    AnalysePending ap;
    analyse_enqueue(c, d_in, n_frames, rate, channels, fmt, F, ap);
    analyse_finish(c, ap, out, iv, iv_cap, n_iv);
}

extern "C" int jt_analyse(jt_ctx *c, const void *pcm_in, int64_t n_frames, int rate, int channels, int fmt, int frame_size,
                          jt_measurements *out, jt_interval *iv, int64_t iv_cap, int64_t *n_iv)
{
    return guarded(c, [&]() {
        if (!pcm_in && n_frames > 0) JT_THROW(JT_ERR_INVALID_ARG, "null input");
        if (!jt_valid_fmt(fmt)) JT_THROW(JT_ERR_UNSUPPORTED, "input sample format %d", fmt);
        const void *d_in = upload(c, pcm_in, (size_t)n_frames * channels * jt_fmt_bytes(fmt));
        analyse_device(c, d_in, n_frames, rate, channels, fmt, frame_size, out, iv, iv_cap, n_iv);
    });
}

// ---------------------------------------------------------------------------------------
// One long stream over several GPUs (BASELINE.json configs[3]): every rank analyses a contiguous
// chunk (plus context) with jt_analyse_chunk, the blobs are exchanged with ONE all-gather, and any rank turns
// them into the jt_measurements / intervals a single-GPU jt_analyse gives with jt_analyse_merge.  What is
// exchanged are the mergeable per-tick / per-hop / per-frame values, not finished statistics: 400 ms / 3 s
// windows, gating and LRA percentiles are evaluated on the merged tick list, so chunking does not change them.
// ---------------------------------------------------------------------------------------
#define JT_CHUNK_MAGIC 0x4a54434855 /* "JTCHU" */
struct JtChunkHdr {
    int64_t magic, bytes, total_frames, first, owned;
    int32_t rate, channels, fmt, frame_size, tick, link_fmt, tc, reserved;
    int64_t tick0, n_ticks, src0, n_src, n_rows, astats_bytes, astats_n, astats_upto;
};
// astats is cumulative and a sink frame shows the snapshot of the decoder frame holding its first sample: the whole-file
// values are those of the LAST sink frame, i.e. statistics of samples [0, upto) with upto possibly short of the end
static int64_t pass1_astats_upto(int64_t total, int tick, int F)
{
    if (total <= 0) return 0;
    const int64_t s_last = (total - 1) / tick * tick;
    return std::min<int64_t>((s_last / F + 1) * (int64_t)F, total);
}
static int64_t gcd64(int64_t a, int64_t b) { while (b) { int64_t t = a % b; a = b; b = t; } return a; }

extern "C" int64_t jt_analyse_chunk_unit(int rate)
{   // chunk boundaries: whole 100 ms ticks, whole 4096-sample decoder frames, whole 1024-sample spectral hops
    if (rate <= 0 || rate % 10) return 0;
    const int64_t tick = rate / 10;
    return tick / gcd64(tick, 4096) * 4096;
}
extern "C" int64_t jt_analyse_chunk_bytes(int64_t owned_frames, int rate)
{
    if (rate <= 0 || rate % 10 || owned_frames < 0) return 0;
    const int64_t nt = owned_frames / (rate / 10) + 2, ns = owned_frames / 4096 + 2;
    return (int64_t)sizeof(JtChunkHdr) + nt * (3 * 8 + 8 + JT_SP_COUNT * 4) + ns * 16 + (int64_t)jt_astats_host_bytes() + 64;
}

// d_in: the local window on the DEVICE.  blob: resized to the chunk's mergeable values.  mono_out (optional): the downmixed
// window in the link's format (what the band graphs of the adaptive path read).
static void analyse_chunk_core(jt_ctx *c, const void *d_in, int64_t n_local, int rate, int channels, int fmt,
                               int64_t local_first, int64_t own_first, int64_t owned, int64_t total_frames,
                               std::vector<char> &blob_out, Sig *mono_out)
{
    {
        const int F = 4096;
        const int64_t U = jt_analyse_chunk_unit(rate);
        if (!d_in || U <= 0) JT_THROW(JT_ERR_INVALID_ARG, "null argument or unsupported rate");
        if (!jt_valid_fmt(fmt)) JT_THROW(JT_ERR_UNSUPPORTED, "input sample format %d", fmt);
        if (own_first % U || local_first % U || local_first > own_first || own_first + owned > total_frames || owned <= 0 ||
            local_first + n_local < own_first + owned || local_first + n_local > total_frames)
            JT_THROW(JT_ERR_INVALID_ARG, "chunk [%lld,+%lld) / local [%lld,+%lld): boundaries must be multiples of jt_analyse_chunk_unit = %lld frames",
                     (long long)own_first, (long long)owned, (long long)local_first, (long long)n_local, (long long)U);
        const bool last = own_first + owned == total_frames;
        if (!last && owned % U) JT_THROW(JT_ERR_INVALID_ARG, "only the last chunk may hold a partial unit");
        if (own_first > 0 && own_first - local_first < U) JT_THROW(JT_ERR_INVALID_ARG, "a chunk in mid-stream needs one unit (%lld frames) of left context", (long long)U);
        if (!last && local_first + n_local - (own_first + owned) < 4096) JT_THROW(JT_ERR_INVALID_ARG, "a chunk before the stream's end needs >= 4096 frames of right context");
        const int tick = rate / 10;
        const int64_t off = own_first - local_first;
        // a2: raw per-decoder-frame statistics of the owned frames
        const int64_t nsrc = (owned + F - 1) / F;
        double *d_ss = jt_dalloc<double>(c, nsrc), *d_pk = jt_dalloc<double>(c, nsrc);
        jt_raw_frame_stats(c, (const char *)d_in + (size_t)off * channels * jt_fmt_bytes(fmt), owned, channels, fmt, F, d_ss, d_pk, nsrc);
        double *h_ss = jt_pinned<double>(c, nsrc), *h_pk = jt_pinned<double>(c, nsrc);
        JT_CUDA(cudaMemcpyAsync(h_ss, d_ss, sizeof(double) * nsrc, cudaMemcpyDeviceToHost, c->stream));
        JT_CUDA(cudaMemcpyAsync(h_pk, d_pk, sizeof(double) * nsrc, cudaMemcpyDeviceToHost, c->stream));
        // the Pass-1 filters on the local stream (aformat mono -> astats -> aspectralstats -> ebur128)
        Sig mono = jt_downmix(c, d_in, n_local, channels, fmt, rate);
        // astats sees the link's own format; aspectralstats converts to flt and ebur128 reads THAT link (s16 / flt widen
        // exactly, 32-bit integers are rounded to float first)
        const Sig mono_as = mono;
        if (mono_out) *mono_out = mono_as;
        if (mono.fmt == JT_FMT_S32) mono = jt_convert(c, mono, JT_FMT_FLT);
        const int64_t as_upto = pass1_astats_upto(total_frames, tick, F);
        const int64_t as_n = std::max<int64_t>(0, std::min(own_first + owned, as_upto) - own_first);
        AstatsPending ap; jt_astats_chunk_launch(c, mono_as, off, as_n, own_first, ap);
        // sink frames (100 ms, the last one possibly partial) whose first sample is owned, and the hop each shows
        const int64_t tick0 = own_first / tick;
        const int64_t n_sink = (owned + tick - 1) / tick, n_ticks = last ? (total_frames / tick - tick0) : owned / tick;
        std::vector<int64_t> wanted(n_sink);
        for (int64_t k = 0; k < n_sink; k++) wanted[k] = (own_first + k * tick) / 1024 - local_first / 1024;
        SpectralPending sp; jt_aspectralstats_launch(c, mono, 2048, &wanted, sp);
        R128Pending rp; jt_ebur128_launch(c, mono, true, true, rp);
        // ---- wait, pack ----
        std::vector<float> rows; int64_t n_hops = 0;
        jt_aspectralstats_finish(c, sp, rows, n_hops);
        JT_CUDA(cudaStreamSynchronize(c->stream));
        JtChunkHdr h; memset(&h, 0, sizeof(h));
        h.magic = JT_CHUNK_MAGIC; h.total_frames = total_frames; h.first = own_first; h.owned = owned;
        h.rate = rate; h.channels = channels; h.fmt = fmt; h.frame_size = F; h.tick = tick; h.link_fmt = mono_as.fmt; h.tc = ap.tc;
        h.tick0 = tick0; h.n_ticks = std::max<int64_t>(n_ticks, 0); h.src0 = own_first / F; h.n_src = nsrc; h.n_rows = n_sink;
        h.astats_bytes = (int64_t)jt_astats_host_bytes(); h.astats_n = ap.host ? as_n : 0; h.astats_upto = as_upto;
        if (!h.tc) h.tc = (int)std::fmax(0.05 * rate + .5, 1);
        const int64_t need = (int64_t)sizeof(h) + h.n_ticks * 24 + nsrc * 16 + n_sink * (8 + JT_SP_COUNT * 4) + h.astats_bytes;
        h.bytes = need;
        blob_out.resize((size_t)need);
        char *w = blob_out.data();
        memcpy(w, &h, sizeof(h)); w += sizeof(h);
        const int64_t lt0 = off / tick;                       // local index of the first owned tick
        if (h.n_ticks > 0 && lt0 + h.n_ticks > rp.nt) JT_THROW(JT_ERR_INVALID_ARG, "internal: local tick range");
        memcpy(w, rp.hp + lt0, 8 * h.n_ticks); w += 8 * h.n_ticks;
        memcpy(w, rp.hk + lt0, 8 * h.n_ticks); w += 8 * h.n_ticks;
        memcpy(w, rp.ht + lt0, 8 * h.n_ticks); w += 8 * h.n_ticks;
        memcpy(w, h_ss, 8 * nsrc); w += 8 * nsrc;
        memcpy(w, h_pk, 8 * nsrc); w += 8 * nsrc;
        for (int64_t k = 0; k < n_sink; k++) { const int64_t gh = (own_first + k * tick) / 1024; memcpy(w, &gh, 8); w += 8; }
        for (int64_t k = 0; k < n_sink; k++) {
            const int64_t lh = wanted[k];
            if (lh >= 0 && lh < n_hops) memcpy(w, &rows[(size_t)lh * JT_SP_COUNT], JT_SP_COUNT * 4); else memset(w, 0, JT_SP_COUNT * 4);
            w += JT_SP_COUNT * 4;
        }
        if (ap.host) memcpy(w, ap.host, h.astats_bytes); else memset(w, 0, h.astats_bytes);
    }
}

extern "C" int jt_analyse_chunk(jt_ctx *c, const void *pcm_local, int64_t n_local, int rate, int channels, int fmt,
                                int64_t local_first, int64_t own_first, int64_t owned, int64_t total_frames,
                                void *blob, int64_t blob_cap, int64_t *blob_bytes)
{
    return guarded(c, [&]() {
        if (!pcm_local || !blob) JT_THROW(JT_ERR_INVALID_ARG, "null argument");
        if (!jt_valid_fmt(fmt)) JT_THROW(JT_ERR_UNSUPPORTED, "input sample format %d", fmt);
        const void *d_in = upload(c, pcm_local, (size_t)n_local * channels * jt_fmt_bytes(fmt));
        std::vector<char> b;
        analyse_chunk_core(c, d_in, n_local, rate, channels, fmt, local_first, own_first, owned, total_frames, b, nullptr);
        if ((int64_t)b.size() > blob_cap) JT_THROW(JT_ERR_BUFFER, "chunk blob needs %lld bytes", (long long)b.size());
        memcpy(blob, b.data(), b.size());
        if (blob_bytes) *blob_bytes = (int64_t)b.size();
    });
}


extern "C" int jt_analyse_merge(int n_chunks, const void *const *blobs, jt_measurements *out, jt_interval *iv, int64_t iv_cap, int64_t *n_iv)
{
    if (n_chunks <= 0 || !blobs) return JT_ERR_INVALID_ARG;
    try {
        std::vector<const JtChunkHdr *> hs;
        for (int i = 0; i < n_chunks; i++) {
            const JtChunkHdr *h = (const JtChunkHdr *)blobs[i];
            if (!h || h->magic != JT_CHUNK_MAGIC) return JT_ERR_INVALID_ARG;
            hs.push_back(h);
        }
        std::sort(hs.begin(), hs.end(), [](const JtChunkHdr *a, const JtChunkHdr *b) { return a->first < b->first; });
        const JtChunkHdr &h0 = *hs[0];
        int64_t expect = 0;
        for (const JtChunkHdr *h : hs) {
            if (h->first != expect || h->rate != h0.rate || h->channels != h0.channels || h->total_frames != h0.total_frames || h->fmt != h0.fmt)
                return JT_ERR_INVALID_ARG;                      // chunks must tile the stream
            expect = h->first + h->owned;
        }
        if (expect != h0.total_frames) return JT_ERR_INVALID_ARG;
        const int64_t total = h0.total_frames; const int rate = h0.rate, tick = h0.tick, F = h0.frame_size;
        const int64_t nt = total / tick, nsrc = (total + F - 1) / F, n_hops = (total + 1023) / 1024;
        std::vector<double> hp(nt), hk(nt), ht(nt), ss(nsrc), pk(nsrc);
        std::vector<float> rows((size_t)n_hops * JT_SP_COUNT, 0.f);
        std::vector<char> as(jt_astats_host_bytes());
        bool first_as = true;
        for (const JtChunkHdr *h : hs) {
            const char *r = (const char *)h + sizeof(JtChunkHdr);
            // a truncated or corrupt all-gather payload must not drive a memcpy: every count non-negative and in range,
            // and the sections must add up to the size the header declares
            if (h->n_ticks < 0 || h->n_src < 0 || h->n_rows < 0 || h->tick0 < 0 || h->src0 < 0 || h->astats_bytes != (int64_t)as.size()) return JT_ERR_INVALID_ARG;
            if (h->n_ticks > nt || h->n_src > nsrc || h->n_rows > n_hops + 2) return JT_ERR_INVALID_ARG;
            if (h->bytes != (int64_t)sizeof(JtChunkHdr) + h->n_ticks * 24 + h->n_src * 16 + h->n_rows * (8 + JT_SP_COUNT * 4) + h->astats_bytes) return JT_ERR_INVALID_ARG;
            if (h->tick0 + h->n_ticks > nt || h->src0 + h->n_src > nsrc) return JT_ERR_INVALID_ARG;
            memcpy(&hp[h->tick0], r, 8 * h->n_ticks); r += 8 * h->n_ticks;
            memcpy(&hk[h->tick0], r, 8 * h->n_ticks); r += 8 * h->n_ticks;
            memcpy(&ht[h->tick0], r, 8 * h->n_ticks); r += 8 * h->n_ticks;
            memcpy(&ss[h->src0], r, 8 * h->n_src); r += 8 * h->n_src;
            memcpy(&pk[h->src0], r, 8 * h->n_src); r += 8 * h->n_src;
            const char *hops = r; r += 8 * h->n_rows;
            for (int64_t k = 0; k < h->n_rows; k++) {
                int64_t gh; memcpy(&gh, hops + 8 * k, 8);
                if (gh >= 0 && gh < n_hops) memcpy(&rows[(size_t)gh * JT_SP_COUNT], r + (size_t)k * JT_SP_COUNT * 4, JT_SP_COUNT * 4);
            }
            r += (size_t)h->n_rows * JT_SP_COUNT * 4;
            if ((size_t)h->astats_bytes != as.size()) return JT_ERR_INVALID_ARG;
            if (h->astats_n > 0) { if (first_as) { memcpy(as.data(), r, as.size()); first_as = false; } else jt_astats_host_merge(as.data(), r); }
        }
        R128Result r128;
        jt_ebur128_host_finalize(nullptr, hp.data(), hk.data(), ht.data(), nt, tick, true, r128);
        AstatsResult ar;
        jt_astats_host_finalize(first_as ? nullptr : as.data(), h0.astats_upto, h0.link_fmt, h0.tc, ar);
        GraphResult g;
        jt_pass1_records(nullptr, total, rate, F, r128, rows, n_hops, &ar, g);
        analyse_accumulate(nullptr, g, ss.data(), pk.data(), nsrc, total, rate, h0.channels, F, out, iv, iv_cap, n_iv);
    } catch (const JtError &e) { return e.code; }
    catch (const std::bad_alloc &) { return JT_ERR_NOMEM; }
    return JT_OK;
}

// ---------------------------------------------------------------------------------------
// Passes 2-4 of one long stream over several GPUs: any graph on a window of the stream (jt_graph_chunk), host merge
// of the chunks' measurement values (jt_graph_merge).  See include/jtdsp.h for the contract.
// ---------------------------------------------------------------------------------------
#define JT_GCHUNK_MAGIC 0x4a544743484b /* "JTGCHK" */
struct JtGraphChunkHdr {
    int64_t magic, bytes, total_frames, first, owned;
    int32_t rate, channels, fmt, frame_size;
    int64_t r_tick0, r_nticks;                  // ebur128: energy, sample peak, true peak per 100 ms tick
    int64_t n_rows;                             // aspectralstats: (global hop, 13 floats) per row
    int64_t astats_bytes, astats_n;             // partial astats over astats_n owned samples
    int32_t astats_fmt, astats_tc;
    int64_t li_tick0, li_nticks, lo_tick0, lo_nticks;   // loudnorm input / output meter: energy, peak per 100 ms
    int64_t out_first, out_n;                   // owned sink samples
};

extern "C" void jt_set_exchange(jt_ctx *c, jt_exchange_fn fn, void *user, int n_ranks)
{
    if (!c) return;
    c->exchange = fn; c->exchange_user = user; c->exchange_ranks = n_ranks > 0 ? n_ranks : 1;
}

static int64_t lcm64(int64_t a, int64_t b) { return a / gcd64(a, b) * b; }
// smallest U with U * link_rate / rate a multiple of `grid` (positions on that link land on its grid)
static int64_t unit_for(int rate, int link_rate, int64_t grid)
{
    const int64_t num = (int64_t)rate * grid;
    return num / gcd64(num, link_rate);
}

struct ChunkGeometry { int64_t unit = 0; int exchanges = 0; };
static ChunkGeometry chunk_geometry(const char *spec, int rate)
{
    ChunkGeometry G;
    if (rate <= 0 || rate % 10) JT_THROW(JT_ERR_UNSUPPORTED, "chunked graphs need a rate that is a multiple of 10 (got %d)", rate);
    int64_t U = rate / 10;
    int r = rate;
    for (const FilterNode &f : jt_parse_spec(spec ? spec : "")) {
        int nr = r;
        if (f.name == "aformat") nr = (int)f.num("sample_rates", "r", r);
        else if (f.name == "aresample") { nr = (int)f.num("sample_rate", "", r); if (const std::string *p = f.get("")) nr = atoi(p->c_str()); }
        else if (f.name == "afftdn") { U = lcm64(U, unit_for(rate, r, std::max(r / 80, 1))); if (f.flag("tn", "track_noise", false)) G.exchanges++; }
        else if (f.name == "adeclick") {
            const double w = f.num("w", "window", 55), o = f.num("o", "overlap", 75);
            const int ws = (int)(r * w / 1000.), hop = std::max((int)(ws * (1. - o / 100.)), 1);
            U = lcm64(U, unit_for(rate, r, hop));
        } else if (f.name == "ebur128") { if (r % 10) JT_THROW(JT_ERR_UNSUPPORTED, "ebur128 at %d Hz", r); U = lcm64(U, unit_for(rate, r, r / 10)); }
        else if (f.name == "loudnorm") {
            if (!jt_loudnorm_linear_mode(f)) nr = 192000;
            U = lcm64(U, unit_for(rate, nr, (nr + 5) / 10));
        } else if (f.name == "atrim") JT_THROW(JT_ERR_UNSUPPORTED, "atrim in a chunked graph");
        if (nr != r) { if (nr <= 0) JT_THROW(JT_ERR_SPEC, "bad rate in %s", f.name.c_str()); U = lcm64(U, unit_for(rate, nr, 1)); r = nr; }
        if (U > ((int64_t)1 << 40)) JT_THROW(JT_ERR_UNSUPPORTED, "chunk unit overflow");
    }
    G.unit = U;
    return G;
}

extern "C" int64_t jt_graph_chunk_unit(const char *spec, int rate)
{
    try { return chunk_geometry(spec, rate).unit; } catch (const JtError &) { return 0; }
}
extern "C" int jt_graph_exchanges(const char *spec)
{
    try { return chunk_geometry(spec, 48000).exchanges; } catch (const JtError &) { return 0; }
}
extern "C" int jt_graph_chunk_context(const char *spec, int rate, int64_t *left, int64_t *right)
{
    try {
        const int64_t U = chunk_geometry(spec, rate).unit;
        // left: 37 time constants of the slowest envelope follower (200 ms release) = 7.4 s -> 8 s; right: look-ahead of
        // the windowed filters (afftdn 37.5 ms, adeclick 55 ms, spectral 2048 samples, resampler taps) -> 1 s
        if (left) *left = ((int64_t)8 * rate + U - 1) / U * U;
        if (right) *right = ((int64_t)rate + U - 1) / U * U;
    } catch (const JtError &e) { return e.code; }
    return JT_OK;
}
extern "C" int64_t jt_graph_chunk_bytes(const char *spec, int64_t owned, int rate)
{
    (void)spec;
    if (rate <= 0 || owned < 0) return 0;
    // ticks on a link of at most 192 kHz (same count as on the input link), spectral rows: at most one per input tick
    const int64_t nt = owned / std::max(rate / 10, 1) + 4 + 32;       // + the re-metered flush frame of dynamic loudnorm (2.9 s)
    return (int64_t)sizeof(JtGraphChunkHdr) + nt * (3 * 8 + 2 * 8 + 2 * 8 + 8 + JT_SP_COUNT * 4) + (int64_t)jt_astats_host_bytes() + 256;
}

// d_in: the local window on the DEVICE.  want_pcm: produce the sink audio; its owned part goes to pcm_out (host, `cap`
// frames) and / or to *own_dev (a fresh device buffer of *n_out frames that outlives the chunk's temporaries -- the caller
// releases it).  want_blob: fill blob_out with the chunk's mergeable measurement values.
static void graph_chunk_core(jt_ctx *c, const char *spec, const void *d_in, int64_t n_local, int rate, int channels, int fmt,
                             int64_t local_first, int64_t own_first, int64_t owned, int64_t total, int frame_size,
                             bool want_pcm, void *pcm_out, int64_t cap, Sig *own_dev,
                             int64_t *out_first, int64_t *n_out, int *out_rate, int *out_fmt,
                             bool want_blob, std::vector<char> &blob_out, const GraphResume *head = nullptr)
{
    {
        void *blob = want_blob ? (void *)&blob_out : nullptr;      // non-null = measurement kernels run
        if (!spec || !d_in) JT_THROW(JT_ERR_INVALID_ARG, "null spec or input");
        if (!jt_valid_fmt(fmt)) JT_THROW(JT_ERR_UNSUPPORTED, "input sample format %d", fmt);
        if (frame_size <= 0) frame_size = 4096;
        const ChunkGeometry G = chunk_geometry(spec, rate);
        const int64_t U = G.unit;
        if (owned <= 0 || own_first < 0 || own_first % U || local_first % U || local_first < 0 || local_first > own_first ||
            own_first + owned > total || local_first + n_local < own_first + owned || local_first + n_local > total)
            JT_THROW(JT_ERR_INVALID_ARG, "chunk [%lld,+%lld) / window [%lld,+%lld) of %lld: boundaries must be multiples of jt_graph_chunk_unit = %lld frames",
                     (long long)own_first, (long long)owned, (long long)local_first, (long long)n_local, (long long)total, (long long)U);
        const bool last = own_first + owned == total;
        if (!last && owned % U) JT_THROW(JT_ERR_INVALID_ARG, "only the last chunk may hold a partial unit");
        if (own_first > 0 && own_first - local_first < 2 * (int64_t)rate) JT_THROW(JT_ERR_INVALID_ARG, "a chunk in mid-stream needs >= 2 s of left context");
        if (own_first == 0 && local_first != 0) JT_THROW(JT_ERR_INVALID_ARG, "bad window");
        if (!last && local_first + n_local - (own_first + owned) < rate / 4 && local_first + n_local != total)
            JT_THROW(JT_ERR_INVALID_ARG, "a chunk before the stream's end needs >= 0.25 s of right context");

        // the whole stream's link sizes and sink-frame cadence (no device work)
        GraphRun gd;
        jt_graph_build(c, spec, nullptr, total, rate, channels, fmt, frame_size, want_pcm, true, JT_GRAPH_DRY, nullptr, gd);
        // the audio filters on the local window
        GraphChunk ck; ck.local_first = local_first; ck.own_first = own_first; ck.owned = owned; ck.total = total; ck.rate = rate; ck.last = last;
        ck.exchange = c->exchange; ck.exchange_user = c->exchange_user; ck.n_ranks = c->exchange_ranks;
        GraphRun gl;
        jt_graph_build(c, spec, d_in, n_local, rate, channels, fmt, frame_size, want_pcm, false, JT_GRAPH_CHUNK, &ck, gl, head);

        // positions of the window / the owned range on a link of rate r whose whole-stream length is n_link
        struct Range { int64_t loc0, a, b; };          // link position of the window's first sample; owned = [a, b)
        auto range_on = [&](int r, int64_t n_link) {
            Range R; R.loc0 = ck.link_pos(local_first, r); R.a = ck.link_pos(own_first, r);
            R.b = last ? n_link : ck.link_pos(own_first + owned, r);
            if (R.b < R.a) R.b = R.a;
            return R;
        };
        JtGraphChunkHdr h; memset(&h, 0, sizeof(h));
        h.magic = JT_GCHUNK_MAGIC; h.total_frames = total; h.first = own_first; h.owned = owned;
        h.rate = rate; h.channels = channels; h.fmt = fmt; h.frame_size = frame_size;

        // ---- ebur128: per-tick values of the owned ticks (1 s of context ahead of them) ----
        R128Pending rp; int64_t r_loc_tick0 = 0;
        if (gd.has_r128 && blob) {
            const Sig &sg = gl.r128_sig; const int tick = sg.rate / 10;
            const Range R = range_on(sg.rate, gd.r128_sig.n);
            const int64_t ctx = std::min<int64_t>(10 * (int64_t)tick, (R.a - R.loc0) / tick * tick);
            const int64_t s0 = R.a - ctx - R.loc0;
            if (R.a % tick || s0 < 0 || s0 > sg.n) JT_THROW(JT_ERR_INVALID_ARG, "internal: ebur128 tick grid");
            jt_ebur128_launch(c, jt_slice(sg, s0, sg.n - s0), gd.r128_dual, gd.r128_tp, rp);
            r_loc_tick0 = ctx / tick;
            h.r_tick0 = R.a / tick;
            h.r_nticks = last ? gd.r128_sig.n / tick - h.r_tick0 : (R.b - R.a) / tick;
            if (h.r_nticks < 0) h.r_nticks = 0;
            if (r_loc_tick0 + h.r_nticks > rp.nt) JT_THROW(JT_ERR_INVALID_ARG, "internal: local tick range (%lld + %lld > %lld)", (long long)r_loc_tick0, (long long)h.r_nticks, (long long)rp.nt);
        }
        // ---- aspectralstats: the rows the stream's sink frames show whose hop starts in the owned range ----
        SpectralPending sp; std::vector<int64_t> want_g, want_l; bool have_sp = false;
        if (gd.has_spec && blob) {
            const Sig &sg = gl.spec_sig; const int hs = gd.spec_win / 2;
            const Range R = range_on(sg.rate, gd.spec_sig.n);
            int64_t s_g = R.a / hs * hs - 4 * (int64_t)hs;                              // window start, on the global hop grid
            const int64_t lo_g = (R.loc0 + hs - 1) / hs * hs;
            if (s_g < lo_g) s_g = lo_g;
            for (const FrameRef &fr : gd.frames) if (fr.hop >= 0) {
                const int64_t pos = (int64_t)fr.hop * hs;
                if (pos >= R.a && pos < R.b && (want_g.empty() || want_g.back() != fr.hop)) want_g.push_back(fr.hop);
            }
            for (int64_t gh : want_g) want_l.push_back(gh - s_g / hs);
            const int64_t s0 = s_g - R.loc0;
            if (s0 < 0 || s0 > sg.n) JT_THROW(JT_ERR_INVALID_ARG, "internal: spectral window");
            if (!want_l.empty()) { jt_aspectralstats_launch(c, jt_slice(sg, s0, sg.n - s0), gd.spec_win, &want_l, sp); have_sp = true; }
        }
        // ---- astats: partial statistics of the owned samples up to the last snapshot a sink frame shows ----
        AstatsPending ap;
        if (gd.has_astats && blob && gd.last_astats_frame >= 0) {
            const Sig &sg = gl.astats_sig;
            const Range R = range_on(sg.rate, gd.astats_sig.n);
            const int64_t upto = gd.frames[gd.last_astats_frame].astats_pos;
            const int64_t as_n = std::max<int64_t>(0, std::min(R.b, upto) - R.a);
            if (as_n > 0) jt_astats_chunk_launch(c, sg, R.a - R.loc0, as_n, R.a, ap);
            h.astats_n = ap.host ? as_n : 0; h.astats_fmt = sg.fmt; h.astats_tc = ap.tc ? ap.tc : (int)std::fmax(0.05 * sg.rate + .5, 1);
        }
        h.astats_bytes = (int64_t)jt_astats_host_bytes();
        // ---- loudnorm meters ----
        LoudnormPending li, lo; int64_t li_loc = 0, lo_loc = 0;
        auto meter = [&](const Sig &sg, int64_t n_link, LoudnormPending &pd, int64_t &loc_tick0, int64_t &tick0, int64_t &nticks) {
            const int s100 = (sg.rate + 5) / 10;
            const Range R = range_on(sg.rate, n_link);
            const int64_t ctx = std::min<int64_t>(10 * (int64_t)s100, (R.a - R.loc0) / s100 * s100);
            const int64_t s0 = R.a - ctx - R.loc0;
            if (R.a % s100 || s0 < 0 || s0 > sg.n) JT_THROW(JT_ERR_INVALID_ARG, "internal: loudnorm tick grid");
            jt_loudnorm_meter_launch(c, jt_slice(sg, s0, sg.n - s0), gd.ln_dual, pd);
            loc_tick0 = ctx / s100; tick0 = R.a / s100;
            nticks = last ? (n_link + s100 - 1) / s100 - tick0 : (R.b - R.a) / s100;
            if (nticks < 0) nticks = 0;
            if (loc_tick0 + nticks > pd.nt) JT_THROW(JT_ERR_INVALID_ARG, "internal: local meter tick range");
        };
        LoudnormPending ltail; int64_t ltail_first = 0, li_split = 0;     // dynamic mode: the flush frame is metered twice (last chunk)
        if (gd.has_ln && blob) {
            meter(gl.ln_in_sig, gd.ln_in_sig.n, li, li_loc, h.li_tick0, h.li_nticks);
            if (gd.ln_linear) meter(gl.ln_out_sig, gd.ln_out_sig.n, lo, lo_loc, h.lo_tick0, h.lo_nticks);
            if (last && gd.ln_dynamic && gd.ln_in_extra) {
                const int s100 = (gl.ln_in_sig.rate + 5) / 10;
                jt_loudnorm_tail_launch(c, gl.ln_in_sig, gd.ln_in_sig.n, gd.ln_dual, ltail, &ltail_first);
                li_split = gd.ln_in_sig.n / s100;                     // ticks from here on come from the tail run
                h.li_nticks = ltail_first + ltail.nt - h.li_tick0;
            }
        }
        // ---- the owned part of the sink audio ----
        {
            const Range R = range_on(gd.out.rate, gd.out.n);
            h.out_first = R.a; h.out_n = R.b - R.a;
            if (out_first) *out_first = R.a;
            if (n_out) *n_out = h.out_n;
            if (out_rate) *out_rate = gd.out.rate;
            if (out_fmt) *out_fmt = gd.out.fmt;
            if (want_pcm && (pcm_out || own_dev)) {
                if (pcm_out && h.out_n > cap) JT_THROW(JT_ERR_BUFFER, "pcm_out holds %lld frames, the chunk owns %lld", (long long)cap, (long long)h.out_n);
                if (gl.out.fmt != gd.out.fmt || !gl.out.d) JT_THROW(JT_ERR_INVALID_ARG, "internal: sink format");
                const size_t bs = jt_fmt_bytes(gl.out.fmt);
                const int64_t s0 = R.a - R.loc0, have = std::max<int64_t>(0, std::min(h.out_n, gl.out.n - s0));
                if (s0 < 0 || (!last && have < h.out_n)) JT_THROW(JT_ERR_INVALID_ARG, "internal: sink range (%lld of %lld)", (long long)have, (long long)h.out_n);
                if (pcm_out) {
                    if (have) JT_CUDA(cudaMemcpyAsync(pcm_out, (const char *)gl.out.d + (size_t)s0 * bs, (size_t)have * bs, cudaMemcpyDeviceToHost, c->stream));
                    if (have < h.out_n) memset((char *)pcm_out + (size_t)have * bs, 0, (size_t)(h.out_n - have) * bs);     // asetnsamples padding
                }
                if (own_dev) {
                    // the caller pre-allocated own_dev->d (so it survives the release of this chunk's temporaries)
                    if (!own_dev->d || own_dev->n < h.out_n) JT_THROW(JT_ERR_INVALID_ARG, "internal: resident sink buffer");
                    if (have) JT_CUDA(cudaMemcpyAsync(own_dev->d, (const char *)gl.out.d + (size_t)s0 * bs, (size_t)have * bs, cudaMemcpyDeviceToDevice, c->stream));
                    if (have < h.out_n) JT_CUDA(cudaMemsetAsync((char *)own_dev->d + (size_t)have * bs, 0, (size_t)(h.out_n - have) * bs, c->stream));
                    own_dev->n = h.out_n; own_dev->fmt = gl.out.fmt; own_dev->rate = gd.out.rate;
                }
            }
        }
        // ---- wait, pack ----
        std::vector<float> rows; int64_t n_hops = 0;
        if (have_sp) jt_aspectralstats_finish(c, sp, rows, n_hops);
        JT_CUDA(cudaStreamSynchronize(c->stream));
        if (!blob) { blob_out.clear(); return; }
        h.n_rows = have_sp ? (int64_t)want_g.size() : 0;
        const int64_t need = (int64_t)sizeof(h) + h.r_nticks * 24 + h.n_rows * (8 + JT_SP_COUNT * 4) + h.astats_bytes + (h.li_nticks + h.lo_nticks) * 16;
        h.bytes = need;
        blob_out.resize((size_t)need);
        char *w = blob_out.data();
        memcpy(w, &h, sizeof(h)); w += sizeof(h);
        if (h.r_nticks > 0) {
            memcpy(w, rp.hp + r_loc_tick0, 8 * h.r_nticks); w += 8 * h.r_nticks;
            memcpy(w, rp.hk + r_loc_tick0, 8 * h.r_nticks); w += 8 * h.r_nticks;
            if (gd.r128_tp) memcpy(w, rp.ht + r_loc_tick0, 8 * h.r_nticks); else memset(w, 0, 8 * h.r_nticks);
            w += 8 * h.r_nticks;
        }
        for (int64_t k = 0; k < h.n_rows; k++) { memcpy(w, &want_g[k], 8); w += 8; }
        for (int64_t k = 0; k < h.n_rows; k++) {
            const int64_t lh = want_l[k];
            if (lh >= 0 && lh < n_hops) memcpy(w, &rows[(size_t)lh * JT_SP_COUNT], JT_SP_COUNT * 4); else memset(w, 0, JT_SP_COUNT * 4);
            w += JT_SP_COUNT * 4;
        }
        if (ap.host && h.astats_n > 0) memcpy(w, ap.host, h.astats_bytes); else memset(w, 0, h.astats_bytes);
        w += h.astats_bytes;
        if (h.li_nticks > 0) {
            std::vector<double> tp(h.li_nticks), tk(h.li_nticks);
            for (int64_t k = 0; k < h.li_nticks; k++) {
                const int64_t gt = h.li_tick0 + k;
                if (ltail.nt > 0 && gt >= li_split) { tp[k] = ltail.hp[gt - ltail_first]; tk[k] = ltail.hk[gt - ltail_first]; }
                else { tp[k] = li.hp[li_loc + k]; tk[k] = li.hk[li_loc + k]; }
            }
            memcpy(w, tp.data(), 8 * h.li_nticks); w += 8 * h.li_nticks; memcpy(w, tk.data(), 8 * h.li_nticks); w += 8 * h.li_nticks;
        }
        if (h.lo_nticks > 0) { memcpy(w, lo.hp + lo_loc, 8 * h.lo_nticks); w += 8 * h.lo_nticks; memcpy(w, lo.hk + lo_loc, 8 * h.lo_nticks); w += 8 * h.lo_nticks; }
    }
}

extern "C" int jt_graph_chunk(jt_ctx *c, const char *spec, const void *pcm_local, int64_t n_local, int rate, int channels, int fmt,
                              int64_t local_first, int64_t own_first, int64_t owned, int64_t total, int frame_size,
                              void *pcm_out, int64_t cap, int64_t *out_first, int64_t *n_out, int *out_rate, int *out_fmt,
                              void *blob, int64_t blob_cap, int64_t *blob_bytes)
{
    return guarded(c, [&]() {
        if (!spec || !pcm_local) JT_THROW(JT_ERR_INVALID_ARG, "null spec or input");
        if (!jt_valid_fmt(fmt)) JT_THROW(JT_ERR_UNSUPPORTED, "input sample format %d", fmt);
        const void *d_in = upload(c, pcm_local, (size_t)n_local * channels * jt_fmt_bytes(fmt));
        std::vector<char> b;
        graph_chunk_core(c, spec, d_in, n_local, rate, channels, fmt, local_first, own_first, owned, total, frame_size,
                         pcm_out != nullptr, pcm_out, cap, nullptr, out_first, n_out, out_rate, out_fmt, blob != nullptr, b);
        if (blob) {
            if ((int64_t)b.size() > blob_cap) JT_THROW(JT_ERR_BUFFER, "chunk blob needs %lld bytes", (long long)b.size());
            memcpy(blob, b.data(), b.size());
        }
        if (blob_bytes) *blob_bytes = (int64_t)b.size();
    });
}


extern "C" int jt_graph_merge(const char *spec, int64_t total, int rate, int channels, int fmt, int frame_size,
                              int n_chunks, const void *const *blobs, jt_frame_meta *meta, int64_t meta_cap, int64_t *n_meta,
                              jt_loudnorm_stats *ln, jt_measurements *accumulated)
{
    if (n_chunks <= 0 || !blobs || !spec) return JT_ERR_INVALID_ARG;
    try {
        std::vector<const JtGraphChunkHdr *> hs;
        for (int i = 0; i < n_chunks; i++) {
            const JtGraphChunkHdr *h = (const JtGraphChunkHdr *)blobs[i];
            if (!h || h->magic != JT_GCHUNK_MAGIC) return JT_ERR_INVALID_ARG;
            hs.push_back(h);
        }
        std::sort(hs.begin(), hs.end(), [](const JtGraphChunkHdr *a, const JtGraphChunkHdr *b) { return a->first < b->first; });
        int64_t expect = 0;
        for (const JtGraphChunkHdr *h : hs) {
            if (h->first != expect || h->rate != rate || h->channels != channels || h->total_frames != total || h->fmt != fmt) return JT_ERR_INVALID_ARG;
            expect = h->first + h->owned;
        }
        if (expect != total) return JT_ERR_INVALID_ARG;             // chunks must tile the stream
        if (frame_size <= 0) frame_size = 4096;
        GraphRun gd;
        jt_graph_build(nullptr, spec, nullptr, total, rate, channels, fmt, frame_size, false, true, JT_GRAPH_DRY, nullptr, gd);

        const int64_t nt = gd.has_r128 ? gd.r128_sig.n / std::max(gd.r128_sig.rate / 10, 1) : 0;
        const int hs_sp = gd.spec_win / 2;
        const int64_t n_hops = gd.has_spec ? (gd.spec_sig.n + hs_sp - 1) / hs_sp : 0;
        const int s_in = gd.has_ln ? (gd.ln_in_sig.rate + 5) / 10 : 1, s_out = gd.has_ln && gd.ln_linear ? (gd.ln_out_sig.rate + 5) / 10 : 1;
        const int64_t li_n = gd.has_ln ? gd.ln_in_sig.n + gd.ln_in_extra : 0;      // dynamic mode meters the flush frame a second time
        const int64_t li_nt = gd.has_ln ? (li_n + s_in - 1) / s_in : 0, lo_nt = gd.has_ln && gd.ln_linear ? (gd.ln_out_sig.n + s_out - 1) / s_out : 0;
        std::vector<double> hp(nt), hk(nt), ht(nt), lip(li_nt), lik(li_nt), lop(lo_nt), lok(lo_nt);
        std::vector<float> rows((size_t)n_hops * JT_SP_COUNT, 0.f);
        std::vector<char> as(jt_astats_host_bytes());
        bool first_as = true; int as_fmt = 0, as_tc = 0;
        int64_t got_r = 0, got_li = 0, got_lo = 0;
        for (const JtGraphChunkHdr *h : hs) {
            const char *r = (const char *)h + sizeof(JtGraphChunkHdr);
            if (h->r_nticks < 0 || h->r_tick0 < 0 || h->r_tick0 + h->r_nticks > nt) return JT_ERR_INVALID_ARG;
            if (h->n_rows < 0 || h->n_rows > n_hops + 2 || h->li_nticks < 0 || h->lo_nticks < 0 || h->li_tick0 < 0 || h->lo_tick0 < 0 ||
                h->astats_bytes != (int64_t)as.size()) return JT_ERR_INVALID_ARG;
            if (h->bytes != (int64_t)sizeof(JtGraphChunkHdr) + h->r_nticks * 24 + h->n_rows * (8 + JT_SP_COUNT * 4) + h->astats_bytes +
                            (h->li_nticks + h->lo_nticks) * 16) return JT_ERR_INVALID_ARG;
            if (h->r_nticks > 0) {
                memcpy(&hp[h->r_tick0], r, 8 * h->r_nticks); r += 8 * h->r_nticks;
                memcpy(&hk[h->r_tick0], r, 8 * h->r_nticks); r += 8 * h->r_nticks;
                memcpy(&ht[h->r_tick0], r, 8 * h->r_nticks); r += 8 * h->r_nticks;
                got_r += h->r_nticks;
            }
            const char *hops = r; r += 8 * h->n_rows;
            for (int64_t k = 0; k < h->n_rows; k++) {
                int64_t gh; memcpy(&gh, hops + 8 * k, 8);
                if (gh >= 0 && gh < n_hops) memcpy(&rows[(size_t)gh * JT_SP_COUNT], r + (size_t)k * JT_SP_COUNT * 4, JT_SP_COUNT * 4);
            }
            r += (size_t)h->n_rows * JT_SP_COUNT * 4;
            if ((size_t)h->astats_bytes != as.size()) return JT_ERR_INVALID_ARG;
            if (h->astats_n > 0) {
                if (first_as) { memcpy(as.data(), r, as.size()); first_as = false; as_fmt = h->astats_fmt; as_tc = h->astats_tc; }
                else jt_astats_host_merge(as.data(), r);
            }
            r += h->astats_bytes;
            if (h->li_nticks < 0 || h->li_tick0 + h->li_nticks > li_nt || h->lo_nticks < 0 || h->lo_tick0 + h->lo_nticks > lo_nt) return JT_ERR_INVALID_ARG;
            if (h->li_nticks > 0) { memcpy(&lip[h->li_tick0], r, 8 * h->li_nticks); r += 8 * h->li_nticks; memcpy(&lik[h->li_tick0], r, 8 * h->li_nticks); r += 8 * h->li_nticks; got_li += h->li_nticks; }
            if (h->lo_nticks > 0) { memcpy(&lop[h->lo_tick0], r, 8 * h->lo_nticks); r += 8 * h->lo_nticks; memcpy(&lok[h->lo_tick0], r, 8 * h->lo_nticks); r += 8 * h->lo_nticks; got_lo += h->lo_nticks; }
        }
        if (got_r != nt || got_li != li_nt || got_lo != lo_nt) return JT_ERR_INVALID_ARG;       // every tick exactly once

        GraphResult res; memset(&res.ln, 0, sizeof(res.ln));
        if (gd.has_ln) {
            LoudnormMeter mi, mo;
            jt_loudnorm_meter_host_finalize(lip.data(), lik.data(), li_nt, li_n / s_in, s_in, gd.ln_dual, mi);
            res.ln.valid = 1;
            if (gd.ln_linear) {
                jt_loudnorm_meter_host_finalize(lop.data(), lok.data(), lo_nt, gd.ln_out_sig.n / s_out, s_out, gd.ln_dual, mo);
                res.ln.normalization_type = 0;
                res.ln.output_i = mo.I; res.ln.output_tp = 20. * log10(mo.sample_peak); res.ln.output_lra = mo.LRA; res.ln.output_thresh = mo.thresh;
                res.ln.target_offset = gd.ln_I - mo.I;
            } else {
                res.ln.normalization_type = gd.ln_type;
                res.ln.output_i = res.ln.output_tp = res.ln.output_lra = res.ln.output_thresh = res.ln.target_offset = NAN;
            }
            res.ln.input_i = mi.I; res.ln.input_tp = 20. * log10(mi.sample_peak); res.ln.input_lra = mi.LRA; res.ln.input_thresh = mi.thresh;
        }
        if (ln) *ln = res.ln;
        R128Result r128;
        if (gd.has_r128) jt_ebur128_host_finalize(nullptr, hp.data(), hk.data(), gd.r128_tp ? ht.data() : nullptr, nt, gd.r128_sig.rate / 10, gd.r128_dual, r128);
        if (!meta && !n_meta) {
            // accumulated measurements only: no records are built
            if (accumulated) {
                jt_accumulate_frames(gd.frames, gd.has_r128, r128, gd.has_spec, rows, n_hops, accumulated);
                if (gd.has_astats && gd.last_astats_frame >= 0 && !first_as) {
                    AstatsResult a;
                    jt_astats_host_finalize(as.data(), gd.frames[gd.last_astats_frame].astats_pos, as_fmt, as_tc, a);
                    if (!gd.astats_overall_only) for (int k = 0; k < JT_AS_COUNT; k++) accumulated->astats[k] = std::isnan(a.v[k]) ? NAN : jt_wire("%f", a.v[k]);
                }
                accumulated->duration_s = gd.out.rate > 0 ? (double)gd.out.n / gd.out.rate : 0.0;
            }
            return JT_OK;
        }
        jt_assemble_records(gd.frames, gd.has_r128, r128, gd.has_spec, rows, n_hops, res);
        if (gd.has_astats && gd.last_astats_frame >= 0 && !first_as) {
            AstatsResult a;
            jt_astats_host_finalize(as.data(), gd.frames[gd.last_astats_frame].astats_pos, as_fmt, as_tc, a);
            jt_frame_meta &m = res.meta[gd.last_astats_frame];
            if (!gd.astats_overall_only) for (int k = 0; k < JT_AS_COUNT; k++) m.astats[k] = std::isnan(a.v[k]) ? NAN : jt_wire("%f", a.v[k]);
            m.astats_overall_RMS_level = jt_wire("%f", a.overall_rms);
            m.astats_overall_Peak_level = jt_wire("%f", a.overall_peak);
        }
        if (n_meta) *n_meta = (int64_t)res.meta.size();
        if (meta) {
            if ((int64_t)res.meta.size() > meta_cap) return JT_ERR_BUFFER;
            if (!res.meta.empty()) memcpy(meta, res.meta.data(), res.meta.size() * sizeof(jt_frame_meta));
        }
        if (accumulated) {
            MeasAcc acc; for (auto &m : res.meta) acc.add(m);
            acc.finish(gd.out.rate > 0 ? (double)gd.out.n / gd.out.rate : 0.0);
            *accumulated = acc.m;
        }
    } catch (const JtError &e) { return e.code; }
    catch (const std::bad_alloc &) { return JT_ERR_NOMEM; }
    return JT_OK;
}

// test hook (tests/test_shard_gloo.py, tests/test_gpu_graph_shard.py): the device-free plan of a graph over a stream of
// n frames -- sink length / rate / format, sink-frame count, lengths of the measuring links
extern "C" int jt_debug_graph_plan(const char *spec, int64_t n, int rate, int channels, int fmt, int frame_size, int want_pcm, int64_t *out /* 10 */)
{
    if (!spec || !out) return JT_ERR_INVALID_ARG;
    try {
        GraphRun gd;
        jt_graph_build(nullptr, spec, nullptr, n, rate, channels, fmt, frame_size, want_pcm != 0, true, JT_GRAPH_DRY, nullptr, gd);
        out[0] = gd.out.n; out[1] = gd.out.rate; out[2] = gd.out.fmt; out[3] = (int64_t)gd.frames.size();
        out[4] = gd.has_r128 ? gd.r128_sig.n : -1; out[5] = gd.has_spec ? gd.spec_sig.n : -1; out[6] = gd.has_astats ? gd.astats_sig.n : -1;
        out[7] = gd.has_ln ? gd.ln_in_sig.n : -1; out[8] = gd.has_ln ? gd.ln_in_sig.rate : -1;
        out[9] = gd.last_astats_frame >= 0 ? gd.frames[gd.last_astats_frame].astats_pos : -1;
    } catch (const JtError &e) { return e.code; }
    return JT_OK;
}

// test hook (tests/test_frame_cadence.py): the sink-frame list of the device-free plan, six int64 per frame
// (start, nb, ready, astats_pos, hop, tick); returns the number of sink frames (or a negative error code)
extern "C" int64_t jt_debug_graph_frames(const char *spec, int64_t n, int rate, int channels, int fmt, int frame_size, int want_pcm, int64_t *out, int64_t cap_frames)
{
    if (!spec) return JT_ERR_INVALID_ARG;
    try {
        GraphRun gd;
        jt_graph_build(nullptr, spec, nullptr, n, rate, channels, fmt, frame_size, want_pcm != 0, true, JT_GRAPH_DRY, nullptr, gd);
        const int64_t nf = (int64_t)gd.frames.size();
        if (out) {
            if (nf > cap_frames) return JT_ERR_BUFFER;
            for (int64_t i = 0; i < nf; i++) {
                const FrameRef &f = gd.frames[(size_t)i];
                int64_t *o = out + 6 * i;
                o[0] = f.start; o[1] = f.nb; o[2] = f.ready; o[3] = f.astats_pos; o[4] = f.hop; o[5] = f.tick;
            }
        }
        return nf;
    } catch (const JtError &e) { return e.code; }
    catch (const std::bad_alloc &) { return JT_ERR_NOMEM; }
}

// ---------------------------------------------------------------------------------------
// K20: band RMS batch (measureSpeechBandRMS, analyser_bands.go:33-104)
// ---------------------------------------------------------------------------------------
extern "C" int jt_band_rms(jt_ctx *c, const void *pcm_in, int64_t n_frames, int rate, int channels, int fmt,
                           double start_s, double duration_s, const double *lo, const double *hi, int n_bands,
                           double *rms_db, int32_t *found)
{
    return guarded(c, [&]() {
        if (!pcm_in || !lo || !hi || !rms_db || n_bands <= 0) JT_THROW(JT_ERR_INVALID_ARG, "null argument");
        if (start_s < 0) JT_THROW(JT_ERR_INVALID_ARG, "invalid region: negative start time");
        if (duration_s <= 0) JT_THROW(JT_ERR_INVALID_ARG, "invalid region: non-positive duration");
        const void *d_in = upload(c, pcm_in, (size_t)n_frames * channels * jt_fmt_bytes(fmt));
        Sig mono = jt_downmix(c, d_in, n_frames, channels, fmt, rate);
        const int64_t st_us = llround(start_s * 1e6), du_us = llround(duration_s * 1e6);
        const int64_t s0 = (st_us * rate + 500000) / 1000000, len = (du_us * rate + 500000) / 1000000;
        Sig reg = jt_slice(mono, s0, len);
        std::vector<int32_t> fnd(n_bands, 0);
        jt_band_rms_batch(c, reg, lo, hi, n_bands, rms_db, fnd.data());
        if (found) memcpy(found, fnd.data(), sizeof(int32_t) * n_bands);
    });
}

// ---------------------------------------------------------------------------------------
// a9 planners (normalise.go:373-392, 407-425, 539-561, 583-585, 611-632, 1198-1203)
// ---------------------------------------------------------------------------------------
static const double kMinLimiterCeilingDB = -24.0, kBrickwallHeadroomDB = 0.9, kCushionDB = 0.2, kLinearSafety = 0.1;
static const double kTPMax = 0.0, kTPMin = -9.0;
static double db_to_lin(double db) { return pow(10, db / 20.0); }

static std::string pre_limiter_prefix(double preGain, double ceiling, bool needed)
{
    if (!needed) return "";
    char b[256]; std::string s;
    if (preGain > 0) { snprintf(b, sizeof(b), "volume=%.1fdB,", preGain); s += b; }
    snprintf(b, sizeof(b), "alimiter=limit=%.6f:attack=5:release=100:level_in=1:level_out=1:level=0:latency=1:asc=1:asc_level=0.8", db_to_lin(ceiling));
    return s + b;
}

extern "C" int jt_build_pass3_spec(double outI, double outTP, double tI, double tTP, double tLRA, char *buf, size_t cap, jt_process_result *plan)
{
    if (!buf) return JT_ERR_INVALID_ARG;
    // calculateLimiterCeiling
    const double gainRequired = tI - outI, projectedTP = outTP + gainRequired;
    double ceiling = 0; bool needed = false, clamped = false;
    if (projectedTP > tTP) { ceiling = tTP - gainRequired; needed = true; if (ceiling < kMinLimiterCeilingDB) { ceiling = kMinLimiterCeilingDB; clamped = true; } }
    // calculatePreGain
    double preGain = 0, reCeil = 0;
    { const double ideal = tTP - gainRequired;
      if (ideal < kMinLimiterCeilingDB) { preGain = kMinLimiterCeilingDB - ideal; const double postI = outI + preGain; reCeil = tTP - (tI - postI); } }
    if (clamped) ceiling = reCeil;
    std::string prefix = pre_limiter_prefix(preGain, ceiling, needed);
    char ln[256];
    snprintf(ln, sizeof(ln), "loudnorm=I=%.1f:TP=%.1f:LRA=%.1f:dual_mono=true:print_format=json", tI, tTP, tLRA);
    std::string spec = prefix.empty() ? std::string(ln) : prefix + "," + ln;
    if (plan) {
        plan->limiter_ceiling_db = ceiling; plan->limiter_pregain_db = preGain; plan->gain_db = gainRequired;
        plan->limiter_needed = needed; plan->limiter_clamped = clamped;
    }
    return copy_str(spec.c_str(), buf, cap);
}

extern "C" int jt_build_pass4_spec(const jt_process_result *plan, const jt_loudnorm_stats *p3, double tI, double tTP, double tLRA,
                                   int source_rate, char *buf, size_t cap, double *eff_out, double *off_out)
{
    if (!plan || !p3 || !buf) return JT_ERR_INVALID_ARG;
    // the Go side parses the JSON strings, i.e. sees "%.2f"-rounded values (normalise.go:321-343)
    const double mI = jt_wire("%.2f", p3->input_i), mTP = jt_wire("%.2f", p3->input_tp);
    const double mLRA = jt_wire("%.2f", p3->input_lra), mTh = jt_wire("%.2f", p3->input_thresh);
    const double internalTP = mTP + (tI - mI) + kLinearSafety + kCushionDB;             // loudnormInternalTargetTP
    const double maxLinear = internalTP - mTP + mI - kLinearSafety;                      // calculateLinearModeTarget
    double eff = tI; if (!(tI <= maxLinear)) eff = maxLinear;
    const double offset = eff - mI;
    const double emittedTP = std::max(kTPMin, std::min(internalTP, kTPMax));             // loudnormTPTargets
    const double brickwall = tTP - kBrickwallHeadroomDB;
    std::string spec = pre_limiter_prefix(plan->limiter_pregain_db, plan->limiter_ceiling_db, plan->limiter_needed != 0);
    if (!spec.empty()) spec += ",";
    char b[1024];
    snprintf(b, sizeof(b), "loudnorm=I=%.2f:TP=%.2f:LRA=%.1f:measured_I=%.2f:measured_TP=%.2f:measured_LRA=%.2f:measured_thresh=%.2f:offset=%.2f:dual_mono=true:linear=true:print_format=json",
             eff, emittedTP, tLRA, mI, mTP, mLRA, mTh, offset);
    spec += b;
    if (source_rate > 0) { snprintf(b, sizeof(b), ",aresample=%d", source_rate); spec += b; }
    spec += ",adeclick=t=1.7:w=55:o=50:m=s";
    snprintf(b, sizeof(b), ",alimiter=limit=%.6f:attack=1:release=50:level_in=1:level_out=1:level=0:latency=1:asc=1:asc_level=0.8", db_to_lin(brickwall));
    spec += b;
    spec += ",astats=metadata=1:measure_perchannel=all,aspectralstats=win_size=2048:win_func=hann:measure=all,"
            "ebur128=metadata=1:peak=sample+true:dualmono=true,"
            "aformat=sample_rates=44100:channel_layouts=mono:sample_fmts=s16,asetnsamples=n=4096";
    if (eff_out) *eff_out = eff;
    if (off_out) *off_out = offset;
    return copy_str(spec.c_str(), buf, cap);
}

// ---------------------------------------------------------------------------------------
// a7: measureOutputRegionFromReader (analyser_output.go:95-233) -- the region graph of analyser_output.go:18 and the
// Go-side reduction of its sink frames into a RegionSample
// ---------------------------------------------------------------------------------------
static double go_seconds(int64_t ns) { return (double)(ns / 1000000000LL) + (double)(ns % 1000000000LL) / 1e9; }     // time.Duration.Seconds

// sample_a / sample_b >= 0: the region as an explicit sample range of the buffer (a buffer that starts on the stream's
// decoder-frame grid: the sharded path, which assembles a region from the ranks that own its parts)
// a region re-measure whose kernels are queued but whose host part has not run: lets ProcessAudio put all four of them behind
// Pass 4 without the GPU waiting for the host in between
struct RegionPending { GraphRun g; bool queued = false; };
static void region_measure_enqueue(jt_ctx *c, const void *d_pcm, int64_t n_frames, int rate, int channels, int fmt,
                                   int64_t start_ns, int64_t dur_ns, RegionPending &p, int64_t sample_a = -1, int64_t sample_b = -1,
                                   bool keep_buffers = false /* graphs on the side stream: the arena must not hand their blocks to the main stream */)
{
    char spec[512];
    if (sample_a >= 0) {
        snprintf(spec, sizeof(spec), "atrim=start_sample=%lld:end_sample=%lld,asetpts=PTS-STARTPTS,astats=metadata=1:measure_perchannel=0,"
                                     "aspectralstats=measure=all,ebur128=metadata=1:peak=sample+true", (long long)sample_a, (long long)sample_b);
    } else {
    if (start_ns < 0) JT_THROW(JT_ERR_INVALID_ARG, "invalid region: negative start time");
    if (dur_ns <= 0) JT_THROW(JT_ERR_INVALID_ARG, "invalid region: non-positive duration");
    snprintf(spec, sizeof(spec), "atrim=start=%f:duration=%f,asetpts=PTS-STARTPTS,astats=metadata=1:measure_perchannel=0,"
                                 "aspectralstats=measure=all,ebur128=metadata=1:peak=sample+true", go_seconds(start_ns), go_seconds(dur_ns));
    }
    const size_t mark = c->allocs.size();
    jt_graph_enqueue(c, spec, d_pcm, n_frames, rate, channels, fmt, 4096, false, true, p.g);
    if (!keep_buffers) jt_release_since(c, mark, nullptr);
    p.queued = true;
}
static void region_measure_finish(jt_ctx *c, RegionPending &p, jt_region_sample *out, int64_t *frames_out)
{
    GraphResult g;
    jt_graph_finish(c, p.g, g);
    double rms = 0, peak = 0, M = 0, S = 0, tp = 0, sp = 0, spec_sum[JT_SP_COUNT] = {0}; bool rms_found = false; int64_t spec_n = 0;
    for (const jt_frame_meta &f : g.meta) {
        if (!std::isnan(f.astats_overall_RMS_level)) { rms = f.astats_overall_RMS_level; rms_found = true; }
        if (!std::isnan(f.astats_overall_Peak_level)) peak = f.astats_overall_Peak_level;
        bool sf = false;
        for (int k = 0; k < JT_SP_COUNT; k++) if (!std::isnan(f.spectral[k])) sf = true;
        if (sf) { for (int k = 0; k < JT_SP_COUNT; k++) spec_sum[k] += std::isnan(f.spectral[k]) ? 0.0 : f.spectral[k]; spec_n++; }
        if (!std::isnan(f.r128_M)) M = f.r128_M;
        if (!std::isnan(f.r128_S)) S = f.r128_S;
        if (!std::isnan(f.r128_true_peak)) tp = f.r128_true_peak;
        if (!std::isnan(f.r128_sample_peak)) sp = f.r128_sample_peak;
    }
    if (frames_out) *frames_out = (int64_t)g.meta.size();
    if (g.meta.empty()) JT_THROW(JT_ERR_INVALID_ARG, "no frames processed in region");
    jt_region_sample r; memset(&r, 0, sizeof(r));
    r.rms_level = rms_found ? rms : -60.0;                                   // conservative fallback (analyser_output.go:222-224)
    r.peak_level = peak;
    r.crest_factor = (rms_found && peak != 0) ? peak - rms : 0.0;            // lavfi.astats.Overall has no Crest_factor key
    for (int k = 0; k < JT_SP_COUNT; k++) r.spectral[k] = spec_n ? spec_sum[k] / (double)spec_n : 0.0;
    r.momentary_lufs = M; r.short_term_lufs = S; r.true_peak = ratio_db(tp); r.sample_peak = ratio_db(sp);
    if (out) *out = r;
}
static void region_measure_device(jt_ctx *c, const void *d_pcm, int64_t n_frames, int rate, int channels, int fmt,
                                  int64_t start_ns, int64_t dur_ns, jt_region_sample *out, int64_t *frames_out,
                                  int64_t sample_a = -1, int64_t sample_b = -1)
{
    RegionPending p;
    region_measure_enqueue(c, d_pcm, n_frames, rate, channels, fmt, start_ns, dur_ns, p, sample_a, sample_b);
    region_measure_finish(c, p, out, frames_out);
}

extern "C" int jt_measure_output_region(jt_ctx *c, const void *pcm, int64_t n_frames, int rate, int channels, int fmt,
                                        int64_t start_ns, int64_t dur_ns, jt_region_sample *out, int64_t *frames)
{
    return guarded(c, [&]() {
        if (!pcm && n_frames > 0) JT_THROW(JT_ERR_INVALID_ARG, "null input");
        if (!jt_valid_fmt(fmt)) JT_THROW(JT_ERR_UNSUPPORTED, "input sample format %d", fmt);
        const void *d_in = upload(c, pcm, (size_t)n_frames * channels * jt_fmt_bytes(fmt));
        region_measure_device(c, d_in, n_frames, rate, channels, fmt, start_ns, dur_ns, out, frames);
    });
}
extern "C" int jt_measure_output_region_dev(jt_ctx *c, const void *d_pcm, int64_t n_frames, int rate, int channels, int fmt,
                                        int64_t start_ns, int64_t dur_ns, jt_region_sample *out, int64_t *frames)
{
    return guarded(c, [&]() { region_measure_device(c, d_pcm, n_frames, rate, channels, fmt, start_ns, dur_ns, out, frames); });
}

// MeasureOutputRegions (analyser_output.go:261-297): the elected room-tone and speech regions of Pass 1 re-measured on a
// pass's s16 output; a failing region is a warning, not an error (its sample stays absent)
struct OutputRegionsPending { RegionPending room, speech; };
// which: 1 = room tone, 2 = speech, 3 = both.  on_side: the graphs go to the context's side stream, ordered behind everything
// queued on the main stream so far (their buffers come from the arena, which hands blocks out again relying on stream order).
// A region graph is a dozen tiny launches over <= 60 s of audio: on the main stream the four of a call were 3.4 ms of a GPU
// that sat mostly empty; on the side stream they run under the next pass's kernels.
static void measure_output_regions_enqueue(jt_ctx *c, const void *d_pcm, int64_t n, const jt_voice_activity &va, OutputRegionsPending &p,
                                           int which = 3, bool on_side = false)
{
    cudaStream_t main_stream = c->stream;
    if (on_side && c->side_stream) {
        cudaEvent_t behind = jt_record_event(c);
        JT_CUDA(cudaStreamWaitEvent(c->side_stream, behind, 0));
        c->stream = c->side_stream;
    }
    const bool keep = c->stream != main_stream;
    try {
        if ((which & 1) && va.has_noise_profile) {
            try { region_measure_enqueue(c, d_pcm, n, 44100, 1, JT_FMT_S16, va.noise_profile.start_ns, va.noise_profile.duration_ns, p.room, -1, -1, keep); }
            catch (const JtError &e) { p.room.queued = false; if (e.code == JT_ERR_CUDA || e.code == JT_ERR_CANCELLED) throw; }
        }
        if ((which & 2) && va.has_speech_profile) {
            try { region_measure_enqueue(c, d_pcm, n, 44100, 1, JT_FMT_S16, va.speech_profile.region.start_ns, va.speech_profile.region.duration_ns, p.speech, -1, -1, keep); }
            catch (const JtError &e) { p.speech.queued = false; if (e.code == JT_ERR_CUDA || e.code == JT_ERR_CANCELLED) throw; }
        }
    } catch (...) { c->stream = main_stream; throw; }
    c->stream = main_stream;
}
static void measure_output_regions_finish(jt_ctx *c, OutputRegionsPending &p, jt_output_regions *o)
{
    memset(o, 0, sizeof(*o));
    if (p.room.queued) {
        try { region_measure_finish(c, p.room, &o->room_tone, nullptr); o->has_room_tone = 1; }
        catch (const JtError &e) { if (e.code == JT_ERR_CUDA || e.code == JT_ERR_CANCELLED) throw; }
    }
    if (p.speech.queued) {
        try { region_measure_finish(c, p.speech, &o->speech, nullptr); o->has_speech = 1; }
        catch (const JtError &e) { if (e.code == JT_ERR_CUDA || e.code == JT_ERR_CANCELLED) throw; }
    }
}
static void measure_output_regions(jt_ctx *c, const void *d_pcm, int64_t n, const jt_voice_activity &va, jt_output_regions *o)
{
    OutputRegionsPending p;
    measure_output_regions_enqueue(c, d_pcm, n, va, p);
    measure_output_regions_finish(c, p, o);
}

// ---------------------------------------------------------------------------------------
// the four-pass chain (ProcessAudio, processor.go:78-216)
// ---------------------------------------------------------------------------------------
// The four passes are data-dependent only through a handful of scalars (Pass 3 is planned from Pass 2's I / TP,
// Pass 4 from Pass 3's loudnorm measurement), so the host-side assembly of one pass's metadata runs while the
// GPU is already working on the next pass: graphs are enqueued (jt_graph_enqueue) ahead of being finished.
static void process_device(jt_ctx *c, const void *d_in, int64_t n_frames, int rate, int channels, int fmt, const char *pass2_spec,
                           int16_t *pcm_out, bool out_on_device, int64_t cap, jt_process_result *res,
                           const jt_measurements *pass1_done = nullptr, const jt_filter_config *cfg = nullptr, jt_analysis *an = nullptr,
                           const GraphResume *head = nullptr, size_t head_mark = (size_t)-1)
{
    jt_process_result R; memset(&R, 0, sizeof(R));
    // defaultLoudnormConfig, filters.go:523-532 (or the caller's base config on the adaptive path)
    const double tI = cfg ? cfg->loudnorm.target_i : -16.0, tTP = cfg ? cfg->loudnorm.target_tp : -1.0, tLRA = cfg ? cfg->loudnorm.target_lra : 20.0;
    const size_t mark = head_mark != (size_t)-1 ? head_mark : c->allocs.size();
    // Pass 1 and Pass 2: device work.  Both read the input only (Pass 2's spec comes from the caller), so Pass 2's
    // long kernels go first and the integer bookkeeping of Pass 1's 200 000 frames happens behind them
    // (adaptive path: the spec-independent head of Pass 2 is already running, `head` is the state behind it)
    GraphRun g2;
    jt_graph_enqueue(c, pass2_spec ? pass2_spec : PASS2_DEFAULT_SPEC, d_in, n_frames, rate, channels, fmt, 4096, true, true, g2, head);
    if (g2.out.fmt != JT_FMT_S16 || g2.out.rate != 44100) JT_THROW(JT_ERR_SPEC, "Pass-2 spec must end in the s16/44.1 kHz output stage (processor.go:379-384)");
    jt_release_since(c, mark, g2.out.d);          // keep only the Pass-2 output ("the FLAC on disk")
    jt_trace(c, "pass2 enqueued");
    jt_check_cancel(c);
    // processor.go:150-160: the elected regions re-measured on Pass 2's output.  On the side stream, behind Pass 2's last kernel:
    // they run under Pass 3 / Pass 4.  (Enqueued here, before Pass 3 takes and returns its buffers, so that nothing they hold was
    // handed back while a main-stream kernel could still be using it.)
    OutputRegionsPending reg2, reg4;
    const bool regions_on_side = an && c->side_stream && !c->timing && !getenv("JT_REGIONS_MAIN");
    if (an && regions_on_side) measure_output_regions_enqueue(c, g2.out.d, g2.out.n, an->voice_activity, reg2, 3, true);
    const size_t mark1 = c->allocs.size();
    AnalysePending p1;
    if (!pass1_done) {
        analyse_enqueue(c, d_in, n_frames, rate, channels, fmt, 4096, p1);
        jt_release_since(c, mark1, nullptr);
    }
    // Pass 3 is planned from Pass 2's integrated loudness and true peak as the sink frames report them
    // (last-seen values of lavfi.r128.I / true_peak, "%.3f": analyser_metrics.go:898-923)
    double out_i = 0.0, out_tp = 0.0;
    {
        const R128Result &r = jt_graph_r128_early(c, g2);
        int64_t last_tick = -1;
        for (const FrameRef &fr : g2.frames) if (fr.tick >= 0 && fr.tick < r.n_ticks) last_tick = std::max<int64_t>(last_tick, fr.tick);
        if (last_tick >= 0) { out_i = jt_wire("%.3f", r.I); out_tp = ratio_db(jt_wire("%.3f", r.tp_cum[last_tick])); }
    }
    char spec3[1024], spec4[4096];
    int rc = jt_build_pass3_spec(out_i, out_tp, tI, tTP, tLRA, spec3, sizeof(spec3), &R);
    if (rc) JT_THROW(rc, "pass-3 spec");
    // Pass 3: the Pass-2 output is re-read as the s16 FLAC the reference wrote (processor.go:126-146)
    const size_t mark3 = c->allocs.size();
    GraphRun g3;
    jt_trace(c, "pass3 planned");
    jt_graph_enqueue(c, spec3, g2.out.d, g2.out.n, 44100, 1, JT_FMT_S16, 4096, false, false, g3);
    jt_release_since(c, mark3, nullptr);
    jt_trace(c, "pass3 enqueued");
    // Pass 1: host part, while the GPU runs Pass 1 and Pass 3
    if (pass1_done) R.input = *pass1_done; else analyse_finish(c, p1, &R.input, nullptr, 0, nullptr);
    jt_check_cancel(c);
    // Pass 3's four numbers gate everything that follows: wait for them first (the GPU is still busy with Pass 2's
    // analysis tail and Pass 3 itself), so Pass 4 can be enqueued before any other host work
    {
        GraphResult r3;
        jt_graph_finish(c, g3, r3);
        R.pass3 = r3.ln;
    }
    jt_trace(c, "pass3 finished");
    const double mI = jt_wire("%.2f", R.pass3.input_i);
    if (std::isinf(mI) || std::isnan(mI) || mI < -70.0) JT_THROW(JT_ERR_INVALID_ARG, "cannot normalise silent audio (measured %.1f LUFS)", mI);
    jt_check_cancel(c);
    // Pass 4
    double eff = 0, off = 0;
    rc = jt_build_pass4_spec(&R, &R.pass3, tI, tTP, tLRA, 44100, spec4, sizeof(spec4), &eff, &off);
    if (rc) JT_THROW(rc, "pass-4 spec");
    R.effective_target_i = eff; R.linear_possible = eff == tI;
    GraphRun g4;
    c->meter_after_output = pcm_out != nullptr && !out_on_device && !getenv("JT_METER_FIRST");     // nothing waits for Pass 4's meter but the end of the call
    try { jt_graph_enqueue(c, spec4, g2.out.d, g2.out.n, 44100, 1, JT_FMT_S16, 4096, true, true, g4); } catch (...) { c->meter_after_output = false; throw; }
    c->meter_after_output = false;
    jt_trace(c, "pass4 enqueued");
    R.n_out = g4.out.n;
    if (pcm_out && g4.out.n > cap) JT_THROW(JT_ERR_BUFFER, "pcm_out holds %lld samples, chain produced %lld", (long long)cap, (long long)g4.out.n);
    // processor.go:150-160 and normalise.go's Pass-4 re-measure of the same regions: queued behind Pass 4 now, so the GPU goes
    // straight on with them while the host is still assembling Pass 2's and Pass 4's metadata
    if (an && !regions_on_side) {
        measure_output_regions_enqueue(c, g2.out.d, g2.out.n, an->voice_activity, reg2);
        measure_output_regions_enqueue(c, g4.out.d, g4.out.n, an->voice_activity, reg4);
    } else if (an) {
        // normalise.go's re-measure of the same regions on Pass 4's output: one on each stream
        measure_output_regions_enqueue(c, g4.out.d, g4.out.n, an->voice_activity, reg4, 2, true);
        measure_output_regions_enqueue(c, g4.out.d, g4.out.n, an->voice_activity, reg4, 1, false);
    }
    jt_trace(c, "regions enqueued");
    // Pass 2: host part (sink-frame records, accumulators), while the GPU runs Pass 4
    jt_graph_finish_acc(c, g2, &R.filtered, nullptr);
    jt_trace(c, "pass2 accumulated");
    if (pcm_out) {                                 // the result leaves while the host assembles Pass 4's metadata
        const size_t ob = (size_t)g4.out.n * sizeof(int16_t);
        // the copy engine takes the result as soon as its last sample exists (ev_out), while the compute stream goes on
        // with Pass 4's analysis tail
        if (ob && g4.out_ready) {
            JT_CUDA(cudaStreamWaitEvent(c->copy_stream, g4.out_ready, 0));
            if (out_on_device) JT_CUDA(cudaMemcpyAsync(pcm_out, g4.out.d, ob, cudaMemcpyDeviceToDevice, c->copy_stream));
            else JT_CUDA(copy_in_pieces(pcm_out, g4.out.d, ob, cudaMemcpyDeviceToHost, c->copy_stream));
        } else if (ob) JT_CUDA(cudaMemcpyAsync(pcm_out, g4.out.d, ob, out_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, c->stream));
    }
    jt_graph_finish_acc(c, g4, &R.final, &R.pass4);
    jt_trace(c, "pass4 accumulated");
    if (an) {
        measure_output_regions_finish(c, reg2, &an->filtered_regions);
        measure_output_regions_finish(c, reg4, &an->final_regions);
    }
    jt_trace(c, "regions finished");
    JT_CUDA(cudaStreamSynchronize(c->stream));
    if (c->side_stream) JT_CUDA(cudaStreamSynchronize(c->side_stream));
    JT_CUDA(cudaStreamSynchronize(c->copy_stream));
    jt_trace(c, "end");
    if (res) *res = R;
}

extern "C" int jt_process_audio(jt_ctx *c, const void *pcm_in, int64_t n_frames, int rate, int channels, int fmt,
                                const char *pass2_spec, int16_t *pcm_out, int64_t cap, jt_process_result *res)
{
    return guarded(c, [&]() {
        if (!pcm_in && n_frames > 0) JT_THROW(JT_ERR_INVALID_ARG, "null input");
        if (!jt_valid_fmt(fmt)) JT_THROW(JT_ERR_UNSUPPORTED, "input sample format %d", fmt);
        const void *d_in = upload(c, pcm_in, (size_t)n_frames * channels * jt_fmt_bytes(fmt));
        process_device(c, d_in, n_frames, rate, channels, fmt, pass2_spec, pcm_out, false, cap, res);
    });
}
extern "C" int jt_process_audio_dev(jt_ctx *c, const void *d_in, int64_t n_frames, int rate, int channels, int fmt,
                                const char *pass2_spec, int16_t *d_out, int64_t cap, jt_process_result *res)
{
    return guarded(c, [&]() {
        if (!d_in && n_frames > 0) JT_THROW(JT_ERR_INVALID_ARG, "null input");
        process_device(c, d_in, n_frames, rate, channels, fmt, pass2_spec, d_out, true, cap, res);
    });
}

// ---------------------------------------------------------------------------------------
// The adaptive path on the library side: AnalyseAudio + AdaptConfig (analyser.go:325-372, processor.go:37-69) and
// ProcessAudio with the adapted Pass-2 spec (processor.go:78-216).  Pass 2 depends on Pass 1 through the detector,
// so the two cannot overlap here as they do in process_device with a caller-supplied spec.
// ---------------------------------------------------------------------------------------
// issue the enclosed work on another stream of the context; on leaving, everything issued there is done (so its buffers can be
// released) and the context is back on its own stream
struct StreamSwap {
    jt_ctx *c; cudaStream_t main;
    StreamSwap(jt_ctx *ctx, cudaStream_t to) : c(ctx), main(ctx->stream) { c->stream = to; }
    ~StreamSwap() { cudaStreamSynchronize(c->stream); c->stream = main; }
};

static void analyse_adaptive_device(jt_ctx *c, const void *d_in, int64_t n_frames, int rate, int channels, int fmt, int F,
                                    const jt_filter_config *base, jt_analysis *out, jt_interval *iv_out, int64_t iv_cap, int64_t *n_iv_out,
                                    AnalysePending *pending = nullptr /* Pass 1 already enqueued by the caller */,
                                    cudaEvent_t input_ready = nullptr /* with it: the band graphs run on the side stream */)
{
    if (!out) JT_THROW(JT_ERR_INVALID_ARG, "null analysis");
    memset(out, 0, sizeof(*out));
    const size_t mark = c->allocs.size();
    std::vector<jt_interval> own;
    jt_interval *iv = iv_out; int64_t cap = iv_cap;
    if (!iv) { cap = (int64_t)((double)n_frames / rate / 0.25) + 16; own.resize((size_t)cap); iv = own.data(); }
    int64_t n_iv = 0;
    // astats_later: Pass 1's astats runs on the low-priority stream (process_adaptive_device); the intervals, the detector and
    // the band graphs do not read it, so its values are collected after them
    const bool astats_later = pending && pending->g.astats_later;
    if (pending) analyse_finish(c, *pending, &out->measurements, iv, cap, &n_iv);
    else analyse_device(c, d_in, n_frames, rate, channels, fmt, F, &out->measurements, iv, cap, &n_iv);
    jt_trace(c, "pass1 intervals done");
    if (n_iv_out) *n_iv_out = n_iv;
    if (out->measurements.sink_frames == 0 || std::isnan(out->measurements.input_i))
        JT_THROW(JT_ERR_INVALID_ARG, "ebur128 measurements not found in metadata (analyser.go:397-399)");
    int rc;
    { JtHost hdet(c, "voice_activity_detector"); rc = jt_detect_voice_activity(&out->measurements, iv, n_iv, &out->voice_activity, nullptr, 0, nullptr, 0); }
    if (rc) JT_THROW(rc, "voice-activity detector");
    jt_trace(c, "detector done");
    jt_voice_activity &va = out->voice_activity;
    // measureSpeechBands + measureNoiseBands: the 2 + 15 band graphs of analyser_bands.go:33 over the elected regions
    double lo[17], hi[17]; jt_band_plan(lo, hi);
    const bool want_speech = va.has_speech_profile && va.speech_profile.region.duration_ns > 0;
    const bool want_noise = va.has_noise_profile && va.noise_profile.duration_ns > 0;
    if (want_speech || want_noise) {
        // The band graphs read the input only.  When Pass 2's head is already queued on the main stream they go to the side
        // stream (ordered behind the input's upload by `input_ready`), so their result does not wait for anlmdn.
        std::unique_ptr<StreamSwap> swap;
        if (input_ready && c->side_stream && !c->timing) {
            JT_CUDA(cudaStreamWaitEvent(c->side_stream, input_ready, 0));
            swap.reset(new StreamSwap(c, c->side_stream));
        }
        Sig mono = jt_downmix(c, d_in, n_frames, channels, fmt, rate);
        auto region = [&](int64_t start_ns, int64_t dur_ns) {
            // atrim=start=%f:duration=%f (analyser_bands.go:54-60): seconds printed with six decimals
            const double st = jt_wire("%f", (double)(start_ns / 1000000000LL) + (double)(start_ns % 1000000000LL) / 1e9);
            const double du = jt_wire("%f", (double)(dur_ns / 1000000000LL) + (double)(dur_ns % 1000000000LL) / 1e9);
            const int64_t st_us = llround(st * 1e6), du_us = llround(du * 1e6);
            const int64_t s0 = (st_us * rate + 500000) / 1000000, len = (du_us * rate + 500000) / 1000000;
            return jt_slice(mono, s0, len);
        };
        double rms[17] = {0}; int32_t found[17] = {0};
        if (want_speech) { Sig r = region(va.speech_profile.region.start_ns, va.speech_profile.region.duration_ns); jt_band_rms_batch(c, r, lo, hi, 2, rms, found); }
        if (want_noise) { Sig r = region(va.noise_profile.start_ns, va.noise_profile.duration_ns); jt_band_rms_batch(c, r, lo + 2, hi + 2, 15, rms + 2, found + 2); }
        swap.reset();
        jt_trace(c, "band graphs done");
        jt_apply_band_rms(&va, want_speech ? rms : nullptr, want_speech ? found : nullptr, want_noise ? rms + 2 : nullptr, want_noise ? found + 2 : nullptr);
    }
    if (astats_later && pending->g.has_astats && pending->g.last_astats_frame >= 0) {
        AstatsResult a;
        jt_astats_finish(c, pending->g.astp, a);
        // what MeasAcc::add takes from the record of the last astats frame (latest wins, absent keys stay NaN)
        if (!pending->g.astats_overall_only)
            for (int k = 0; k < JT_AS_COUNT; k++) { const double w = std::isnan(a.v[k]) ? NAN : jt_wire("%f", a.v[k]); if (!std::isnan(w)) out->measurements.astats[k] = w; }
        jt_vad_assign_astats(&out->measurements, &va);
    }
    { JtHost hcfg(c, "adapt_config"); rc = jt_adapt_config(base, &out->measurements, &va, &out->config, &out->diagnostics); }
    if (rc) JT_THROW(rc, "AdaptConfig");
    rc = jt_build_filter_spec(&out->config, out->pass2_spec, sizeof(out->pass2_spec));
    if (rc) JT_THROW(rc, "BuildFilterSpec");
    jt_trace(c, "spec built");
    jt_release_since(c, mark, nullptr);
}

extern "C" int jt_analyse_adaptive(jt_ctx *c, const void *pcm_in, int64_t n_frames, int rate, int channels, int fmt, int frame_size,
                                   const jt_filter_config *base, jt_analysis *out, jt_interval *iv, int64_t iv_cap, int64_t *n_iv)
{
    return guarded(c, [&]() {
        if (!pcm_in && n_frames > 0) JT_THROW(JT_ERR_INVALID_ARG, "null input");
        if (!jt_valid_fmt(fmt)) JT_THROW(JT_ERR_UNSUPPORTED, "input sample format %d", fmt);
        const void *d_in = upload(c, pcm_in, (size_t)n_frames * channels * jt_fmt_bytes(fmt));
        analyse_adaptive_device(c, d_in, n_frames, rate, channels, fmt, frame_size, base, out, iv, iv_cap, n_iv);
    });
}

// The filters of a Pass-2 spec that AdaptConfig never touches: everything before the first adaptive filter
// (Pass2FilterOrder, filters.go:58-68: downmix, rumble high-pass, band-limit low-pass, anlmdn | afftdn, gate, compressor,
// de-esser, analysis, resample).  Returns the head as a spec string ("" when there is none).
static std::string pass2_static_head(const std::string &spec)
{
    std::vector<FilterNode> nodes = jt_parse_spec(spec);
    size_t n = 0, pos = 0, end = 0;
    for (; n < nodes.size(); n++) {
        const std::string &nm = nodes[n].name;
        if (!(nm == "aformat" || nm == "highpass" || nm == "lowpass" || nm == "anlmdn")) break;
        if (nm == "aformat" && n > 0) break;                    // the output stage, not the downmix
        // the node's text ends at the next top-level comma (none of these filters has escaped commas)
        const size_t comma = spec.find(',', pos);
        end = comma == std::string::npos ? spec.size() : comma;
        pos = end + 1;
    }
    return n ? spec.substr(0, end) : std::string();
}

static void process_adaptive_device(jt_ctx *c, const void *d_in, int64_t n_frames, int rate, int channels, int fmt, const jt_filter_config *base,
                                    int16_t *pcm_out, bool out_on_device, int64_t cap, jt_process_result *res, jt_analysis *analysis)
{
    std::vector<jt_analysis> own(analysis ? 0 : 1);
    jt_analysis *an = analysis ? analysis : own.data();
    // Pass 1's kernels first; behind them, without waiting, the head of Pass 2 that no measurement can change (predicted
    // from AdaptConfig on empty measurements and checked against the real spec below), so the GPU keeps working while
    // the host assembles Pass 1's metadata, runs the detector and derives the spec's adaptive tail.
    std::string head_spec, predicted_spec;
    {
        jt_measurements m0; memset(&m0, 0, sizeof(m0));
        jt_voice_activity va0; memset(&va0, 0, sizeof(va0));
        jt_filter_config cfg0; char spec0[2048];
        if (jt_adapt_config(base, &m0, &va0, &cfg0, nullptr) == JT_OK && jt_build_filter_spec(&cfg0, spec0, sizeof(spec0)) == JT_OK) {
            head_spec = pass2_static_head(spec0);
            predicted_spec = spec0;             // afftdn's forward transforms (independent of its adaptive parameters) join the head
        }
    }
    AnalysePending p1;
    const size_t mark1 = c->allocs.size();
    jt_trace(c, "start");
    // Pass 1's astats is the one product of Pass 1 the host needs LAST (whole-file values for AdaptConfig; the intervals and the
    // detector run on the meter's and aspectralstats' rows): it goes to the low-priority stream, where it fills the SMs the
    // meter / spectral / head kernels leave free and the gap in which the main stream waits for the detector.  Its buffers --
    // and everything else Pass 1 allocated -- are therefore released only after the host has seen Pass 1's last event.
    // (with per-kernel timing on, everything stays on the main stream: an event pair around a kernel that shares the GPU with
    // another stream's kernels measures the sharing, not the kernel)
    const bool defer = c->low_stream != nullptr && !c->timing && !getenv("JT_NO_DEFER_ASTATS");
    c->defer_astats = defer;
    try { analyse_enqueue(c, d_in, n_frames, rate, channels, fmt, 4096, p1); } catch (...) { c->defer_astats = false; throw; }
    c->defer_astats = false;
    p1.g.astats_later = defer && p1.g.astats_on_low;
    if (!defer) jt_release_since(c, mark1, nullptr);
    const size_t mark1_end = c->allocs.size();
    // the side stream starts behind Pass 1's kernels: the input is in HBM by then, and every arena block Pass 1 has already
    // given back (the arena hands blocks out again at once, relying on stream order) is no longer in use
    cudaEvent_t input_ready = jt_record_event(c);
    jt_trace(c, "pass1 enqueued");
    const size_t head_mark = c->allocs.size();
    GraphResume head; bool have_head = false;
    if (!head_spec.empty()) { jt_graph_head(c, head_spec, d_in, n_frames, rate, channels, fmt, 4096, head, &predicted_spec); have_head = true; }
    jt_trace(c, "head enqueued");
    size_t head_mark_now = head_mark;
    try {
        analyse_adaptive_device(c, d_in, n_frames, rate, channels, fmt, 4096, base, an, nullptr, 0, nullptr, &p1, have_head ? input_ready : nullptr);
    } catch (...) {
        if (defer) cudaStreamSynchronize(c->low_stream);
        throw;
    }
    if (defer) {
        // analyse_finish has waited for every event of Pass 1 (astats' included): its blocks can go back to the arena
        JT_CUDA(cudaStreamSynchronize(c->low_stream));
        jt_release_range(c, mark1, mark1_end);
        head_mark_now = head_mark - (mark1_end - mark1);
    }
    jt_check_cancel(c);
    const std::string spec = an->pass2_spec;
    const bool match = have_head && spec.compare(0, head_spec.size(), head_spec) == 0 && (spec.size() == head_spec.size() || spec[head_spec.size()] == ',');
    if (have_head && !match) jt_release_since(c, head_mark_now, nullptr);       // the prediction missed: Pass 2 runs from the input
    process_device(c, d_in, n_frames, rate, channels, fmt, an->pass2_spec, pcm_out, out_on_device, cap, res, &an->measurements, &an->config, an,
                   match ? &head : nullptr, match ? head_mark_now : (size_t)-1);
}
extern "C" int jt_process_audio_adaptive(jt_ctx *c, const void *pcm_in, int64_t n_frames, int rate, int channels, int fmt,
                                         const jt_filter_config *base, int16_t *pcm_out, int64_t cap, jt_process_result *res, jt_analysis *analysis)
{
    return guarded(c, [&]() {
        if (!pcm_in && n_frames > 0) JT_THROW(JT_ERR_INVALID_ARG, "null input");
        if (!jt_valid_fmt(fmt)) JT_THROW(JT_ERR_UNSUPPORTED, "input sample format %d", fmt);
        const void *d_in = upload(c, pcm_in, (size_t)n_frames * channels * jt_fmt_bytes(fmt));
        process_adaptive_device(c, d_in, n_frames, rate, channels, fmt, base, pcm_out, false, cap, res, analysis);
    });
}
extern "C" int jt_process_audio_adaptive_dev(jt_ctx *c, const void *d_in, int64_t n_frames, int rate, int channels, int fmt,
                                         const jt_filter_config *base, int16_t *d_out, int64_t cap, jt_process_result *res, jt_analysis *analysis)
{
    return guarded(c, [&]() {
        if (!d_in && n_frames > 0) JT_THROW(JT_ERR_INVALID_ARG, "null input");
        process_adaptive_device(c, d_in, n_frames, rate, channels, fmt, base, d_out, true, cap, res, analysis);
    });
}

// ---------------------------------------------------------------------------------------
// ONE stream over several GPUs, orchestrated inside the library (BASELINE.json configs[3]): every rank calls
// jt_process_audio_sharded with its window of the input and an all-gather callback (jt_set_exchange); the call runs
// ProcessAudio (processor.go:78-216) with every pass cut into one chunk per rank.  What crosses ranks:
//   - the mergeable measurement blobs of each measuring pass (a few MB per hour of audio), one all-gather each;
//   - afftdn's noise-floor carry (32 bytes per rank), inside Pass 2;
//   - the samples of the elected regions (<= 60 s of speech, <= 18 s of room tone) for the 17 band graphs and for the
//     region re-measures of the Pass-2 / Pass-4 outputs: each rank contributes the part it owns;
//   - the HALO of the Pass-2 output that a neighbour needs for Pass 3 / 4 (whose chunk grid at 44.1 kHz differs from
//     Pass 2's): context + grid shift, tens of seconds -- never the stream.
// The Pass-2 output stays resident on the rank that made it; each rank returns the part of the result it owns.
// ---------------------------------------------------------------------------------------
namespace {
struct ShardComm {
    jt_ctx *c; int rank, world; double seconds = 0; int calls = 0;
    // all-gather of byte strings of any length, built on the fixed-size primitive: lengths first, then padded payloads
    std::vector<std::vector<char>> allgather(const std::vector<char> &send)
    {
        std::vector<std::vector<char>> out((size_t)world);
        if (world == 1 || !c->exchange) { out[0] = send; return out; }
        const double t0 = host_seconds();
        std::vector<int64_t> lens((size_t)world, 0);
        const int64_t mine = (int64_t)send.size();
        int rc = c->exchange(c->exchange_user, &mine, sizeof(int64_t), lens.data());
        if (rc) JT_THROW(JT_ERR_INVALID_ARG, "exchange callback failed (%d)", rc);
        int64_t cap = 0; for (int64_t l : lens) { if (l < 0 || l > ((int64_t)1 << 40)) JT_THROW(JT_ERR_INVALID_ARG, "exchange: bad length"); cap = std::max(cap, l); }
        if (cap > 0) {
            std::vector<char> sb((size_t)cap, 0), rb((size_t)cap * world);
            if (mine) memcpy(sb.data(), send.data(), (size_t)mine);
            rc = c->exchange(c->exchange_user, sb.data(), cap, rb.data());
            if (rc) JT_THROW(JT_ERR_INVALID_ARG, "exchange callback failed (%d)", rc);
            for (int r = 0; r < world; r++) out[(size_t)r].assign(rb.begin() + (size_t)r * cap, rb.begin() + (size_t)r * cap + lens[(size_t)r]);
        }
        seconds += host_seconds() - t0; calls += 2;
        return out;
    }
    static double host_seconds() { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }
};

struct Chunks { std::vector<int64_t> first, owned; };
static Chunks plan_chunks(int64_t total, int64_t unit, int world)
{
    Chunks ck; ck.first.resize((size_t)world); ck.owned.resize((size_t)world);
    const int64_t units = std::max<int64_t>(1, (total + unit - 1) / unit), per = units / world, extra = units % world;
    int64_t first = 0;
    for (int r = 0; r < world; r++) {
        const int64_t nu = per + (r < extra ? 1 : 0);
        const int64_t own = std::min(nu * unit, std::max<int64_t>(total - first, 0));
        ck.first[(size_t)r] = first; ck.owned[(size_t)r] = own; first += own;
    }
    return ck;
}
static int64_t sharded_unit(int rate)
{
    const int64_t u1 = jt_analyse_chunk_unit(rate), u2 = chunk_geometry(PASS2_DEFAULT_SPEC, rate).unit;
    if (u1 <= 0 || u2 <= 0) JT_THROW(JT_ERR_UNSUPPORTED, "sharded processing at %d Hz", rate);
    return lcm64(u1, u2);
}
// sample range of a region on a link, the way trim.c cuts it from "%f"-printed seconds (analyser_output.go:18, analyser_bands.go:54-60)
static void region_range(int64_t start_ns, int64_t dur_ns, int rate, int64_t n_total, int64_t *a, int64_t *b)
{
    const double st = jt_wire("%f", go_seconds(start_ns)), du = jt_wire("%f", go_seconds(dur_ns));
    const int64_t st_us = llround(st * 1e6), du_us = llround(du * 1e6);
    const int64_t s0 = (st_us * rate + 500000) / 1000000, len = du_us > 0 ? (du_us * rate + 500000) / 1000000 : INT64_MAX / 4;
    *a = std::min(std::max<int64_t>(s0, 0), n_total); *b = std::min(n_total, s0 + len);
    if (*b < *a) *b = *a;
}
}   // namespace

extern "C" int jt_sharded_plan(int64_t total_frames, int rate, int n_ranks, int rank, jt_shard_plan *out)
{
    if (!out || n_ranks <= 0 || rank < 0 || rank >= n_ranks || total_frames <= 0) return JT_ERR_INVALID_ARG;
    try {
        const int64_t U = sharded_unit(rate);
        const Chunks ck = plan_chunks(total_frames, U, n_ranks);
        const int64_t left = ((int64_t)8 * rate + U - 1) / U * U, right = ((int64_t)rate + U - 1) / U * U;
        memset(out, 0, sizeof(*out));
        out->unit = U; out->own_first = ck.first[(size_t)rank]; out->owned = ck.owned[(size_t)rank];
        out->local_first = std::max<int64_t>(0, out->own_first - left);
        out->n_local = out->owned > 0 ? std::min(total_frames, out->own_first + out->owned + right) - out->local_first : 0;
    } catch (const JtError &e) { return e.code; }
    return JT_OK;
}

static void process_sharded_device(jt_ctx *c, const void *d_local, int64_t n_local, int rate, int channels, int fmt,
                                   int64_t total, int world, int rank, const jt_filter_config *base, int adaptive,
                                   int16_t *pcm_out, bool out_on_device, int64_t cap, int64_t *out_first, int64_t *n_out,
                                   jt_process_result *res, jt_analysis *analysis, jt_shard_timing *tm)
{
    if (world <= 0 || rank < 0 || rank >= world) JT_THROW(JT_ERR_INVALID_ARG, "bad rank / world");
    if (world > 1 && (!c->exchange || c->exchange_ranks != world)) JT_THROW(JT_ERR_INVALID_ARG, "jt_set_exchange must provide an all-gather over %d ranks", world);
    jt_shard_plan P;
    int rc = jt_sharded_plan(total, rate, world, rank, &P);
    if (rc) JT_THROW(rc, "sharded plan");
    if (P.owned <= 0) JT_THROW(JT_ERR_INVALID_ARG, "the stream is too short for %d ranks (one unit of %lld frames per rank at least)", world, (long long)P.unit);
    if (n_local != P.n_local) JT_THROW(JT_ERR_INVALID_ARG, "window of %lld frames, jt_sharded_plan asks for %lld", (long long)n_local, (long long)P.n_local);
    ShardComm comm{c, rank, world};
    jt_shard_timing T; memset(&T, 0, sizeof(T));
    double t_prev = ShardComm::host_seconds();
    auto lap = [&](double &slot, bool sync = true) {
        if (sync) JT_CUDA(cudaStreamSynchronize(c->stream));
        const double t = ShardComm::host_seconds(); slot += t - t_prev; t_prev = t;
    };
    std::vector<jt_analysis> own_an(analysis ? 0 : 1);
    jt_analysis *an = analysis ? analysis : own_an.data();
    memset(an, 0, sizeof(*an));
    jt_process_result R; memset(&R, 0, sizeof(R));
    const double tI = base ? base->loudnorm.target_i : -16.0, tTP = base ? base->loudnorm.target_tp : -1.0, tLRA = base ? base->loudnorm.target_lra : 20.0;
    const Chunks ck1 = plan_chunks(total, P.unit, world);

    // ---- Pass 1: chunk analysis, one all-gather, merge on every rank ----
    std::vector<jt_interval> iv((size_t)((double)total / rate / 0.25) + 16);
    int64_t n_iv = 0;
    Sig mono;                                             // the downmixed window (device), kept for the band graphs
    GraphResume head2; bool have_head2 = false;
    {
        std::vector<char> blob;
        analyse_chunk_core(c, d_local, n_local, rate, channels, fmt, P.local_first, P.own_first, P.owned, total, blob, &mono);
        lap(T.pass1_chunk);
        // the head of Pass 2 that no measurement can change (downmix, both biquads, anlmdn -- the heaviest kernel of the pass)
        // runs on the window while the host exchanges and merges Pass 1 and derives the spec's adaptive tail
        {
            jt_measurements m0; memset(&m0, 0, sizeof(m0));
            jt_voice_activity va0; memset(&va0, 0, sizeof(va0));
            jt_filter_config cfg0; char spec0[2048];
            std::string hs;
            if (!adaptive) hs = pass2_static_head(PASS2_DEFAULT_SPEC);
            else if (jt_adapt_config(base, &m0, &va0, &cfg0, nullptr) == JT_OK && jt_build_filter_spec(&cfg0, spec0, sizeof(spec0)) == JT_OK) hs = pass2_static_head(spec0);
            if (!hs.empty()) { jt_graph_head(c, hs, d_local, n_local, rate, channels, fmt, 4096, head2); have_head2 = true; }
        }
        std::vector<std::vector<char>> all = comm.allgather(blob);
        std::vector<const void *> ptrs; for (auto &b : all) if (!b.empty()) ptrs.push_back(b.data());
        rc = jt_analyse_merge((int)ptrs.size(), ptrs.data(), &an->measurements, iv.data(), (int64_t)iv.size(), &n_iv);
        if (rc) JT_THROW(rc, "Pass-1 merge");
        R.input = an->measurements;
        lap(T.pass1_merge, false);                        // (no device sync: Pass 2's head is running)
    }
    // ---- detector, band graphs over the elected regions (their samples gathered from the ranks that own them), AdaptConfig ----
    std::string spec2 = PASS2_DEFAULT_SPEC;
    if (adaptive) {
        if (an->measurements.sink_frames == 0 || std::isnan(an->measurements.input_i)) JT_THROW(JT_ERR_INVALID_ARG, "ebur128 measurements not found in metadata (analyser.go:397-399)");
        rc = jt_detect_voice_activity(&an->measurements, iv.data(), n_iv, &an->voice_activity, nullptr, 0, nullptr, 0);
        if (rc) JT_THROW(rc, "voice-activity detector");
        jt_voice_activity &va = an->voice_activity;
        const bool want_speech = va.has_speech_profile && va.speech_profile.region.duration_ns > 0;
        const bool want_noise = va.has_noise_profile && va.noise_profile.duration_ns > 0;
        if (want_speech || want_noise) {
            int64_t ra[2] = {0, 0}, rb[2] = {0, 0};
            if (want_speech) region_range(va.speech_profile.region.start_ns, va.speech_profile.region.duration_ns, rate, total, &ra[0], &rb[0]);
            if (want_noise) region_range(va.noise_profile.start_ns, va.noise_profile.duration_ns, rate, total, &ra[1], &rb[1]);
            // every rank band-filters the part of each region it owns (the filters warm up on the samples before it: the
            // region's own when they are in reach, so a part that starts the region starts from rest as the reference's
            // graph does) and contributes 17 sums of squares
            double lo_hz[17], hi_hz[17]; jt_band_plan(lo_hz, hi_hz);
            std::vector<char> send(17 * sizeof(double), 0);
            double *mine = (double *)send.data();
            // (on the side stream: Pass 2's head occupies the main one; Pass 1's kernels, whose released blocks the arena may
            //  hand out here, finished before the last lap)
            std::unique_ptr<StreamSwap> swap;
            if (have_head2 && c->side_stream) swap.reset(new StreamSwap(c, c->side_stream));
            for (int k = 0; k < 2; k++) {
                const int64_t lo = std::max(ra[k], P.own_first), hi = std::min(rb[k], P.own_first + P.owned);
                if (hi <= lo) continue;
                const int64_t cs = std::max(std::max(ra[k], lo - JT_BAND_WARM_MAX), P.local_first);
                Sig part = jt_slice(mono, cs - P.local_first, hi - cs);
                if (k == 0) jt_band_sumsq(c, part, lo - cs, lo_hz, hi_hz, 2, mine);
                else jt_band_sumsq(c, part, lo - cs, lo_hz + 2, hi_hz + 2, 15, mine + 2);
            }
            swap.reset();
            std::vector<std::vector<char>> all = comm.allgather(send);
            double rms[17] = {0}; int32_t found[17] = {0};
            for (int b = 0; b < 17; b++) {
                const int k = b < 2 ? 0 : 1;
                const int64_t len = rb[k] - ra[k];
                if (len <= 0 || !(k == 0 ? want_speech : want_noise)) continue;
                double sum = 0;
                for (int r = 0; r < world; r++) if (all[(size_t)r].size() == 17 * sizeof(double)) sum += ((const double *)all[(size_t)r].data())[b];
                rms[b] = jt_wire("%f", log10(sqrt(sum / (double)len)) * 20); found[b] = 1;
            }
            jt_apply_band_rms(&va, want_speech ? rms : nullptr, want_speech ? found : nullptr, want_noise ? rms + 2 : nullptr, want_noise ? found + 2 : nullptr);
        }
        rc = jt_adapt_config(base, &an->measurements, &va, &an->config, &an->diagnostics);
        if (rc) JT_THROW(rc, "AdaptConfig");
        rc = jt_build_filter_spec(&an->config, an->pass2_spec, sizeof(an->pass2_spec));
        if (rc) JT_THROW(rc, "BuildFilterSpec");
        spec2 = an->pass2_spec;
    } else copy_str(PASS2_DEFAULT_SPEC, an->pass2_spec, sizeof(an->pass2_spec));
    lap(T.adapt, false);

    // the regions the reference re-measures on the Pass-2 and Pass-4 outputs (MeasureOutputRegions, analyser_output.go:261-297)
    auto measure_regions = [&](const Sig &own, int64_t own_first44, const Chunks &ck44, int64_t n44, jt_output_regions *o) {
        memset(o, 0, sizeof(*o));
        if (!adaptive) return;
        const jt_voice_activity &va = an->voice_activity;
        int64_t a[2] = {0, 0}, b[2] = {0, 0}, g0[2] = {0, 0}; bool want[2] = {false, false};
        int64_t s_ns[2] = {va.noise_profile.start_ns, va.speech_profile.region.start_ns}, d_ns[2] = {va.noise_profile.duration_ns, va.speech_profile.region.duration_ns};
        want[0] = va.has_noise_profile && d_ns[0] > 0 && s_ns[0] >= 0; want[1] = va.has_speech_profile && d_ns[1] > 0 && s_ns[1] >= 0;
        std::vector<char> send;
        for (int k = 0; k < 2; k++) {
            if (!want[k]) continue;
            region_range(s_ns[k], d_ns[k], 44100, n44, &a[k], &b[k]);
            g0[k] = a[k] / 4096 * 4096;                    // the decoder frame holding the region's first sample: same frame grid as the whole stream
            const int64_t lo = std::max(g0[k], own_first44), hi = std::min(b[k], own_first44 + own.n);
            if (hi > lo) {
                const size_t off = send.size(); send.resize(off + (size_t)(hi - lo) * 2);
                JT_CUDA(cudaMemcpyAsync(send.data() + off, (const char *)own.d + (size_t)(lo - own_first44) * 2, (size_t)(hi - lo) * 2, cudaMemcpyDeviceToHost, c->stream));
            }
        }
        JT_CUDA(cudaStreamSynchronize(c->stream));
        std::vector<std::vector<char>> all = comm.allgather(send);
        std::vector<size_t> cursor((size_t)world, 0);
        for (int k = 0; k < 2; k++) {
            if (!want[k]) continue;
            const int64_t len = b[k] - g0[k];
            jt_region_sample *dst = k == 0 ? &o->room_tone : &o->speech;
            if (b[k] - a[k] <= 0) continue;
            std::vector<int16_t> reg((size_t)len);
            for (int r = 0; r < world; r++) {
                const int64_t lo = std::max(g0[k], ck44.first[(size_t)r]), hi = std::min(b[k], ck44.first[(size_t)r] + ck44.owned[(size_t)r]);
                if (hi > lo) {
                    if (cursor[(size_t)r] + (size_t)(hi - lo) * 2 > all[(size_t)r].size()) JT_THROW(JT_ERR_INVALID_ARG, "internal: region gather");
                    memcpy(reg.data() + (lo - g0[k]), all[(size_t)r].data() + cursor[(size_t)r], (size_t)(hi - lo) * 2);
                    cursor[(size_t)r] += (size_t)(hi - lo) * 2;
                }
            }
            try {
                const void *d_reg = upload(c, reg.data(), reg.size() * 2);
                JT_CUDA(cudaStreamSynchronize(c->stream));
                region_measure_device(c, d_reg, len, 44100, 1, JT_FMT_S16, 0, 0, dst, nullptr, a[k] - g0[k], b[k] - g0[k]);
                if (k == 0) o->has_room_tone = 1; else o->has_speech = 1;
            } catch (const JtError &e) { if (e.code == JT_ERR_CUDA || e.code == JT_ERR_CANCELLED) throw; }
        }
    };

    // ---- Pass 2 on the window: the owned part of the 44.1 kHz output stays on this GPU ----
    GraphRun gd2;
    jt_graph_build(c, spec2, nullptr, total, rate, channels, fmt, 4096, true, true, JT_GRAPH_DRY, nullptr, gd2);
    if (gd2.out.fmt != JT_FMT_S16 || gd2.out.rate != 44100) JT_THROW(JT_ERR_SPEC, "Pass-2 spec must end in the s16/44.1 kHz output stage (processor.go:379-384)");
    const int64_t n44 = gd2.out.n;
    Chunks ck2; ck2.first.resize((size_t)world); ck2.owned.resize((size_t)world);       // ownership of the Pass-2 output
    for (int r = 0; r < world; r++) {
        const int64_t f = (int64_t)((__int128)ck1.first[(size_t)r] * 44100 / rate);
        const bool last = ck1.first[(size_t)r] + ck1.owned[(size_t)r] == total;
        const int64_t e = last ? n44 : (int64_t)((__int128)(ck1.first[(size_t)r] + ck1.owned[(size_t)r]) * 44100 / rate);
        ck2.first[(size_t)r] = f; ck2.owned[(size_t)r] = ck1.owned[(size_t)r] > 0 ? e - f : 0;
    }
    Sig own2; own2.fmt = JT_FMT_S16; own2.rate = 44100; own2.n = ck2.owned[(size_t)rank]; own2.d = jt_dalloc<int16_t>(c, (size_t)own2.n);
    {
        const size_t mark = c->allocs.size();
        std::vector<char> blob; int64_t of = 0, on = 0; int orate = 0, ofmt = 0;
        const bool head_ok = have_head2 && spec2.compare(0, head2.head.size(), head2.head) == 0 &&
                             (spec2.size() == head2.head.size() || spec2[head2.head.size()] == ',');
        graph_chunk_core(c, spec2.c_str(), d_local, n_local, rate, channels, fmt, P.local_first, P.own_first, P.owned, total, 4096,
                         true, nullptr, 0, &own2, &of, &on, &orate, &ofmt, true, blob, head_ok ? &head2 : nullptr);
        if (of != ck2.first[(size_t)rank] || on != ck2.owned[(size_t)rank]) JT_THROW(JT_ERR_INVALID_ARG, "internal: Pass-2 ownership (%lld+%lld vs %lld+%lld)", (long long)of, (long long)on, (long long)ck2.first[(size_t)rank], (long long)ck2.owned[(size_t)rank]);
        JT_CUDA(cudaStreamSynchronize(c->stream));
        jt_release_since(c, mark, nullptr);
        lap(T.pass2_chunk);
        std::vector<std::vector<char>> all = comm.allgather(blob);
        std::vector<const void *> ptrs; for (auto &b : all) if (!b.empty()) ptrs.push_back(b.data());
        rc = jt_graph_merge(spec2.c_str(), total, rate, channels, fmt, 4096, (int)ptrs.size(), ptrs.data(), nullptr, 0, nullptr, nullptr, &R.filtered);
        if (rc) JT_THROW(rc, "Pass-2 merge");
        lap(T.pass2_merge);
    }
    measure_regions(own2, ck2.first[(size_t)rank], ck2, n44, &an->filtered_regions);
    lap(T.regions);

    // ---- Pass 3 / 4 run on their own chunk grid at 44.1 kHz: every rank fetches the halo it lacks ----
    char spec3[1024], spec4[4096];
    rc = jt_build_pass3_spec(R.filtered.input_i, R.filtered.input_tp, tI, tTP, tLRA, spec3, sizeof(spec3), &R);
    if (rc) JT_THROW(rc, "pass-3 spec");
    // Pass 4's grid (whole adeclick hops, resampler periods, ticks) is the coarser one and a multiple of Pass 3's
    jt_process_result plan_probe = R; jt_loudnorm_stats p3_probe; memset(&p3_probe, 0, sizeof(p3_probe));
    p3_probe.input_i = -20; p3_probe.input_tp = -6; p3_probe.input_lra = 5; p3_probe.input_thresh = -30;
    rc = jt_build_pass4_spec(&plan_probe, &p3_probe, tI, tTP, tLRA, 44100, spec4, sizeof(spec4), nullptr, nullptr);
    if (rc) JT_THROW(rc, "pass-4 spec");
    const int64_t U34 = lcm64(chunk_geometry(spec3, 44100).unit, chunk_geometry(spec4, 44100).unit);
    const Chunks ck4 = plan_chunks(n44, U34, world);
    int64_t ctxL = 0, ctxR = 0;
    { const int64_t l = (int64_t)8 * 44100, r = 44100; ctxL = (l + U34 - 1) / U34 * U34; ctxR = (r + U34 - 1) / U34 * U34; }
    auto window_of = [&](int r, int64_t *lo, int64_t *hi) {
        if (ck4.owned[(size_t)r] <= 0) { *lo = *hi = 0; return; }
        *lo = std::max<int64_t>(0, ck4.first[(size_t)r] - ctxL); *hi = std::min(n44, ck4.first[(size_t)r] + ck4.owned[(size_t)r] + ctxR);
    };
    int64_t wlo = 0, whi = 0; window_of(rank, &wlo, &whi);
    Sig win; win.fmt = JT_FMT_S16; win.rate = 44100; win.n = whi - wlo; win.d = jt_dalloc<int16_t>(c, (size_t)std::max<int64_t>(win.n, 1));
    {
        // piece(s -> d) = window(d) ^ owned2(s), for d != s, in rank order of d: every rank can compute every piece's extent
        std::vector<char> send;
        const int64_t mo = ck2.first[(size_t)rank], me = mo + ck2.owned[(size_t)rank];
        for (int d = 0; d < world; d++) {
            if (d == rank) continue;
            int64_t lo, hi; window_of(d, &lo, &hi);
            lo = std::max(lo, mo); hi = std::min(hi, me);
            if (hi > lo) {
                const size_t off = send.size(); send.resize(off + (size_t)(hi - lo) * 2);
                JT_CUDA(cudaMemcpyAsync(send.data() + off, (const char *)own2.d + (size_t)(lo - mo) * 2, (size_t)(hi - lo) * 2, cudaMemcpyDeviceToHost, c->stream));
            }
        }
        JT_CUDA(cudaStreamSynchronize(c->stream));
        T.halo_bytes = (int64_t)send.size();
        std::vector<std::vector<char>> all = comm.allgather(send);
        for (int s = 0; s < world; s++) {
            const int64_t so = ck2.first[(size_t)s], se = so + ck2.owned[(size_t)s];
            if (s == rank) {
                const int64_t lo = std::max(wlo, so), hi = std::min(whi, se);
                if (hi > lo) JT_CUDA(cudaMemcpyAsync((char *)win.d + (size_t)(lo - wlo) * 2, (const char *)own2.d + (size_t)(lo - so) * 2, (size_t)(hi - lo) * 2, cudaMemcpyDeviceToDevice, c->stream));
                continue;
            }
            size_t cur = 0;
            for (int d = 0; d < world; d++) {
                if (d == s) continue;
                int64_t lo, hi; window_of(d, &lo, &hi);
                lo = std::max(lo, so); hi = std::min(hi, se);
                if (hi <= lo) continue;
                if (d == rank) {
                    if (cur + (size_t)(hi - lo) * 2 > all[(size_t)s].size()) JT_THROW(JT_ERR_INVALID_ARG, "internal: halo exchange");
                    JT_CUDA(cudaMemcpyAsync((char *)win.d + (size_t)(lo - wlo) * 2, all[(size_t)s].data() + cur, (size_t)(hi - lo) * 2, cudaMemcpyHostToDevice, c->stream));
                }
                cur += (size_t)(hi - lo) * 2;
            }
        }
        JT_CUDA(cudaStreamSynchronize(c->stream));
        lap(T.halo);
    }
    const bool have4 = ck4.owned[(size_t)rank] > 0;
    // ---- Pass 3 (measure only) ----
    {
        const size_t mark = c->allocs.size();
        std::vector<char> blob;
        if (have4) {
            int64_t of = 0, on = 0; int orate = 0, ofmt = 0;
            graph_chunk_core(c, spec3, win.d, win.n, 44100, 1, JT_FMT_S16, wlo, ck4.first[(size_t)rank], ck4.owned[(size_t)rank], n44, 4096,
                             false, nullptr, 0, nullptr, &of, &on, &orate, &ofmt, true, blob);
            JT_CUDA(cudaStreamSynchronize(c->stream));
        }
        jt_release_since(c, mark, nullptr);
        lap(T.pass3_chunk);
        std::vector<std::vector<char>> all = comm.allgather(blob);
        std::vector<const void *> ptrs; for (auto &b : all) if (!b.empty()) ptrs.push_back(b.data());
        rc = jt_graph_merge(spec3, n44, 44100, 1, JT_FMT_S16, 4096, (int)ptrs.size(), ptrs.data(), nullptr, 0, nullptr, &R.pass3, nullptr);
        if (rc) JT_THROW(rc, "Pass-3 merge");
        lap(T.pass3_merge);
    }
    jt_trace(c, "pass3 finished");
    const double mI = jt_wire("%.2f", R.pass3.input_i);
    if (std::isinf(mI) || std::isnan(mI) || mI < -70.0) JT_THROW(JT_ERR_INVALID_ARG, "cannot normalise silent audio (measured %.1f LUFS)", mI);
    double eff = 0, off = 0;
    rc = jt_build_pass4_spec(&R, &R.pass3, tI, tTP, tLRA, 44100, spec4, sizeof(spec4), &eff, &off);
    if (rc) JT_THROW(rc, "pass-4 spec");
    R.effective_target_i = eff; R.linear_possible = eff == tI;
    // ---- Pass 4 ----
    GraphRun gd4;
    jt_graph_build(c, spec4, nullptr, n44, 44100, 1, JT_FMT_S16, 4096, true, true, JT_GRAPH_DRY, nullptr, gd4);
    const int64_t n_final = gd4.out.n;
    Chunks ck5; ck5.first.resize((size_t)world); ck5.owned.resize((size_t)world);       // ownership of the final output (44.1 kHz in and out)
    for (int r = 0; r < world; r++) {
        const bool last = ck4.owned[(size_t)r] > 0 && ck4.first[(size_t)r] + ck4.owned[(size_t)r] == n44;
        ck5.first[(size_t)r] = ck4.first[(size_t)r]; ck5.owned[(size_t)r] = ck4.owned[(size_t)r] > 0 ? (last ? n_final : ck4.first[(size_t)r] + ck4.owned[(size_t)r]) - ck4.first[(size_t)r] : 0;
    }
    Sig own4; own4.fmt = JT_FMT_S16; own4.rate = 44100; own4.n = ck5.owned[(size_t)rank]; own4.d = jt_dalloc<int16_t>(c, (size_t)std::max<int64_t>(own4.n, 1));
    {
        const size_t mark = c->allocs.size();
        std::vector<char> blob;
        if (have4) {
            int64_t of = 0, on = 0; int orate = 0, ofmt = 0;
            graph_chunk_core(c, spec4, win.d, win.n, 44100, 1, JT_FMT_S16, wlo, ck4.first[(size_t)rank], ck4.owned[(size_t)rank], n44, 4096,
                             true, nullptr, 0, &own4, &of, &on, &orate, &ofmt, true, blob);
            if (of != ck5.first[(size_t)rank] || on != ck5.owned[(size_t)rank]) JT_THROW(JT_ERR_INVALID_ARG, "internal: Pass-4 ownership");
            JT_CUDA(cudaStreamSynchronize(c->stream));
        }
        jt_release_since(c, mark, nullptr);
        lap(T.pass4_chunk);
        std::vector<std::vector<char>> all = comm.allgather(blob);
        std::vector<const void *> ptrs; for (auto &b : all) if (!b.empty()) ptrs.push_back(b.data());
        rc = jt_graph_merge(spec4, n44, 44100, 1, JT_FMT_S16, 4096, (int)ptrs.size(), ptrs.data(), nullptr, 0, nullptr, &R.pass4, &R.final);
        if (rc) JT_THROW(rc, "Pass-4 merge");
        lap(T.pass4_merge);
    }
    measure_regions(own4, ck5.first[(size_t)rank], ck5, n_final, &an->final_regions);
    lap(T.regions);
    // ---- the owned part of the result ----
    R.n_out = n_final;
    if (out_first) *out_first = ck5.first[(size_t)rank];
    if (n_out) *n_out = own4.n;
    if (pcm_out && own4.n > 0) {
        if (own4.n > cap) JT_THROW(JT_ERR_BUFFER, "pcm_out holds %lld samples, this rank owns %lld", (long long)cap, (long long)own4.n);
        JT_CUDA(cudaMemcpyAsync(pcm_out, own4.d, (size_t)own4.n * 2, out_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, c->stream));
    }
    lap(T.download);
    T.exchange = comm.seconds; T.exchange_calls = comm.calls;
    if (res) *res = R;
    if (tm) *tm = T;
}

static int sharded_common(jt_ctx *c, const void *pcm_local, bool on_device, int64_t n_local, int rate, int channels, int fmt,
                          int64_t total, int world, int rank, const jt_filter_config *base, int adaptive,
                          int16_t *pcm_out, int64_t cap, int64_t *out_first, int64_t *n_out,
                          jt_process_result *res, jt_analysis *analysis, jt_shard_timing *tm)
{
    return guarded(c, [&]() {
        if (!pcm_local && n_local > 0) JT_THROW(JT_ERR_INVALID_ARG, "null input");
        if (!jt_valid_fmt(fmt)) JT_THROW(JT_ERR_UNSUPPORTED, "input sample format %d", fmt);
        // the ranks of a sharded stream usually share a node: split its cores among them for the host-side merges (unless the
        // caller has set a number)
        { static bool once = false; if (!once && world > 1 && !getenv("JT_HOST_THREADS")) { jt_set_host_threads((int)std::max(1u, std::thread::hardware_concurrency() / (unsigned)world)); once = true; } }
        const double t0 = ShardComm::host_seconds();
        const void *d_in = on_device ? pcm_local : upload(c, pcm_local, (size_t)n_local * channels * jt_fmt_bytes(fmt));
        JT_CUDA(cudaStreamSynchronize(c->stream));
        const double t_up = ShardComm::host_seconds() - t0;
        process_sharded_device(c, d_in, n_local, rate, channels, fmt, total, world, rank, base, adaptive, pcm_out, on_device, cap, out_first, n_out, res, analysis, tm);
        if (tm) tm->upload = t_up;
    });
}
extern "C" int jt_process_audio_sharded(jt_ctx *c, const void *pcm_local, int64_t n_local, int rate, int channels, int fmt,
                                        int64_t total, int world, int rank, const jt_filter_config *base, int adaptive,
                                        int16_t *pcm_out, int64_t cap, int64_t *out_first, int64_t *n_out,
                                        jt_process_result *res, jt_analysis *analysis, jt_shard_timing *tm)
{
    return sharded_common(c, pcm_local, false, n_local, rate, channels, fmt, total, world, rank, base, adaptive, pcm_out, cap, out_first, n_out, res, analysis, tm);
}
extern "C" int jt_process_audio_sharded_dev(jt_ctx *c, const void *d_local, int64_t n_local, int rate, int channels, int fmt,
                                        int64_t total, int world, int rank, const jt_filter_config *base, int adaptive,
                                        int16_t *d_out, int64_t cap, int64_t *out_first, int64_t *n_out,
                                        jt_process_result *res, jt_analysis *analysis, jt_shard_timing *tm)
{
    return sharded_common(c, d_local, true, n_local, rate, channels, fmt, total, world, rank, base, adaptive, d_out, cap, out_first, n_out, res, analysis, tm);
}

// ---------------------------------------------------------------------------------------
// FLAC container of the chain's output (Encoder: internal/processor/encoder.go:92-101)
// ---------------------------------------------------------------------------------------
extern "C" int64_t jt_flac_max_bytes(int64_t n_samples, int block_size)
{
    if (n_samples < 0 || block_size < 16) return JT_ERR_INVALID_ARG;
    return jt_flac_bound(n_samples, block_size);
}
static int flac_common(jt_ctx *c, const int16_t *pcm, bool on_device, int64_t n, int rate, int block_size, void *out, int64_t cap, int64_t *n_bytes)
{
    return guarded(c, [&]() {
        if ((!pcm && n > 0) || !out || n < 0) JT_THROW(JT_ERR_INVALID_ARG, "null argument");
        const int16_t *d_in = on_device ? pcm : (const int16_t *)upload(c, pcm, (size_t)n * sizeof(int16_t));
        int64_t total = 0;
        void *d = jt_flac_encode_device(c, d_in, n, rate, block_size, &total);
        if (n_bytes) *n_bytes = total;
        if (total > cap) JT_THROW(JT_ERR_BUFFER, "out holds %lld bytes, the stream has %lld", (long long)cap, (long long)total);
        if (on_device) JT_CUDA(cudaMemcpyAsync(out, d, (size_t)total, cudaMemcpyDeviceToDevice, c->stream));
        else download(c, out, d, (size_t)total);
    });
}
extern "C" int jt_flac_encode(jt_ctx *c, const int16_t *pcm, int64_t n, int rate, int block_size, void *out, int64_t cap, int64_t *n_bytes)
{
    return flac_common(c, pcm, false, n, rate, block_size, out, cap, n_bytes);
}
extern "C" int jt_flac_encode_dev(jt_ctx *c, const int16_t *d_pcm, int64_t n, int rate, int block_size, void *d_out, int64_t cap, int64_t *n_bytes)
{
    return flac_common(c, d_pcm, true, n, rate, block_size, d_out, cap, n_bytes);
}

// ---------------------------------------------------------------------------------------
// Input-side containers (audio.Reader: internal/audio/reader.go:29-188): FLAC and WAV file images -> interleaved PCM
// ---------------------------------------------------------------------------------------
extern "C" int jt_flac_stream_info(const void *bytes, int64_t n_bytes, int *sample_fmt, int *sample_rate, int *channels, int *bits_per_sample,
                                   int64_t *n_frames, int64_t *audio_offset)
{
    if (!bytes || n_bytes < 0) return JT_ERR_INVALID_ARG;
    try { jt_flac_info_host(bytes, n_bytes, sample_fmt, sample_rate, channels, bits_per_sample, n_frames, audio_offset); }
    catch (const JtError &e) { return e.code; }
    return JT_OK;
}
static int flac_decode_common(jt_ctx *c, const void *bytes, bool on_device, int64_t n_bytes, void *out, int64_t cap_frames, int64_t *n_frames,
                              int *sample_fmt, int *sample_rate, int *channels)
{
    return guarded(c, [&]() {
        if (!bytes || !out || n_bytes < 0 || cap_frames < 0) JT_THROW(JT_ERR_INVALID_ARG, "null argument");
        const uint8_t *d_in = on_device ? (const uint8_t *)bytes : (const uint8_t *)upload(c, bytes, (size_t)n_bytes);
        int fmt = 0, rate = 0, ch = 0; int64_t frames = 0;
        void *d = jt_flac_decode_device(c, d_in, n_bytes, &fmt, &rate, &ch, &frames);
        if (n_frames) *n_frames = frames; if (sample_fmt) *sample_fmt = fmt; if (sample_rate) *sample_rate = rate; if (channels) *channels = ch;
        if (frames > cap_frames) JT_THROW(JT_ERR_BUFFER, "out holds %lld frames, the stream has %lld", (long long)cap_frames, (long long)frames);
        const size_t bytes_out = (size_t)frames * ch * jt_fmt_bytes(fmt);
        if (on_device) JT_CUDA(cudaMemcpyAsync(out, d, bytes_out, cudaMemcpyDeviceToDevice, c->stream));
        else download(c, out, d, bytes_out);
    });
}
extern "C" int jt_flac_decode(jt_ctx *c, const void *bytes, int64_t n_bytes, void *pcm_out, int64_t cap_frames, int64_t *n_frames,
                              int *sample_fmt, int *sample_rate, int *channels)
{
    return flac_decode_common(c, bytes, false, n_bytes, pcm_out, cap_frames, n_frames, sample_fmt, sample_rate, channels);
}
extern "C" int jt_flac_decode_dev(jt_ctx *c, const void *d_bytes, int64_t n_bytes, void *d_pcm_out, int64_t cap_frames, int64_t *n_frames,
                                  int *sample_fmt, int *sample_rate, int *channels)
{
    return flac_decode_common(c, d_bytes, true, n_bytes, d_pcm_out, cap_frames, n_frames, sample_fmt, sample_rate, channels);
}

// WAV: what libavformat's wav demuxer + libavcodec's pcm_* decoders hand the reference: s16 / s32 (24-bit samples shifted up by 8) /
// flt / dbl, interleaved.  The header is walked on the host (jt_wav_walk), the samples are copied / unpacked on the device.
int jt_wav_walk(const void *bytes, int64_t n_readable, int64_t n_total, int *sample_fmt, int *sample_rate, int *channels, int *bits_per_sample,
                int64_t *data_offset, int64_t *n_frames);                       // jt_wav.cu
static int wav_decode_common(jt_ctx *c, const void *bytes, bool on_device, int64_t n_bytes, void *out, bool out_on_device, int64_t cap_frames,
                             int64_t *n_frames, int *sample_fmt, int *sample_rate, int *channels)
{
    return guarded(c, [&]() {
        if (!bytes || !out || n_bytes < 0 || cap_frames < 0) JT_THROW(JT_ERR_INVALID_ARG, "null argument");
        // the header chunks sit in the first bytes of the file; 1 MB covers any LIST / bext / iXML block in front of "data"
        std::vector<uint8_t> head;
        const uint8_t *h = (const uint8_t *)bytes;
        int64_t n_head = n_bytes;
        if (on_device) {
            n_head = std::min<int64_t>(n_bytes, 1 << 20);
            head.resize((size_t)n_head);
            JT_CUDA(cudaMemcpyAsync(head.data(), bytes, (size_t)n_head, cudaMemcpyDeviceToHost, c->stream));
            JT_CUDA(cudaStreamSynchronize(c->stream));
            h = head.data();
        }
        int fmt = 0, rate = 0, ch = 0, bits = 0; int64_t off = 0, frames = 0;
        const int rc = jt_wav_walk(h, n_head, n_bytes, &fmt, &rate, &ch, &bits, &off, &frames);
        if (rc != JT_OK) JT_THROW(rc, "WAV header not understood (%s)", rc == JT_ERR_UNSUPPORTED ? "unsupported sample format, RF64, or no data chunk in the first MB" : "malformed");
        if (n_frames) *n_frames = frames; if (sample_fmt) *sample_fmt = fmt; if (sample_rate) *sample_rate = rate; if (channels) *channels = ch;
        if (frames > cap_frames) JT_THROW(JT_ERR_BUFFER, "out holds %lld frames, the file has %lld", (long long)cap_frames, (long long)frames);
        const int64_t n_samples = frames * ch;
        const size_t in_bytes = (size_t)n_samples * (bits / 8), out_bytes = (size_t)n_samples * jt_fmt_bytes(fmt);
        const uint8_t *d_in = on_device ? (const uint8_t *)bytes + off : (const uint8_t *)upload(c, (const uint8_t *)bytes + off, in_bytes);
        void *d_res = out_on_device ? out : jt_dalloc_bytes(c, out_bytes + 16);
        if (bits == 24) jt_unpack_s24(c, d_in, n_samples, (int32_t *)d_res);
        else JT_CUDA(cudaMemcpyAsync(d_res, d_in, out_bytes, cudaMemcpyDeviceToDevice, c->stream));
        if (!out_on_device) download(c, out, d_res, out_bytes);
    });
}
extern "C" int jt_wav_decode(jt_ctx *c, const void *bytes, int64_t n_bytes, void *pcm_out, int64_t cap_frames, int64_t *n_frames,
                             int *sample_fmt, int *sample_rate, int *channels)
{
    return wav_decode_common(c, bytes, false, n_bytes, pcm_out, false, cap_frames, n_frames, sample_fmt, sample_rate, channels);
}
extern "C" int jt_wav_decode_dev(jt_ctx *c, const void *d_bytes, int64_t n_bytes, void *d_pcm_out, int64_t cap_frames, int64_t *n_frames,
                                 int *sample_fmt, int *sample_rate, int *channels)
{
    return wav_decode_common(c, d_bytes, true, n_bytes, d_pcm_out, true, cap_frames, n_frames, sample_fmt, sample_rate, channels);
}
