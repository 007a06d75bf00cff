// Filter-spec parser and whole-stream graph executor (the S4 seam:
// setupFilterGraph/runFilterGraph, internal/processor/frame_processor.go:64-216).
#pragma once
#include "jt_internal.h"
#include <string>
#include <vector>
#include <map>

struct FilterNode {
    std::string name;
    std::vector<std::pair<std::string, std::string>> opts;   // in order; positional args have empty key
    const std::string *get(const char *k1, const char *k2 = nullptr) const;
    double num(const char *k1, const char *k2, double dflt) const;
    std::string str(const char *k1, const char *k2, const char *dflt) const;
    bool flag(const char *k1, const char *k2, bool dflt) const;
};
std::vector<FilterNode> jt_parse_spec(const std::string &spec);
// af_loudnorm.c init(): linear mode only when all four measured_* were given and the projected peak / LRA fit the targets
bool jt_loudnorm_linear_mode(const FilterNode &f);

// One frame on a link, with the metadata it inherited (indices into the producers' outputs).
struct FrameRef {
    int64_t start = 0;       // first sample on this link
    int32_t nb = 0;
    int64_t ready = 0;       // source samples that must have been pushed before this frame can be pulled
    int64_t astats_pos = -1; // astats cumulative over [0, astats_pos) of the astats input signal
    int32_t hop = -1;        // aspectralstats hop index
    int32_t tick = -1;       // ebur128 100 ms tick index
};

// A pending uniform re-framing of a link (ff_inlink_consume_samples with min = max = F on a stream of n samples) and what the
// filter that asked for it stamps on the new frames.  Consecutive levels are collapsed in ONE pass over the final frames
// (frames_collapse): a Pass-2 graph goes source 4096 -> anlmdn 577 -> afftdn 600 -> aspectralstats 1024 -> ebur128 4800, and
// materialising the 300 000 + 288 000 + 169 000 intermediate frames of an hour of audio was 10 ms of host time per pass.
enum { JT_LVL_ASTATS = 1, JT_LVL_HOP = 2, JT_LVL_TICK = 4 };
struct FrameLvl { int64_t F = 1, n = 0; unsigned tag = 0; };

struct GraphResult {
    Sig out;                                  // sink signal (device)
    std::vector<jt_frame_meta> meta;          // one per sink frame
    std::vector<int64_t> meta_ready;          // FrameRef::ready per sink frame (INT64_MAX = only at flush)
    jt_loudnorm_stats ln;
};

// A graph whose kernels and device->host copies have been enqueued but whose host part (windowing / gating of the
// tick values, statistics, sink-frame metadata) has not run yet.  jt_graph_enqueue never waits for the device;
// jt_graph_finish does.  In between the caller can enqueue the next graph so the GPU stays busy while the host
// assembles this one's metadata (jt_graph_r128_early gives the loudness values the next pass is planned from).
struct GraphRun {
    std::vector<FrameRef> frames;
    Sig out; int out_fmt = 0;
    cudaEvent_t out_ready = nullptr;      // recorded once the sink audio is complete (the analysis kernels come after it)
    bool want_meta = false;
    bool has_astats = false, has_spec = false, has_r128 = false, astats_overall_only = false;
    Sig astats_sig, spec_sig; int spec_win = 2048;
    R128Pending r128p; SpectralPending specp; AstatsPending astp;
    long last_astats_frame = -1;
    bool astats_on_low = false;           // the executor put astats on the context's low-priority stream (jt_ctx::defer_astats)
    bool astats_later = false;            // jt_graph_finish leaves astats' values out: the caller collects them (jt_astats_finish on astp) when it needs them
    R128Result r128; bool r128_done = false;
    // loudnorm
    bool has_ln = false, ln_linear = false, ln_dual = false; double ln_I = 0;
    LoudnormPending ln_in, ln_out, ln_tail;
    // dynamic mode (af_loudnorm.c outside its linear-mode preconditions): both links at 192 kHz; the flush frame --
    // the stream's last 2.9 s -- is metered a second time on the input side (ln_in_extra samples; ln_tail holds
    // their per-100 ms values from global tick ln_tail_first on); ln_type is what uninit() prints (0 "linear": a
    // stream shorter than 3 s falls back to a single gain)
    bool ln_dynamic = false, ln_has_out = false; int ln_type = 0; int64_t ln_in_extra = 0, ln_tail_first = 0;
    // the signals the measuring filters see (DRY: sizes / formats only; CHUNK: device signals of the local window)
    Sig r128_sig, ln_in_sig, ln_out_sig; bool r128_dual = false, r128_tp = false;
    int exchanges = 0;          // cross-chunk exchange steps the graph went through (CHUNK)
};

enum { JT_GRAPH_NORMAL = 0, JT_GRAPH_DRY = 1, JT_GRAPH_CHUNK = 2 };
// The executor's state after the first n_nodes filters of a spec: lets the spec-independent head of Pass 2 (downmix, both
// biquads, anlmdn -- filters.go:58-68) run while the host still derives the adaptive tail of the spec from Pass 1.
struct GraphResume { int n_nodes = 0; std::string head; Sig cur; int link_fmt = 0; std::vector<FrameRef> frames; std::vector<FrameLvl> chain; AfftdnFwd fwd; };
// predicted_spec: the whole spec the head was cut from; when afftdn follows the head, its parameter-independent forward transforms
// run here too (jt_afftdn_forward) and ride along in `out`
void jt_graph_head(jt_ctx *c, const std::string &head_spec, const void *d_in, int64_t n_frames, int rate, int channels,
                   int fmt, int frame_size, GraphResume &out, const std::string *predicted_spec = nullptr);
// Where a local window sits in its stream (input-link sample indices) and how carries cross chunk boundaries
struct GraphChunk {
    int64_t local_first = 0, own_first = 0, owned = 0, total = 0; int rate = 0; bool last = false;
    jt_exchange_fn exchange = nullptr; void *exchange_user = nullptr; int n_ranks = 1;
    // position on a link running at `link_rate` of input-link position `pos` (exact: boundaries sit on the chunk unit)
    int64_t link_pos(int64_t pos, int link_rate) const { return (int64_t)((__int128)pos * link_rate / rate); }
};
void jt_graph_build(jt_ctx *c, const std::string &spec, const void *d_in, int64_t n_frames, int rate, int channels,
                    int fmt, int frame_size, bool want_pcm, bool want_meta, int mode, const GraphChunk *chunk, GraphRun &g,
                    const GraphResume *resume = nullptr, GraphResume *capture = nullptr);
void jt_assemble_records(const std::vector<FrameRef> &frames, bool has_r128, const R128Result &r128, bool has_spec,
                         const std::vector<float> &spec_rows, int64_t spec_hops, GraphResult &res);
void jt_graph_enqueue(jt_ctx *c, const std::string &spec, const void *d_in, int64_t n_frames, int rate, int channels,
                      int fmt, int frame_size, bool want_pcm, bool want_meta, GraphRun &g, const GraphResume *resume = nullptr);
const R128Result &jt_graph_r128_early(jt_ctx *c, GraphRun &g);     // waits for the ebur128 values only
void jt_graph_finish(jt_ctx *c, GraphRun &g, GraphResult &res);
void jt_graph_finish_acc(jt_ctx *c, GraphRun &g, jt_measurements *acc, jt_loudnorm_stats *ln);
void jt_accumulate_frames(const std::vector<FrameRef> &frames, bool has_r128, const R128Result &r128, bool has_spec,
                          const std::vector<float> &spec_rows, int64_t spec_hops, jt_measurements *out);

void jt_pass1_records(jt_ctx *c, int64_t n_frames, int rate, int frame_size, const R128Result &r128,
                      const std::vector<float> &spec_rows, int64_t spec_hops, const AstatsResult *astats, GraphResult &res);

// d_in: device pointer to interleaved input.  want_pcm=false lets measure-only graphs skip
// work whose only product is discarded audio (loudnorm dynamic mode in Pass 3).
void jt_graph_run(jt_ctx *c, const std::string &spec, const void *d_in, int64_t n_frames, int rate, int channels,
                  int fmt, int frame_size, bool want_pcm, bool want_meta, GraphResult &res);

double jt_wire(const char *fmt, double v);    // value as the Go side parses it back from FFmpeg's printf
unsigned jt_host_threads();                   // host threads a call may use for per-frame metadata work (jt_set_host_threads)
