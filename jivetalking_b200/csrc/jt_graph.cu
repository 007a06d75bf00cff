// Filter-spec parser + whole-stream graph executor + sink-frame metadata assembly.
// Replaces setupFilterGraph / runFilterGraph (internal/processor/frame_processor.go:64-216):
// the spec strings are the reference's own (filters.go:968-989, normalise.go:257-264,
// 1231-1334, analyser_bands.go:33, analyser_output.go:18); instead of pumping 4096-sample
// frames through libavfilter, each filter runs as whole-stream kernels, and the frame /
// metadata cadence libavfilter would have produced is reconstructed with integer arithmetic
// (which sink frame inherits which astats / aspectralstats / ebur128 stamp, and after which
// pushed input frame it becomes pullable).
#include "jt_graph.h"
#include <atomic>
#include <cstdlib>
#include <algorithm>
#include <array>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <climits>
#include <thread>

// ---------------------------------------------------------------------------------------
// parsing
// ---------------------------------------------------------------------------------------
const std::string *FilterNode::get(const char *k1, const char *k2) const
{
    for (auto &kv : opts) if (kv.first == k1 || (k2 && kv.first == k2)) return &kv.second;
    return nullptr;
}
double FilterNode::num(const char *k1, const char *k2, double dflt) const
{
    const std::string *s = get(k1, k2);
    if (!s) return dflt;
    char *end = nullptr;
    double v = strtod(s->c_str(), &end);
    if (end == s->c_str()) JT_THROW(JT_ERR_SPEC, "filter %s: option %s=%s is not a number", name.c_str(), k1, s->c_str());
    return v;
}
std::string FilterNode::str(const char *k1, const char *k2, const char *dflt) const
{
    const std::string *s = get(k1, k2);
    return s ? *s : std::string(dflt);
}
bool FilterNode::flag(const char *k1, const char *k2, bool dflt) const
{
    const std::string *s = get(k1, k2);
    if (!s) return dflt;
    if (*s == "1" || *s == "true" || *s == "yes" || *s == "on") return true;
    if (*s == "0" || *s == "false" || *s == "no" || *s == "off") return false;
    JT_THROW(JT_ERR_SPEC, "filter %s: option %s=%s is not a boolean", name.c_str(), k1, s->c_str());
}

// split on `sep` at top level, honouring backslash escapes and single quotes
static std::vector<std::string> split_escaped(const std::string &s, char sep, bool unescape)
{
    std::vector<std::string> out; std::string cur; bool quote = false;
    for (size_t i = 0; i < s.size(); i++) {
        char ch = s[i];
        if (ch == '\\' && i + 1 < s.size()) { if (!unescape) cur += ch; cur += s[++i]; continue; }
        if (ch == '\'') { quote = !quote; if (!unescape) cur += ch; continue; }
        if (ch == sep && !quote) { out.push_back(cur); cur.clear(); continue; }
        cur += ch;
    }
    out.push_back(cur);
    return out;
}

std::vector<FilterNode> jt_parse_spec(const std::string &spec)
{
    std::vector<FilterNode> nodes;
    if (spec.empty()) return nodes;
    if (spec.find(';') != std::string::npos || spec.find('[') != std::string::npos)
        JT_THROW(JT_ERR_SPEC, "only linear filter chains are supported");
    for (const std::string &f : split_escaped(spec, ',', false)) {
        if (f.empty()) JT_THROW(JT_ERR_SPEC, "empty filter in spec");
        FilterNode n;
        size_t eq = f.find('=');
        n.name = f.substr(0, eq);
        for (char ch : n.name) if (!(isalnum((unsigned char)ch) || ch == '_')) JT_THROW(JT_ERR_SPEC, "bad filter name '%s'", n.name.c_str());
        if (eq != std::string::npos) {
            for (const std::string &o : split_escaped(f.substr(eq + 1), ':', true)) {
                size_t e2 = o.find('=');
                if (e2 == std::string::npos) n.opts.push_back({"", o});
                else n.opts.push_back({o.substr(0, e2), o.substr(e2 + 1)});
            }
        }
        nodes.push_back(n);
    }
    return nodes;
}

bool jt_loudnorm_linear_mode(const FilterNode &f)
{
    if (!f.flag("linear", "", true)) return false;
    const double I = f.num("I", "i", -24), TP = f.num("TP", "tp", -2), LRA = f.num("LRA", "lra", 7);
    const double mI = f.num("measured_I", "measured_i", 0), mTP = f.num("measured_TP", "measured_tp", 99);
    const double mLRA = f.num("measured_LRA", "measured_lra", 0), mTh = f.num("measured_thresh", "", -70);
    const double off = I - mI, off_tp = mTP + off;
    return mTP != 99 && mTh != -70 && mLRA != 0 && mI != 0 && off_tp <= TP && mLRA <= LRA;
}

static double wire_slow(const char *fmt, double v)
{
    char b[512];          // "%f" of DBL_MAX (astats Min_difference on a 1-sample stream) is 316 characters
    snprintf(b, sizeof(b), fmt, v);
    return strtod(b, nullptr);
}

// snprintf + strtod round trip without the text: scale to an integer, round, scale back.  Powers of ten
// up to 1e22 are exact doubles and IEEE division / multiplication round once, so q / 10^k is the double
// nearest to the printed decimal -- exactly what strtod returns.  Values whose scaled fraction sits within
// 1e-6 of a rounding tie (where printf's exact-decimal rounding could differ) take the text path.
static const double kPow10[23] = {1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11, 1e12, 1e13, 1e14, 1e15,
                                  1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};
static bool wire_fixed(double v, int decimals, double *out)
{
    if (!std::isfinite(v) || fabs(v) > 1e9) return false;
    const double t = v * kPow10[decimals], r = nearbyint(t);
    if (fabs(fabs(t - r) - 0.5) < 1e-6) return false;
    *out = r / kPow10[decimals];
    return true;
}
static bool wire_g6(double v, double *out)
{
    if (v == 0.0) { *out = v; return true; }
    if (!std::isfinite(v)) return false;
    const double a = fabs(v);
    if (a < 1e-15 || a > 1e15) return false;
    int e = (int)floor(log10(a));
    for (int pass = 0; pass < 2; pass++) {
        const int k = 5 - e;                                  // scale so that 6 significant digits are integral
        const double t = k >= 0 ? a * kPow10[k] : a / kPow10[-k];
        if (t < 99999.5 - 1e-3) { e--; continue; }
        if (t >= 999999.5 - 1e-3) { if (t < 999999.5 + 1e-3) return false; e++; continue; }
        const double r = nearbyint(t);
        if (fabs(fabs(t - r) - 0.5) < 1e-6) return false;
        const double m = k >= 0 ? r / kPow10[k] : r * kPow10[-k];
        *out = v < 0 ? -m : m;
        return true;
    }
    return false;
}

// Host threads one call may use for its per-frame metadata work (the "%g" rounding of ~13 values per sink frame).  Default: up to 8;
// a process that shares its node with other ranks of a sharded stream divides the cores among them (jt_set_host_threads, or
// JT_HOST_THREADS) -- eight ranks x eight threads on sixteen cores was most of pass1_merge in profiles/bench_r2i_n8.json.
static std::atomic<int> g_host_threads{0};
extern "C" void jt_set_host_threads(int n) { g_host_threads.store(n < 0 ? 0 : n); }
unsigned jt_host_threads()
{
    int n = g_host_threads.load();
    if (n <= 0) { static const char *e = getenv("JT_HOST_THREADS"); if (e) n = atoi(e); }
    const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    if (n <= 0) n = 8;
    return std::max(1u, std::min((unsigned)n, hw));
}

double jt_wire(const char *fmt, double v)
{
    double o;
    if (fmt[0] == '%' && fmt[1] == '.' && fmt[2] == '3' && fmt[3] == 'f' && !fmt[4]) { if (wire_fixed(v, 3, &o)) return o; }
    else if (fmt[0] == '%' && fmt[1] == '.' && fmt[2] == '2' && fmt[3] == 'f' && !fmt[4]) { if (wire_fixed(v, 2, &o)) return o; }
    else if (fmt[0] == '%' && fmt[1] == 'f' && !fmt[2]) { if (wire_fixed(v, 6, &o)) return o; }
    else if (fmt[0] == '%' && fmt[1] == 'g' && !fmt[2]) { if (wire_g6(v, &o)) return o; }
    return wire_slow(fmt, v);
}
double jt_wire_text(const char *fmt, double v) { return wire_slow(fmt, v); }
// test hook (tests/test_wire.py): fast path vs the literal snprintf/strtod round trip
extern "C" double jt_debug_wire(const char *fmt, double v, int text) { return text ? wire_slow(fmt, v) : jt_wire(fmt, v); }

// ---------------------------------------------------------------------------------------
// frame bookkeeping
// ---------------------------------------------------------------------------------------
static std::vector<FrameRef> source_frames(int64_t n, int F)
{
    std::vector<FrameRef> v;
    v.reserve((size_t)((n + F - 1) / std::max(F, 1)));
    for (int64_t s = 0; s < n; s += F) {
        FrameRef f; f.start = s; f.nb = (int32_t)std::min<int64_t>(F, n - s); f.ready = s + f.nb;
        v.push_back(f);
    }
    return v;
}

// libavfilter's ff_inlink_consume_samples(min=max=F): new frames of F samples (last partial);
// properties come from the queued frame holding the first sample (take_samples:
// av_frame_copy_props(buf, frame0)); a frame can be built once its last sample has arrived.
static std::vector<FrameRef> reframe(const std::vector<FrameRef> &old, int64_t n, int F)
{
    std::vector<FrameRef> v;
    v.reserve((size_t)((n + F - 1) / std::max(F, 1)));
    size_t a = 0, b = 0;
    for (int64_t s = 0; s < n; s += F) {
        FrameRef f; f.start = s; f.nb = (int32_t)std::min<int64_t>(F, n - s);
        while (a + 1 < old.size() && old[a + 1].start <= s) a++;
        const int64_t last = s + f.nb - 1;
        if (b < a) b = a;
        while (b + 1 < old.size() && old[b + 1].start <= last) b++;
        if (!old.empty()) {
            f.astats_pos = old[a].astats_pos; f.hop = old[a].hop; f.tick = old[a].tick;
            f.ready = old[b].ready;
        }
        v.push_back(f);
    }
    return v;
}

// `chain` applied to `base`, level by level, without the intermediate lists: frame k of the last level starts at k F; walking
// the levels downwards, the frame holding a frame's FIRST sample hands down its stamps (the stamp of the latest level wins, as
// each reframe copies and each filter then overwrites), the frame holding its LAST sample decides when it can be built.
// Equal to reframe() applied once per level followed by the filters' loops over the new frames.
static std::vector<FrameRef> frames_collapse(const std::vector<FrameRef> &base, const std::vector<FrameLvl> &chain)
{
    if (chain.empty()) return base;
    const FrameLvl &top = chain.back();
    std::vector<FrameRef> v;
    const int64_t cnt = top.F > 0 ? (top.n + top.F - 1) / top.F : 0;
    v.resize((size_t)std::max<int64_t>(cnt, 0));
    const int L = (int)chain.size();
    std::vector<int64_t> lvl_cnt((size_t)L);
    for (int i = 0; i < L; i++) lvl_cnt[i] = (chain[i].n + chain[i].F - 1) / chain[i].F;
    size_t a = 0, b = 0;
    for (int64_t k = 0; k < cnt; k++) {
        FrameRef f; f.start = k * top.F; f.nb = (int32_t)std::min<int64_t>(top.F, top.n - f.start);
        int64_t first = f.start, last = f.start + f.nb - 1;
        bool have_as = false, have_hop = false, have_tick = false, dead = false;
        for (int i = L - 1; i >= 0; i--) {
            const FrameLvl &lv = chain[i];
            if (lvl_cnt[i] <= 0) { dead = true; break; }               // an empty list below: nothing to inherit (reframe: old.empty())
            const int64_t ia = std::min(first / lv.F, lvl_cnt[i] - 1), ib = std::min(last / lv.F, lvl_cnt[i] - 1);
            const int64_t sa = ia * lv.F, na = std::min<int64_t>(lv.F, lv.n - sa);
            if ((lv.tag & JT_LVL_ASTATS) && !have_as) { f.astats_pos = sa + na; have_as = true; }
            if ((lv.tag & JT_LVL_HOP) && !have_hop) { f.hop = (int32_t)ia; have_hop = true; }
            if ((lv.tag & JT_LVL_TICK) && !have_tick && na == lv.F) { f.tick = (int32_t)ia; have_tick = true; }
            first = sa; last = std::min((ib + 1) * lv.F, lv.n) - 1;
        }
        if (!dead && !base.empty()) {
            while (a + 1 < base.size() && base[a + 1].start <= first) a++;
            if (b < a) b = a;
            while (b + 1 < base.size() && base[b + 1].start <= last) b++;
            if (!have_as) f.astats_pos = base[a].astats_pos;
            if (!have_hop) f.hop = base[a].hop;
            if (!have_tick) f.tick = base[a].tick;
            f.ready = base[b].ready;
        }
        v[(size_t)k] = f;
    }
    return v;
}

// ---------------------------------------------------------------------------------------
// executor
// ---------------------------------------------------------------------------------------
namespace {
struct Exec {
    jt_ctx *c; Sig cur; int link_fmt; std::vector<FrameRef> frames;
    std::vector<FrameLvl> chain;          // pending uniform re-framings of `frames` (collapsed by flush())
    void ureframe(int64_t F) { FrameLvl l; l.F = std::max<int64_t>(F, 1); l.n = cur.n; chain.push_back(l); }
    void stamp(unsigned tag)              // the filter's loop over the frames it has just asked for
    {
        if (!chain.empty()) { chain.back().tag |= tag; return; }
        if (tag & JT_LVL_ASTATS) for (FrameRef &fr : frames) fr.astats_pos = fr.start + fr.nb;
        if (tag & JT_LVL_HOP) for (size_t j = 0; j < frames.size(); j++) frames[j].hop = (int32_t)j;
    }
    void flush() { if (!chain.empty()) { std::vector<FrameRef> v = frames_collapse(frames, chain); frames.swap(v); chain.clear(); } }
    // analysis products
    bool has_astats = false, has_spec = false, has_r128 = false, astats_overall_only = false;
    Sig astats_sig, spec_sig; int spec_win = 2048;
    bool dry = false;          // JT_GRAPH_DRY: sizes, formats and frame cadence only, no device work
    void storage(int fmt) { if (dry) cur.fmt = fmt; else cur = jt_convert(c, cur, fmt); link_fmt = fmt; }
    void materialise() { if (cur.fmt != link_fmt) { if (dry) cur.fmt = link_fmt; else cur = jt_convert(c, cur, link_fmt); } }
};
}

static int swr_internal_fmt(int in_fmt, int out_fmt)
{   // libswresample/swresample.c swr_init(): int_sample_fmt when a rate conversion is involved
    const size_t bi = jt_fmt_bytes(in_fmt), bo = jt_fmt_bytes(out_fmt);
    if (bi <= 2 && bo <= 2) return JT_FMT_S16;
    if (bi <= 4) return JT_FMT_FLT;
    return JT_FMT_DBL;
}

static void do_resample(Exec &E, int out_rate, int out_fmt /* 0 = keep link format */, bool keep_work_fmt = false)
{
    jt_ctx *c = E.c;
    const int in_link = E.link_fmt;
    if (!out_fmt) out_fmt = in_link;
    if (out_rate == E.cur.rate) {          // format conversion only
        if (out_fmt != in_link) { E.materialise(); E.storage(out_fmt); }
        return;
    }
    const int work = swr_internal_fmt(in_link, out_fmt);
    if (work == JT_FMT_S16) JT_THROW(JT_ERR_UNSUPPORTED, "s16-internal resampling (s16 -> s16 rate change)");
    SwrPlan p = jt_swr_plan(E.cur.rate, out_rate);
    const int64_t n_in = E.cur.n;
    // swr converts to its internal format first, then resamples
    Sig src = E.cur;
    if (work == JT_FMT_FLT && src.fmt == JT_FMT_DBL) { if (E.dry) src.fmt = JT_FMT_FLT; else src = jt_convert(c, src, JT_FMT_FLT); }
    Sig r;
    if (E.dry) { r = src; r.d = nullptr; r.rate = out_rate; r.n = p.out_count_flush(src.n); r.fmt = (work == JT_FMT_FLT && out_fmt == JT_FMT_S16) ? JT_FMT_S16 : work; }
    else r = jt_swr_resample(c, src, p, work, true, out_fmt);
    // frames: one output frame per input frame (aresample filter_frame), then the EOF flush frame
    E.flush();
    std::vector<FrameRef> nf; int64_t done = 0;
    for (const FrameRef &f : E.frames) {
        const int64_t cnt = std::min<int64_t>(p.out_count(f.start + f.nb), r.n);
        if (cnt > done) { FrameRef o = f; o.start = done; o.nb = (int32_t)(cnt - done); nf.push_back(o); done = cnt; }
    }
    if (r.n > done) { FrameRef o; o.start = done; o.nb = (int32_t)(r.n - done); o.ready = INT64_MAX; nf.push_back(o); }
    (void)n_in;
    E.frames.swap(nf);
    E.cur = r; E.link_fmt = r.fmt;            // r.fmt == out_fmt when the kernel converted on store
    // keep_work_fmt: the consumer widens on load (f32 -> f64 is exact), so the converted copy is never materialised
    if (out_fmt != r.fmt && !keep_work_fmt) E.storage(out_fmt);
}

static std::vector<double> parse_bn(const std::string &s)
{
    std::vector<double> v;
    for (const std::string &t : split_escaped(s, '|', true)) {
        if (t.empty()) continue;
        size_t p = 0; std::string u = t;
        while ((p = u.find(' ')) != std::string::npos) u.erase(p, 1);
        v.push_back(strtod(u.c_str(), nullptr));
    }
    return v;
}

void jt_graph_run(jt_ctx *c, const std::string &spec, const void *d_in, int64_t n_frames, int rate, int channels,
                  int fmt, int frame_size, bool want_pcm, bool want_meta, GraphResult &res)
{
    GraphRun g;
    jt_graph_enqueue(c, spec, d_in, n_frames, rate, channels, fmt, frame_size, want_pcm, want_meta, g);
    jt_graph_finish(c, g, res);
}

void jt_graph_enqueue(jt_ctx *c, const std::string &spec, const void *d_in, int64_t n_frames, int rate, int channels,
                      int fmt, int frame_size, bool want_pcm, bool want_meta, GraphRun &g, const GraphResume *resume)
{
    jt_graph_build(c, spec, d_in, n_frames, rate, channels, fmt, frame_size, want_pcm, want_meta, JT_GRAPH_NORMAL, nullptr, g, resume);
}

// run the filters of `head_spec` and hand back the executor's state behind them (no analysis, no sink)
static AfftdnParams parse_afftdn(const FilterNode &f)
{
    AfftdnParams p;
    p.nr = f.num("nr", "noise_reduction", 12); p.nf = f.num("nf", "noise_floor", -50); p.rf = f.num("rf", "residual_floor", -38);
    p.ad = f.num("ad", "adaptivity", 0.5); p.fo = f.num("fo", "floor_offset", 1.0); p.bm = f.num("bm", "band_multiplier", 1.25);
    p.gs = (int)f.num("gs", "gain_smooth", 0);
    p.tn = f.flag("tn", "track_noise", false);
    if (f.flag("tr", "track_residual", false)) JT_THROW(JT_ERR_UNSUPPORTED, "afftdn track_residual");
    const std::string nt = f.str("nt", "noise_type", "w");
    if (nt == "w" || nt == "white") p.nt = 0; else if (nt == "v" || nt == "vinyl") p.nt = 1; else if (nt == "s" || nt == "shellac") p.nt = 2;
    else if (nt == "c" || nt == "custom") p.nt = 3; else JT_THROW(JT_ERR_SPEC, "afftdn nt=%s", nt.c_str());
    if (const std::string *bn = f.get("bn", "band_noise")) {
        std::vector<double> v = parse_bn(*bn); p.has_bn = true;
        for (size_t i = 0; i < 15 && i < v.size(); i++) p.bn[i] = v[i];
    }
    const std::string om = f.str("om", "output_mode", "o");
    if (om != "o" && om != "output") JT_THROW(JT_ERR_UNSUPPORTED, "afftdn output mode %s", om.c_str());
    return p;
}

void jt_graph_head(jt_ctx *c, const std::string &head_spec, const void *d_in, int64_t n_frames, int rate, int channels,
                   int fmt, int frame_size, GraphResume &out, const std::string *predicted_spec)
{
    GraphRun g;
    out = GraphResume();
    out.head = head_spec;
    jt_graph_build(c, head_spec, d_in, n_frames, rate, channels, fmt, frame_size, true, false, JT_GRAPH_NORMAL, nullptr, g, nullptr, &out);
    if (predicted_spec && out.cur.d && out.cur.fmt == JT_FMT_FLT && out.link_fmt == JT_FMT_FLT) {
        const std::vector<FilterNode> nodes = jt_parse_spec(*predicted_spec);
        if ((size_t)out.n_nodes < nodes.size() && nodes[(size_t)out.n_nodes].name == "afftdn") {
            // without the floor candidates of track_noise: whether the real spec tracks is one of the things the host is still
            // deciding (a spec that does runs its own forward pass, as before)
            AfftdnParams p = parse_afftdn(nodes[(size_t)out.n_nodes]);
            p.tn = 0;
            jt_afftdn_forward(c, out.cur, p, out.fwd);
        }
    }
}

// mono view of the source: what jt_downmix returns, without the device work (dry runs)
static Sig dry_mono(int64_t n_frames, int fmt, int rate) { Sig s; s.fmt = fmt; s.rate = rate; s.n = n_frames; s.d = nullptr; return s; }

// mode NORMAL: the whole stream, analysis kernels launched here.
// mode DRY   : no device work at all -- link sizes / formats and the sink-frame cadence of a stream of n_frames.
// mode CHUNK : [d_in, +n_frames) is a window of a longer stream (jt_graph_chunk): the audio filters run on it as if it
//              were a stream of its own, the signals the analysis filters see are recorded (GraphRun::*_sig) and their
//              kernels are left to the caller, who knows which part of the window is owned; no end-of-stream padding.
void jt_graph_build(jt_ctx *c, const std::string &spec, const void *d_in, int64_t n_frames, int rate, int channels,
                    int fmt, int frame_size, bool want_pcm, bool want_meta, int mode, const GraphChunk *chunk, GraphRun &g,
                    const GraphResume *resume, GraphResume *capture)
{
    if (n_frames < 0 || rate <= 0 || channels <= 0) JT_THROW(JT_ERR_INVALID_ARG, "bad stream description");
    if (frame_size <= 0) frame_size = 4096;
    std::vector<FilterNode> nodes = jt_parse_spec(spec);
    g = GraphRun();
    g.want_meta = want_meta;
    const bool dry = mode == JT_GRAPH_DRY, chunked = mode == JT_GRAPH_CHUNK;
    if (chunked) want_meta = false;              // the caller launches the analysis kernels on the owned part

    Exec E; E.c = c; E.dry = dry;
    bool have_mono = false, r128_prelaunched = false;
    cudaEvent_t tail_fork = nullptr;      // see the ebur128 node
    bool meter_deferred = false, meter_dual = false, meter_tp = false; Sig meter_sig;
    int64_t astats_prelaunched_upto = -1; const void *astats_prelaunched_sig = nullptr;
    const void *raw = d_in;
    if (channels == 1) { E.cur = dry ? dry_mono(n_frames, fmt, rate) : jt_downmix(c, raw, n_frames, 1, fmt, rate); have_mono = true; }
    E.link_fmt = fmt;
    size_t first_node = 0;
    if (resume) {
        // the head of this spec already ran (jt_graph_head): continue from the state behind it
        if (resume->n_nodes < 0 || (size_t)resume->n_nodes > nodes.size()) JT_THROW(JT_ERR_INVALID_ARG, "internal: graph resume");
        E.cur = resume->cur; E.link_fmt = resume->link_fmt; E.frames = resume->frames; E.chain = resume->chain; have_mono = true;
        first_node = (size_t)resume->n_nodes;
    } else E.frames = source_frames(n_frames, frame_size);

    for (size_t ni = first_node; ni < nodes.size(); ni++) {
        const FilterNode &f = nodes[ni];
        const bool last = ni + 1 == nodes.size();
        if (!dry) jt_check_cancel(c);
        if (!dry && c->trace) jt_trace(c, ("   node " + f.name).c_str());
        if (!have_mono) {
            // the only multi-channel-aware filter of the path is the leading downmix
            if (f.name == "aformat" && f.str("channel_layouts", "cl", "") == "mono") {
                E.cur = dry ? dry_mono(n_frames, fmt, rate) : jt_downmix(c, raw, n_frames, channels, fmt, rate);
                have_mono = true;
            } else JT_THROW(JT_ERR_UNSUPPORTED, "filter %s on %d-channel audio (specs of this path start with aformat=channel_layouts=mono)", f.name.c_str(), channels);
        }
        if (f.name == "aformat") {
            const std::string cl = f.str("channel_layouts", "cl", "mono");
            if (cl != "mono") JT_THROW(JT_ERR_UNSUPPORTED, "aformat channel_layouts=%s", cl.c_str());
            int out_rate = (int)f.num("sample_rates", "r", E.cur.rate);
            std::string sf = f.str("sample_fmts", "f", "");
            int out_fmt = 0;
            if (sf == "s16") out_fmt = JT_FMT_S16; else if (sf == "s32") out_fmt = JT_FMT_S32; else if (sf == "flt" || sf == "fltp") out_fmt = JT_FMT_FLT;
            else if (sf == "dbl" || sf == "dblp") out_fmt = JT_FMT_DBL; else if (!sf.empty()) JT_THROW(JT_ERR_UNSUPPORTED, "aformat sample_fmts=%s", sf.c_str());
            if (last && !want_pcm && out_rate != E.cur.rate) {
                // measure-only call: the resampled audio would be discarded, keep the frame cadence only
                SwrPlan p = jt_swr_plan(E.cur.rate, out_rate);
                E.flush();
                std::vector<FrameRef> nf; int64_t done = 0; const int64_t tot = p.out_count_flush(E.cur.n);
                for (const FrameRef &fr : E.frames) { int64_t cnt = std::min(p.out_count(fr.start + fr.nb), tot); if (cnt > done) { FrameRef o = fr; o.start = done; o.nb = (int32_t)(cnt - done); nf.push_back(o); done = cnt; } }
                if (tot > done) { FrameRef o; o.start = done; o.nb = (int32_t)(tot - done); o.ready = INT64_MAX; nf.push_back(o); }
                E.frames.swap(nf); E.cur.n = tot; E.cur.rate = out_rate; E.cur.d = nullptr;
            } else do_resample(E, out_rate, out_fmt);
        } else if (f.name == "aresample") {
            int out_rate = (int)f.num("sample_rate", "", E.cur.rate);
            if (const std::string *p = f.get("")) out_rate = atoi(p->c_str());
            do_resample(E, out_rate, 0);
        } else if (f.name == "asetnsamples") {
            const int n = (int)f.num("n", "nb_out_samples", 1024);
            const bool pad = f.flag("p", "pad", true);
            E.ureframe(n); E.flush();
            if (pad && !chunked && !E.frames.empty() && E.frames.back().nb < n) {
                const int64_t tot = E.frames.back().start + n;
                E.frames.back().nb = n;
                if (E.cur.d) { E.materialise(); E.cur = jt_pad_zero(c, E.cur, tot); } else E.cur.n = tot;
            }
        } else if (f.name == "atrim") {
            // libavfilter/trim.c: start_pts / duration_tb = av_rescale_q(usec, AV_TIME_BASE_Q, 1/rate), nearest
            const double st = f.num("start", "starti", 0.0), du = f.num("duration", "durationi", 0.0);
            const int64_t st_us = llround(st * 1e6), du_us = llround(du * 1e6);
            int64_t s0 = (st_us * E.cur.rate + 500000) / 1000000;
            int64_t len = du_us > 0 ? (du_us * E.cur.rate + 500000) / 1000000 : INT64_MAX / 4;
            if (f.get("start_sample") || f.get("end_sample")) {          // trim.c's sample-exact options
                s0 = (int64_t)f.num("start_sample", "", 0);
                const int64_t e1 = (int64_t)f.num("end_sample", "", (double)(INT64_MAX / 4));
                len = std::max<int64_t>(e1 - s0, 0);
            }
            const int64_t a = std::min(std::max<int64_t>(s0, 0), E.cur.n), b = std::min(E.cur.n, s0 + len);
            E.flush();
            std::vector<FrameRef> nf;
            for (const FrameRef &fr : E.frames) {
                const int64_t lo = std::max(fr.start, a), hi = std::min(fr.start + fr.nb, b);
                if (hi > lo) { FrameRef o = fr; o.start = lo - a; o.nb = (int32_t)(hi - lo); nf.push_back(o); }
            }
            E.frames.swap(nf);
            if (chunked) JT_THROW(JT_ERR_UNSUPPORTED, "atrim in a chunked graph");
            E.materialise();
            if (dry) E.cur.n = std::max<int64_t>(b - a, 0); else E.cur = jt_slice(E.cur, a, std::max<int64_t>(b - a, 0));
        } else if (f.name == "asetpts") {
            // timestamps only
        } else if (f.name == "highpass" || f.name == "lowpass") {
            E.materialise();
            if (!jt_valid_fmt(E.link_fmt)) JT_THROW(JT_ERR_UNSUPPORTED, "biquad format");
            const double freq = f.num("f", "frequency", 3000);
            const int poles = (int)f.num("p", "poles", 2);
            const std::string wt = f.str("t", "width_type", "q");
            const double width = f.num("w", "width", 0.707);
            const bool norm = f.flag("n", "normalize", false);
            const std::string tr = f.str("a", "transform", "di");
            const double mix = f.num("m", "mix", 1.0);
            if (poles != 2 || wt != "q") JT_THROW(JT_ERR_UNSUPPORTED, "%s poles=%d width_type=%s", f.name.c_str(), poles, wt.c_str());
            if (tr != "di" && tr != "tdii") JT_THROW(JT_ERR_UNSUPPORTED, "biquad transform %s", tr.c_str());
            BiquadCoef k = jt_biquad_design(f.name == "highpass", freq, width, E.cur.rate, norm);
            if (!dry) E.cur = jt_biquad(c, E.cur, k, tr == "tdii", mix);
        } else if (f.name == "anlmdn") {
            E.materialise(); E.storage(JT_FMT_FLT);
            const std::string om = f.str("o", "output", "o");
            if (om != "o") JT_THROW(JT_ERR_UNSUPPORTED, "anlmdn output mode %s", om.c_str());
            const double s = f.num("s", "strength", 0.00001), p = f.num("p", "patch", 0.002), r = f.num("r", "research", 0.006), m = f.num("m", "smooth", 11.0);
            if (!dry) E.cur = jt_anlmdn(c, E.cur, s, p, r, m);
            const int K = (int)llround(p * E.cur.rate);      // frames of H = 2K+1 samples
            E.ureframe(2 * K + 1);
        } else if (f.name == "afftdn") {
            E.materialise(); E.storage(JT_FMT_FLT);
            const AfftdnParams p = parse_afftdn(f);
            if (chunked && p.tn) {
                // the tracked noise floor is the one state of the chain with unbounded memory: its carry crosses chunks
                const int64_t A = E.cur.rate / 80;
                const int64_t a = chunk->link_pos(chunk->own_first, E.cur.rate) - chunk->link_pos(chunk->local_first, E.cur.rate);
                const int64_t b = chunk->last ? E.cur.n : chunk->link_pos(chunk->own_first + chunk->owned, E.cur.rate) - chunk->link_pos(chunk->local_first, E.cur.rate);
                if (a % A) JT_THROW(JT_ERR_INVALID_ARG, "chunk boundary is not on afftdn's hop grid (%lld samples)", (long long)A);
                AfftdnCarry cy; cy.hop0 = a / A; cy.hop1 = chunk->last ? (E.cur.n + A - 1) / A : (b + A - 1) / A;
                cy.key = chunk->own_first; cy.fn = chunk->exchange; cy.user = chunk->exchange_user; cy.n_ranks = chunk->n_ranks;
                E.cur = jt_afftdn(c, E.cur, p, &cy);
                g.exchanges++;
            } else if (!dry) E.cur = jt_afftdn(c, E.cur, p, nullptr, (resume && ni == first_node) ? &resume->fwd : nullptr);
            E.ureframe(E.cur.rate / 80);
        } else if (f.name == "agate") {
            E.materialise(); E.storage(JT_FMT_DBL);
            GateParams p;
            p.threshold = f.num("threshold", "", 0.125); p.ratio = f.num("ratio", "", 2); p.attack = f.num("attack", "", 20);
            p.release = f.num("release", "", 250); p.range = f.num("range", "", 0.06125); p.knee = f.num("knee", "", 2.828427125);
            p.makeup = f.num("makeup", "", 1); p.detection_rms = f.str("detection", "", "rms") == "rms";
            if (f.num("level_in", "", 1) != 1 || f.str("mode", "", "downward") != "downward") JT_THROW(JT_ERR_UNSUPPORTED, "agate level_in/mode");
            if (!dry) E.cur = jt_agate(c, E.cur, p);
        } else if (f.name == "acompressor") {
            E.materialise(); E.storage(JT_FMT_DBL);
            CompParams p;
            p.threshold = f.num("threshold", "", 0.125); p.ratio = f.num("ratio", "", 2); p.attack = f.num("attack", "", 20);
            p.release = f.num("release", "", 250); p.makeup = f.num("makeup", "", 1); p.knee = f.num("knee", "", 2.82843);
            p.mix = f.num("mix", "", 1); p.detection_rms = f.str("detection", "", "rms") == "rms";
            if (f.num("level_in", "", 1) != 1 || f.str("mode", "", "downward") != "downward" || f.str("link", "", "average") != "average")
                JT_THROW(JT_ERR_UNSUPPORTED, "acompressor level_in/mode/link");
            if (!dry) E.cur = jt_acompressor(c, E.cur, p);
        } else if (f.name == "deesser") {
            E.materialise(); E.storage(JT_FMT_DBL);
            if (f.str("s", "", "o") != "o") JT_THROW(JT_ERR_UNSUPPORTED, "deesser mode");
            if (!dry) E.cur = jt_deesser(c, E.cur, f.num("i", "", 0.0), f.num("m", "", 0.5), f.num("f", "", 0.5));
        } else if (f.name == "volume") {
            std::string v = f.str("volume", "", "1.0");
            if (const std::string *p = f.get("")) v = *p;
            char *end = nullptr; double g = strtod(v.c_str(), &end);
            if (end && !strcmp(end, "dB")) g = pow(10.0, g / 20.0); else if (end && *end) JT_THROW(JT_ERR_UNSUPPORTED, "volume expression '%s'", v.c_str());
            E.materialise();
            if (dry) E.cur.fmt = JT_FMT_FLT; else E.cur = jt_volume(c, E.cur, g);
            E.link_fmt = JT_FMT_FLT;
        } else if (f.name == "alimiter") {
            E.materialise(); E.storage(JT_FMT_DBL);
            LimiterParams p;
            p.limit = f.num("limit", "", 1); p.attack_ms = f.num("attack", "", 5); p.release_ms = f.num("release", "", 50);
            p.level_in = f.num("level_in", "", 1); p.level_out = f.num("level_out", "", 1); p.auto_level = f.flag("level", "", true);
            p.asc = f.flag("asc", "", false); p.asc_level = f.num("asc_level", "", 0.5); p.latency = f.flag("latency", "", false);
            if (!p.latency) JT_THROW(JT_ERR_UNSUPPORTED, "alimiter latency=0");
            if (!dry) E.cur = jt_alimiter(c, E.cur, p);
        } else if (f.name == "adeclick") {
            E.materialise(); E.storage(JT_FMT_DBL);
            const std::string m = f.str("m", "method", "a");
            const int save = (m == "s" || m == "save") ? 1 : 0;
            const double w = f.num("w", "window", 55), o = f.num("o", "overlap", 75);
            if (!dry) E.cur = jt_adeclick(c, E.cur, w, o, f.num("a", "arorder", 2), f.num("t", "threshold", 2), f.num("b", "burst", 2), save);
            const int ws = (int)(E.cur.rate * w / 1000.), hop = (int)(ws * (1. - o / 100.));
            E.ureframe(std::max(hop, 1));
        } else if (f.name == "loudnorm") {
            const double I = f.num("I", "i", -24), TP = f.num("TP", "tp", -2), LRA = f.num("LRA", "lra", 7);
            const double mI = f.num("measured_I", "measured_i", 0), mTP = f.num("measured_TP", "measured_tp", 99);
            const double mLRA = f.num("measured_LRA", "measured_lra", 0), mTh = f.num("measured_thresh", "", -70);
            double offset = f.num("offset", "", 0);
            const bool dual = f.flag("dual_mono", "", false);
            const bool lin_mode = jt_loudnorm_linear_mode(f);
            if (lin_mode) offset = I - mI;
            g.has_ln = true; g.ln_linear = lin_mode; g.ln_I = I;
            g.ln_dual = dual;
            if (lin_mode) {
                E.materialise(); E.storage(JT_FMT_DBL);
                g.ln_in_sig = E.cur;
                if (mode == JT_GRAPH_NORMAL) jt_loudnorm_meter_launch(c, E.cur, dual, g.ln_in);
                if (!dry) E.cur = jt_gain_f64(c, E.cur, pow(10., offset / 20.));
                g.ln_out_sig = E.cur;
                if (mode == JT_GRAPH_NORMAL) jt_loudnorm_meter_launch(c, E.cur, dual, g.ln_out);
            } else {
                // dynamic mode: af_loudnorm.c query_formats() forces BOTH links to 192 kHz / dbl
                const bool audio = want_pcm || !last;
                if (chunked && audio) JT_THROW(JT_ERR_UNSUPPORTED, "loudnorm dynamic mode with audio output in a chunked graph (linear-mode preconditions not met: measured_I=%g measured_TP=%g measured_LRA=%g measured_thresh=%g)", mI, mTP, mLRA, mTh);
                do_resample(E, 192000, JT_FMT_DBL, true);
                g.ln_in_sig = E.cur; g.ln_dynamic = true;
                const int64_t n192 = E.cur.n;
                const int F100 = 19200, F3000 = 576000;
                g.ln_type = n192 < F3000 ? 0 : 1;
                g.ln_in_extra = n192 < F3000 ? 0 : F3000 - F100;
                if (mode == JT_GRAPH_NORMAL) {
                    jt_loudnorm_meter_launch(c, E.cur, dual, g.ln_in);
                    if (g.ln_in_extra) jt_loudnorm_tail_launch(c, E.cur, n192, dual, g.ln_tail, &g.ln_tail_first);
                }
                if (audio) {
                    jt_loudnorm_opts o; o.I = I; o.TP = TP; o.LRA = LRA; o.measured_I = mI; o.measured_TP = mTP; o.measured_LRA = mLRA;
                    o.measured_thresh = mTh; o.offset = offset; o.dual_mono = dual;
                    if (dry) E.cur.fmt = JT_FMT_DBL;
                    else { int ty = 1; E.cur = jt_loudnorm_dynamic(c, E.cur, o, g.ln_in, &ty); }
                    E.link_fmt = JT_FMT_DBL;
                    g.ln_out_sig = E.cur; g.ln_has_out = true;
                    if (mode == JT_GRAPH_NORMAL) jt_loudnorm_meter_launch(c, E.cur, dual, g.ln_out);
                    if (n192 >= F3000) {
                        // output frames: 100 ms once the 3 s first frame is in, one per consumed 100 ms frame (a short one at
                        // EOF), then the 2.9 s flush frame
                        E.flush();
                        std::vector<FrameRef> nf; size_t a = 0;
                        auto ready_at = [&](int64_t consumed) {        // the link frame holding sample consumed - 1
                            while (a + 1 < E.frames.size() && E.frames[a + 1].start <= consumed - 1) a++;
                            return E.frames.empty() ? (int64_t)0 : E.frames[a].ready;
                        };
                        FrameRef fr; fr.start = 0; fr.nb = F100; fr.ready = ready_at(F3000); nf.push_back(fr);
                        int64_t pos = F100, src = F3000;
                        while (src < n192) {
                            const int nb = (int)std::min<int64_t>(F100, n192 - src); src += nb;
                            FrameRef o2; o2.start = pos; o2.nb = nb; o2.ready = nb == F100 ? ready_at(src) : INT64_MAX; nf.push_back(o2); pos += nb;
                        }
                        FrameRef fl; fl.start = pos; fl.nb = F3000 - F100; fl.ready = INT64_MAX; nf.push_back(fl);
                        E.frames.swap(nf);
                    }
                }
            }
        } else if (f.name == "astats") {
            E.materialise();
            // An analysis tail astats -> aspectralstats -> ebur128 on float storage (Pass 1 of a float file): the meter sees the very same
            // samples whichever of them runs first (only a dbl link is narrowed by aspectralstats' fltp negotiation), so its
            // kernels are queued now and the GPU works through them while the host does the frame bookkeeping of the three
            // nodes (3.5 ms per hour of audio -- at the head of Pass 1 the GPU had nothing else queued).
            if (want_meta && mode == JT_GRAPH_NORMAL && !r128_prelaunched && E.cur.fmt == JT_FMT_FLT) {
                size_t j = ni + 1;
                while (j < nodes.size() && (nodes[j].name == "astats" || nodes[j].name == "aspectralstats")) j++;
                if (j < nodes.size() && nodes[j].name == "ebur128") {
                    const FilterNode &e = nodes[j];
                    jt_ebur128_launch(c, E.cur, e.flag("dualmono", "", false), e.str("peak", "", "none").find("true") != std::string::npos, g.r128p);
                    r128_prelaunched = true;
                    // the same for astats itself when nothing but analysis nodes follow to the end of the spec: the sink's last
                    // frame then carries the statistics of the whole signal (checked against the final cadence below)
                    size_t e2 = j + 1;
                    while (e2 < nodes.size() && (nodes[e2].name == "astats" || nodes[e2].name == "aspectralstats" || nodes[e2].name == "ebur128")) e2++;
                    if (e2 == nodes.size() && E.cur.n > 0) {
                        if (c->defer_astats && c->low_stream) {
                            // behind everything queued so far on the main stream (the signal exists, every block the arena
                            // handed out again is free), then on its own: the caller keeps the pass's buffers until it has
                            // waited for astats' event
                            cudaEvent_t fork = jt_record_event(c);
                            JT_CUDA(cudaStreamWaitEvent(c->low_stream, fork, 0));
                            cudaStream_t main_stream = c->stream;
                            c->stream = c->low_stream;
                            try { jt_astats_launch(c, E.cur, E.cur.n, g.astp); } catch (...) { c->stream = main_stream; throw; }
                            c->stream = main_stream;
                            g.astats_on_low = true;
                        } else jt_astats_launch(c, E.cur, E.cur.n, g.astp);
                        astats_prelaunched_upto = E.cur.n; astats_prelaunched_sig = E.cur.d;
                    }
                }
            }
            E.has_astats = true; E.astats_sig = E.cur;
            { const std::string mp = f.str("measure_perchannel", "", "all"); E.astats_overall_only = (mp == "0" || mp == "none"); }
            E.stamp(JT_LVL_ASTATS);
        } else if (f.name == "aspectralstats") {
            E.materialise(); E.storage(JT_FMT_FLT);
            const int win = (int)f.num("win_size", "", 2048);
            if (f.str("win_func", "", "hann") != "hann" && f.str("win_func", "", "hann") != "hanning") JT_THROW(JT_ERR_UNSUPPORTED, "aspectralstats win_func");
            if (f.num("overlap", "", 0.5) != 0.5) JT_THROW(JT_ERR_UNSUPPORTED, "aspectralstats overlap");
            E.has_spec = true; E.spec_sig = E.cur; E.spec_win = win;    // computed once the sink cadence is known
            E.ureframe(win / 2);
            E.stamp(JT_LVL_HOP);
        } else if (f.name == "ebur128") {
            const std::string peak = f.str("peak", "", "none");
            const bool tp = peak.find("true") != std::string::npos;
            const bool dual = f.flag("dualmono", "", false);
            // input link is dbl; s16/flt storage widens exactly on load
            if (want_meta && mode == JT_GRAPH_NORMAL && !r128_prelaunched) {
                // The meter's kernels (the true-peak oversampler above all: 2 - 7 ms at one CTA of 640 threads per SM) leave most of
                // an SM's warp slots empty, and the graph tail below queues aspectralstats and astats of the very signals that exist
                // by now.  Mark this point of the main stream: the tail forks from here onto the low-priority stream and runs
                // under the meter and the output stage instead of behind them.  (Only when no astats / aspectralstats node follows,
                // i.e. their input -- and its format conversion -- is already queued; small region graphs are not worth two events.)
                bool later_analysis = false;
                for (size_t j = ni + 1; j < nodes.size(); j++) if (nodes[j].name == "astats" || nodes[j].name == "aspectralstats") later_analysis = true;
                static const bool no_fork = getenv("JT_NO_TAIL_FORK") != nullptr;
                if (!no_fork && !c->timing && !later_analysis && c->low_stream && (E.has_astats || E.has_spec) && E.cur.n >= (1 << 22)) tail_fork = jt_record_event(c);
                // the meter does not change the audio: when only the output stage follows and the caller asked for it, its kernels
                // go behind that stage (graph tail), so `out_ready` -- and with it the download of the result -- comes first
                bool only_output_follows = want_pcm && c->meter_after_output;
                for (size_t j = ni + 1; j < nodes.size(); j++)
                    if (!(nodes[j].name == "aformat" || nodes[j].name == "aresample" || nodes[j].name == "asetnsamples")) only_output_follows = false;
                if (only_output_follows) { meter_deferred = true; meter_sig = E.cur; meter_dual = dual; meter_tp = tp; }
                else jt_ebur128_launch(c, E.cur, dual, tp, g.r128p);
            }
            r128_prelaunched = false;
            g.r128_sig = E.cur; g.r128_dual = dual; g.r128_tp = tp;
            E.has_r128 = true; E.link_fmt = JT_FMT_DBL;
            const int tick = E.cur.rate / 10;
            E.ureframe(tick);
            E.stamp(JT_LVL_TICK);
        } else {
            JT_THROW(JT_ERR_UNSUPPORTED, "filter '%s' is not part of the jivetalking hot path", f.name.c_str());
        }
    }
    if (!have_mono) JT_THROW(JT_ERR_UNSUPPORTED, "%d-channel graph without a mono downmix", channels);
    if (capture) { capture->n_nodes = (int)nodes.size(); capture->cur = E.cur; capture->link_fmt = E.link_fmt; capture->frames = E.frames; capture->chain = E.chain; return; }
    E.flush();
    if (E.cur.d && want_pcm) E.materialise();
    if (E.cur.d && want_pcm && mode == JT_GRAPH_NORMAL) g.out_ready = jt_record_event(c);     // the sink audio is complete; analysis kernels follow
    g.out = E.cur; g.out.fmt = E.cur.d ? E.cur.fmt : E.link_fmt;
    g.frames.swap(E.frames);
    g.has_astats = E.has_astats; g.has_spec = E.has_spec; g.has_r128 = E.has_r128; g.astats_overall_only = E.astats_overall_only;
    g.astats_sig = E.astats_sig; g.spec_sig = E.spec_sig; g.spec_win = E.spec_win;
    for (size_t i = 0; i < g.frames.size(); i++) if (g.frames[i].astats_pos >= 0) g.last_astats_frame = (long)i;
    if (!want_meta || mode != JT_GRAPH_NORMAL) return;
    if (c->trace) jt_trace(c, "   graph tail");
    // ---- analysis kernels whose input cadence depends on the final sink framing -------------
    const size_t nf = g.frames.size();
    cudaStream_t main_stream = c->stream;
    if (meter_deferred) jt_ebur128_launch(c, meter_sig, meter_dual, meter_tp, g.r128p);
    if (tail_fork) {
        JT_CUDA(cudaStreamWaitEvent(c->low_stream, tail_fork, 0));
        c->stream = c->low_stream;
    }
    try {
    if (g.has_spec) {
        // aspectralstats emits one row per 1024-sample hop, but a sink frame only ever shows the row of the
        // hop holding its first sample: compute just those (plus predecessors for the flux term)
        std::vector<int64_t> wanted; wanted.reserve(nf);
        for (size_t i = 0; i < nf; i++) if (g.frames[i].hop >= 0) wanted.push_back(g.frames[i].hop);
        jt_aspectralstats_launch(c, g.spec_sig, g.spec_win, &wanted, g.specp);
    }
    if (g.has_astats && g.last_astats_frame >= 0 &&
        !(astats_prelaunched_upto == g.frames[g.last_astats_frame].astats_pos && astats_prelaunched_sig == g.astats_sig.d)) {
        g.astats_on_low = false;                  // (a pre-launch that missed the final cadence is simply not read)
        jt_astats_launch(c, g.astats_sig, g.frames[g.last_astats_frame].astats_pos, g.astp);
    }
    } catch (...) { c->stream = main_stream; throw; }
    if (tail_fork) {
        // join: whatever the caller queues next on the main stream (or hands back to the arena) comes after the forked kernels
        cudaEvent_t done = jt_record_event(c);
        c->stream = main_stream;
        JT_CUDA(cudaStreamWaitEvent(main_stream, done, 0));
    }
}

const R128Result &jt_graph_r128_early(jt_ctx *c, GraphRun &g)
{
    if (!g.r128_done) { if (g.has_r128 && g.want_meta) jt_ebur128_finish(c, g.r128p, g.r128); g.r128_done = true; }
    return g.r128;
}

// one jt_frame_meta per sink frame from the per-tick / per-hop products (astats is added by the caller);
// records are independent: the printf-rounding of ~20 values per sink frame is spread over host threads in `pool`
static void assemble_records(const std::vector<FrameRef> &frames, bool has_r128, const R128Result &r128, bool has_spec,
                             const std::vector<float> &spec_rows, int64_t spec_hops, GraphResult &res, std::vector<std::thread> &pool)
{
    const size_t nf = frames.size();
    res.meta.resize(nf); res.meta_ready.resize(nf);
    int64_t last_tick = -1;
    for (size_t i = 0; i < nf; i++) if (frames[i].tick >= 0) last_tick = std::max<int64_t>(last_tick, frames[i].tick);
    auto fill = [&frames, has_r128, &r128, has_spec, &spec_rows, spec_hops, &res, last_tick](size_t i0, size_t i1) {
        for (size_t i = i0; i < i1; i++) {
            const FrameRef &fr = frames[i];
            jt_frame_meta &m = res.meta[i];
            double *dp = &m.r128_M;
            const size_t ndbl = (sizeof(jt_frame_meta) - offsetof(jt_frame_meta, r128_M)) / sizeof(double);
            for (size_t k = 0; k < ndbl; k++) dp[k] = NAN;
            m.first_sample = fr.start; m.nb_samples = fr.nb; m.reserved = 0;
            res.meta_ready[i] = fr.ready;
            if (has_r128 && fr.tick >= 0 && fr.tick < r128.n_ticks) {
                const int64_t k = fr.tick;
                m.r128_M = jt_wire("%.3f", r128.M[k]); m.r128_S = jt_wire("%.3f", r128.S[k]);
                m.r128_sample_peak = jt_wire("%.3f", r128.sp_cum[k]);
                m.r128_true_peak = jt_wire("%.3f", r128.tp_cum[k]);
                if (k == last_tick) {
                    m.r128_I = jt_wire("%.3f", r128.I); m.r128_LRA = jt_wire("%.3f", r128.LRA);
                    m.r128_LRA_low = jt_wire("%.3f", r128.LRA_low); m.r128_LRA_high = jt_wire("%.3f", r128.LRA_high);
                }
            }
            if (has_spec && fr.hop >= 0 && fr.hop < spec_hops)
                for (int k = 0; k < JT_SP_COUNT; k++) m.spectral[k] = jt_wire("%g", (double)spec_rows[(size_t)fr.hop * JT_SP_COUNT + k]);
        }
    };
    const unsigned hw = jt_host_threads();
    if (nf < 8192 || hw == 1) fill(0, nf);
    else {
        const size_t per = (nf + hw - 1) / hw;
        for (unsigned t = 0; t < hw; t++) { const size_t a = t * per, b = std::min(nf, a + per); if (a < b) pool.emplace_back(fill, a, b); }
    }
}

void jt_graph_finish(jt_ctx *c, GraphRun &g, GraphResult &res)
{
    res = GraphResult();
    memset(&res.ln, 0, sizeof(res.ln));
    res.out = g.out;
    if (g.has_ln) {
        LoudnormMeter mi, mo;
        if (g.ln_dynamic && g.ln_in_extra && g.ln_tail.nt > 0) {
            // the stream's ticks, then -- from the tick holding the stream's end on -- those of the re-metered tail
            JT_CUDA(cudaEventSynchronize(g.ln_in.ev)); JT_CUDA(cudaEventSynchronize(g.ln_tail.ev));
            const int64_t split = g.ln_in_sig.n / g.ln_in.s100, t0 = g.ln_tail_first;
            const int64_t nt = t0 + g.ln_tail.nt, nfull = t0 + g.ln_tail.nfull;
            std::vector<double> hp(nt), hk(nt);
            for (int64_t k = 0; k < nt; k++) {
                if (k < split) { hp[k] = g.ln_in.hp[k]; hk[k] = g.ln_in.hk[k]; }
                else { hp[k] = g.ln_tail.hp[k - t0]; hk[k] = g.ln_tail.hk[k - t0]; }
            }
            jt_loudnorm_meter_host_finalize(hp.data(), hk.data(), nt, nfull, g.ln_in.s100, g.ln_dual, mi);
        } else jt_loudnorm_meter_finish(c, g.ln_in, mi);
        res.ln.valid = 1;
        if (g.ln_linear || g.ln_has_out) {
            jt_loudnorm_meter_finish(c, g.ln_out, mo);
            res.ln.normalization_type = g.ln_linear ? 0 : g.ln_type;
            res.ln.output_i = mo.I; res.ln.output_tp = 20. * log10(mo.sample_peak); res.ln.output_lra = mo.LRA; res.ln.output_thresh = mo.thresh;
            res.ln.target_offset = g.ln_I - mo.I;
        } else {
            // measure-only call (Pass 3): the discarded output is not produced, its four values are not known
            res.ln.normalization_type = g.ln_type;
            res.ln.output_i = res.ln.output_tp = res.ln.output_lra = res.ln.output_thresh = res.ln.target_offset = NAN;
        }
        res.ln.input_i = mi.I; res.ln.input_tp = 20. * log10(mi.sample_peak); res.ln.input_lra = mi.LRA; res.ln.input_thresh = mi.thresh;
    }
    if (!g.want_meta) return;
    const R128Result &r128 = jt_graph_r128_early(c, g);
    jt_trace(c, " finish: meter values");
    std::vector<float> spec_rows; int64_t spec_hops = 0;
    if (g.has_spec) jt_aspectralstats_finish(c, g.specp, spec_rows, spec_hops);
    jt_trace(c, " finish: spectral rows");
    const std::vector<FrameRef> &frames = g.frames;
    JtHost hmeta(c, "meta_assembly");
    const long last_astats_frame = g.last_astats_frame;
    std::vector<std::thread> pool;
    assemble_records(frames, g.has_r128, r128, g.has_spec, spec_rows, spec_hops, res, pool);
    AstatsResult a; bool have_a = false; JtError aerr{0, ""};
    if (g.has_astats && last_astats_frame >= 0 && !g.astats_later) {
        try { jt_astats_finish(c, g.astp, a); have_a = true; }
        catch (const JtError &e) { aerr = e; }
    }
    for (auto &t : pool) t.join();
    if (aerr.code) throw aerr;
    if (have_a) {
        jt_frame_meta &m = res.meta[last_astats_frame];
        if (!g.astats_overall_only) for (int k = 0; k < JT_AS_COUNT; k++) m.astats[k] = std::isnan(a.v[k]) ? NAN : jt_wire("%f", a.v[k]);
        m.astats_overall_RMS_level = jt_wire("%f", a.overall_rms);
        m.astats_overall_Peak_level = jt_wire("%f", a.overall_peak);
    }
}

// What the Go side accumulates from a pass's sink-frame records (latest-wins loudness values, mean of the spectral rows over
// the frames that carry them: analyser_metrics.go:898-942) WITHOUT materialising the records: the per-frame work is the
// "%g" rounding of 13 spectral values, spread over host threads.  Equal to feeding jt_assemble_records' output to the
// accumulator, up to the order in which the spectral sums are added.
void jt_accumulate_frames(const std::vector<FrameRef> &frames, bool has_r128, const R128Result &r128, bool has_spec,
                          const std::vector<float> &spec_rows, int64_t spec_hops, jt_measurements *out)
{
    memset(out, 0, sizeof(*out));
    for (double &a : out->astats) a = NAN;
    const size_t nf = frames.size();
    out->sink_frames = (int64_t)nf;
    int64_t last_tick = -1; long last_tick_frame = -1;
    for (size_t i = 0; i < nf; i++) if (frames[i].tick >= 0 && frames[i].tick < r128.n_ticks && (int64_t)frames[i].tick >= last_tick) { last_tick = frames[i].tick; last_tick_frame = (long)i; }
    if (has_r128) {
        // latest wins: the last frame that carries a tick gives M / S / peaks; I and LRA ride on the stream's last tick
        for (long i = (long)nf - 1; i >= 0; i--) {
            const FrameRef &fr = frames[(size_t)i];
            if (fr.tick >= 0 && fr.tick < r128.n_ticks) {
                const int64_t k = fr.tick;
                out->last_m = jt_wire("%.3f", r128.M[k]); out->last_s = jt_wire("%.3f", r128.S[k]);
                const double sp = jt_wire("%.3f", r128.sp_cum[k]), tp = jt_wire("%.3f", r128.tp_cum[k]);
                out->input_sp = sp <= 0 ? -120.0 : 20 * log10(sp); out->input_tp = tp <= 0 ? -120.0 : 20 * log10(tp);
                break;
            }
        }
        if (last_tick_frame >= 0) { out->input_i = jt_wire("%.3f", r128.I); out->input_lra = jt_wire("%.3f", r128.LRA); }
    }
    if (has_spec && nf) {
        const unsigned hw = jt_host_threads();
        const unsigned nth = nf < 8192 ? 1u : hw;
        std::vector<std::array<double, JT_SP_COUNT + 1>> part(nth);
        auto work = [&](unsigned t) {
            std::array<double, JT_SP_COUNT + 1> acc{}; 
            const size_t per = (nf + nth - 1) / nth, a = t * per, b = std::min(nf, a + per);
            for (size_t i = a; i < b; i++) {
                const FrameRef &fr = frames[i];
                if (fr.hop < 0 || fr.hop >= spec_hops) continue;
                double v[JT_SP_COUNT]; bool any = false;
                for (int k = 0; k < JT_SP_COUNT; k++) { v[k] = jt_wire("%g", (double)spec_rows[(size_t)fr.hop * JT_SP_COUNT + k]); any = any || !std::isnan(v[k]); }
                if (!any) continue;
                for (int k = 0; k < JT_SP_COUNT; k++) acc[k] += std::isnan(v[k]) ? 0.0 : v[k];
                acc[JT_SP_COUNT] += 1.0;
            }
            part[t] = acc;
        };
        std::vector<std::thread> pool;
        for (unsigned t = 1; t < nth; t++) pool.emplace_back(work, t);
        work(0);
        for (auto &t : pool) t.join();
        double sum[JT_SP_COUNT] = {0}; double cnt = 0;
        for (unsigned t = 0; t < nth; t++) { for (int k = 0; k < JT_SP_COUNT; k++) sum[k] += part[t][k]; cnt += part[t][JT_SP_COUNT]; }
        out->spectral_frames = (int64_t)cnt;
        for (int k = 0; k < JT_SP_COUNT; k++) out->spectral_mean[k] = cnt > 0 ? sum[k] / cnt : 0.0;
    }
}

// jt_graph_finish for a caller that wants the accumulated measurements only (Pass 2 / Pass 4 inside the four-pass drivers)
void jt_graph_finish_acc(jt_ctx *c, GraphRun &g, jt_measurements *acc, jt_loudnorm_stats *ln)
{
    const bool want = g.want_meta;
    g.want_meta = false;
    GraphResult res;
    jt_graph_finish(c, g, res);                  // loudnorm statistics; no records
    g.want_meta = want;
    if (ln) *ln = res.ln;
    if (!acc) return;
    const R128Result &r128 = jt_graph_r128_early(c, g);
    std::vector<float> spec_rows; int64_t spec_hops = 0;
    if (g.has_spec) jt_aspectralstats_finish(c, g.specp, spec_rows, spec_hops);
    JtHost hmeta(c, "meta_assembly");
    jt_accumulate_frames(g.frames, g.has_r128, r128, g.has_spec, spec_rows, spec_hops, acc);
    if (g.has_astats && g.last_astats_frame >= 0) {
        AstatsResult a;
        jt_astats_finish(c, g.astp, a);
        if (!g.astats_overall_only) for (int k = 0; k < JT_AS_COUNT; k++) acc->astats[k] = std::isnan(a.v[k]) ? NAN : jt_wire("%f", a.v[k]);
    }
    acc->duration_s = g.out.rate > 0 ? (double)g.out.n / g.out.rate : 0.0;
}

void jt_assemble_records(const std::vector<FrameRef> &frames, bool has_r128, const R128Result &r128, bool has_spec,
                         const std::vector<float> &spec_rows, int64_t spec_hops, GraphResult &res)
{
    std::vector<std::thread> pool;
    assemble_records(frames, has_r128, r128, has_spec, spec_rows, spec_hops, res, pool);
    for (auto &t : pool) t.join();
}

// Pass-1 sink-frame records of a stream whose per-tick / per-hop / astats products were computed elsewhere (several
// GPUs, jt_analyse_chunk): the frame cadence of Pass1FilterOrder (filters.go:42-45) is pure integer bookkeeping.
void jt_pass1_records(jt_ctx *c, int64_t n_frames, int rate, int frame_size, const R128Result &r128,
                      const std::vector<float> &spec_rows, int64_t spec_hops, const AstatsResult *astats, GraphResult &res)
{
    res = GraphResult();
    memset(&res.ln, 0, sizeof(res.ln));
    std::vector<FrameRef> frames = source_frames(n_frames, frame_size);
    for (FrameRef &fr : frames) fr.astats_pos = fr.start + fr.nb;                       // astats
    frames = reframe(frames, n_frames, 1024);                                          // aspectralstats win 2048
    for (size_t j = 0; j < frames.size(); j++) frames[j].hop = (int32_t)j;
    const int tick = rate / 10;
    frames = reframe(frames, n_frames, tick);                                          // ebur128 metadata=1
    for (size_t k = 0; k < frames.size(); k++) if (frames[k].nb == tick) frames[k].tick = (int32_t)k;
    std::vector<std::thread> pool;
    assemble_records(frames, true, r128, true, spec_rows, spec_hops, res, pool);
    for (auto &t : pool) t.join();
    long last_astats_frame = -1;
    for (size_t i = 0; i < frames.size(); i++) if (frames[i].astats_pos >= 0) last_astats_frame = (long)i;
    if (astats && last_astats_frame >= 0) {
        jt_frame_meta &m = res.meta[last_astats_frame];
        for (int k = 0; k < JT_AS_COUNT; k++) m.astats[k] = std::isnan(astats->v[k]) ? NAN : jt_wire("%f", astats->v[k]);
        m.astats_overall_RMS_level = jt_wire("%f", astats->overall_rms);
        m.astats_overall_Peak_level = jt_wire("%f", astats->overall_peak);
    }
    (void)c;
}
