// jt_adapt.cu -- the host logic between Pass 1 and Pass 2: noise-floor seed, voice-activity
// detector, speech / room-tone election, gate statistics, AdaptConfig and BuildFilterSpec.  Pure scalar host code over the
// 250 ms interval records Pass 1 produced on the GPU; nothing here touches the device.
//
// Restated from the reference's behaviour (cited per function), not from its text: one IntervalView with the timestamps
// in integer nanoseconds (Go's time.Duration) and index ranges instead of slice copies.  Go semantics that matter for
// equality are kept explicitly: max/min propagate NaN, slices.Sort puts NaN first, int(x) truncates, Duration.Seconds()
// splits seconds and nanoseconds, fmt's %g prints the shortest digits that round-trip.
#include "../../include/jtdsp.h"
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <string>
#include <vector>

extern "C" void jt_vad_assign_astats(const jt_measurements *m, jt_voice_activity *va);

namespace {

typedef int64_t ns_t;
const ns_t kMs = 1000000LL, kSec = 1000000000LL;

// ---- constants (analyser_vad.go:16-32,57,175-185,306,363-372,673; analyser_candidates_speech.go:9-122,
//      337-352; analyser_noise_seed.go:20-65) ----
const ns_t   kIntervalHop = 250 * kMs, kGoldenInterval = 250 * kMs;
const ns_t   kMinSpeech = 10 * kSec, kGapTolFloor = 2 * kSec, kGapTolCeil = 10 * kSec;
const double kLevelFloorDB = -115.0;
const double kGateVoicedLowPct = 10.0, kGateNoiseHighPct = 95.0, kNoiseFloorPct = 10.0;
const double kHystFraction = 0.25, kHystFallbackDB = 1.0;
const double kVoiceActivatedFraction = 0.20;
const ns_t   kIdealMin = 8 * kSec, kIdealMax = 18 * kSec;
const double kCentroidMin = 200.0, kCentroidMax = 6000.0, kNoiseMarginDB = 2.0, kEntropyMax = 0.70;
const double kVoicingDensityThr = 0.6, kVoicedKurtosis = 4.5;
const double kRollIdealMin = 4000.0, kRollIdealMax = 8000.0, kRollOkMin = 2500.0, kRollOkMax = 10000.0;
const double kFluxStable = 0.004, kFluxNormal = 0.010, kFluxTransient = 0.020, kFluxOk = 0.030;
const double kMinSNR = 20.0, kSNRSaturation = 40.0;
const double kWKurt = 0.15, kWFlat = 0.10, kWCent = 0.10, kWRMS = 0.10, kWCons = 0.10, kWVoic = 0.15, kWRoll = 0.15, kWFlux = 0.15;
const ns_t   kGoldenSpeechWindow = 60 * kSec, kGoldenSpeechMin = 30 * kSec;
const ns_t   kGoldenWindow = 10 * kSec, kGoldenMin = 8 * kSec;
const double kGroundSNRW = 0.6, kGroundDurW = 0.4, kGroundTieMax = 0.02, kGroundVarCap = 25.0;
const double kMinViableScore = 0.3;
const double kRoomAmpDecayDB = 6.0, kRoomAmpW = 0.6, kRoomFluxW = 0.4;
const int    kSilenceMinIntervals = 10, kSeedTopDiv = 5, kSeedMinCount = 8;
const double kSilenceHeadroomDB = 1.0, kSilenceFallbackHeadroom = 6.0, kSilenceMinThr = -70.0, kSilenceMaxThr = -35.0;

// ---- Go arithmetic helpers ----
inline double gomax(double a, double b) { if (std::isnan(a) || std::isnan(b)) return NAN; return a > b ? a : b; }
inline double gomin(double a, double b) { if (std::isnan(a) || std::isnan(b)) return NAN; return a < b ? a : b; }
inline bool   go_less(double x, double y) { return (std::isnan(x) && !std::isnan(y)) || x < y; }          // cmp.Less
inline int    go_cmp(double x, double y) {                                                                 // cmp.Compare
    const bool xn = std::isnan(x), yn = std::isnan(y);
    if (xn) return yn ? 0 : -1;
    if (yn) return 1;
    return x < y ? -1 : (x > y ? 1 : 0);
}
inline void   go_sort(std::vector<double> &v) { std::sort(v.begin(), v.end(), go_less); }
// element k of the slice as slices.Sort would leave it, without sorting the rest (the detector only ever reads single order
// statistics of its 14 400-per-hour level lists: six full sorts were half of its run time)
inline double go_kth(std::vector<double> &v, int64_t k) { std::nth_element(v.begin(), v.begin() + k, v.end(), go_less); return v[(size_t)k]; }
inline double go_pct(std::vector<double> &v, double pct)           // pct_sorted of the sorted slice
{
    if (v.empty()) return 0;
    pct = pct < 0 ? 0 : (pct > 100 ? 100 : pct);
    return go_kth(v, (int64_t)(pct / 100 * (double)(v.size() - 1)));
}
inline double dur_seconds(ns_t d) { return (double)(d / kSec) + (double)(d % kSec) / 1e9; }              // Duration.Seconds
inline bool   is_finite(double v) { return !std::isnan(v) && !std::isinf(v); }
inline double db_to_linear(double db) { return pow(10.0, db / 20.0); }                                    // filters.go:591-593
inline double linear_to_db(double l) { return l <= 0 ? -120.0 : 20.0 * log10(l); }                        // filters.go:597-602
inline double sanitize(double v, double d) { return is_finite(v) ? v : d; }                                  // adaptive_math.go:6-11

inline ns_t   ts_of(const jt_interval &iv) { return (ns_t)llround(iv.timestamp_s * 1e9); }
inline double level_of(const jt_interval &iv, int axis) { return axis == 1 ? iv.rms_level : iv.momentary_lufs; }
inline bool   is_floored(double l) { return std::isinf(l) || std::isnan(l) || l <= kLevelFloorDB; }       // analyser_vad.go:66-68
inline int64_t intervals_for(ns_t d, ns_t hop) { return hop <= 0 ? 0 : (d + hop / 2) / hop; }             // analyser_vad.go:38-43

struct View {                      // the interval stream with exact timestamps
    const jt_interval *iv; int64_t n; std::vector<ns_t> ts;
    View(const jt_interval *p, int64_t cnt) : iv(p), n(p ? cnt : 0), ts((size_t)(p ? cnt : 0)) { for (int64_t i = 0; i < n; i++) ts[i] = ts_of(p[i]); }
};
struct Span { int64_t lo = 0, hi = 0; int64_t len() const { return hi - lo; } };

// getIntervalsInRange (analyser_candidates_shared.go:85-113): timestamps in [start, end)
Span in_range(const View &v, ns_t start, ns_t end)
{
    Span s;
    if (v.n == 0) return s;
    const int64_t lo = std::lower_bound(v.ts.begin(), v.ts.end(), start) - v.ts.begin();
    if (lo >= v.n) return s;
    int64_t hi = lo;
    while (hi < v.n && v.ts[hi] < end) hi++;
    s.lo = lo; s.hi = hi;
    return s;
}

// ---- histogram / Otsu (analyser_vad.go:84-150, 260-304, 374-400) ----
struct Hist { std::vector<int> bins; double bw = 0, lo = 0, hi = 0; int64_t count = 0;
              double centre(int64_t i) const { return lo + ((double)i + 0.5) * bw; } };

Hist build_hist(const View &v, int axis, double bw)
{
    Hist h;
    if (bw <= 0) return h;
    std::vector<double> lv; lv.reserve(v.n);
    double mn = INFINITY, mx = -INFINITY;
    for (int64_t i = 0; i < v.n; i++) {
        const double l = level_of(v.iv[i], axis);
        if (is_floored(l)) continue;
        lv.push_back(l); mn = gomin(mn, l); mx = gomax(mx, l);
    }
    if (lv.empty()) return h;
    const int64_t nb = (int64_t)((mx - mn) / bw) + 1;
    h.bins.assign((size_t)nb, 0); h.bw = bw; h.lo = mn; h.hi = mx;
    for (double l : lv) { int64_t k = (int64_t)((l - mn) / bw); if (k >= nb) k = nb - 1; h.bins[k]++; h.count++; }
    return h;
}

double otsu(const Hist &h)
{
    const int64_t nb = (int64_t)h.bins.size();
    if (h.count == 0 || nb < 2) return (h.lo + h.hi) / 2;
    const double total = (double)h.count;
    double sum_all = 0;
    for (int64_t i = 0; i < nb; i++) sum_all += h.centre(i) * (double)h.bins[i];
    double wb = 0, sb = 0, best = 0; int64_t best_i = -1;
    for (int64_t i = 0; i < nb - 1; i++) {
        wb += (double)h.bins[i]; sb += h.centre(i) * (double)h.bins[i];
        const double wf = total - wb;
        if (wb == 0 || wf == 0) continue;
        const double d = sb / wb - (sum_all - sb) / wf, var = wb * wf * d * d;
        if (var > best) { best = var; best_i = i; }
    }
    if (best_i < 0) return (h.lo + h.hi) / 2;
    return h.lo + (double)(best_i + 1) * h.bw;
}

double hysteresis_margin(const Hist &h, double split)
{
    double w = 0, c = 0;
    for (int64_t i = 0; i < (int64_t)h.bins.size(); i++) { const double ce = h.centre(i); if (ce >= split) { w += ce * (double)h.bins[i]; c += (double)h.bins[i]; } }
    const double upper = c == 0 ? split : w / c, dist = upper - split;
    return dist <= 0 ? kHystFallbackDB : dist * kHystFraction;
}

std::vector<double> vad_levels(const View &v, int axis)
{
    std::vector<double> lv; lv.reserve(v.n);
    for (int64_t i = 0; i < v.n; i++) { const double l = level_of(v.iv[i], axis); if (!is_floored(l)) lv.push_back(l); }
    go_sort(lv);
    return lv;
}
double pct_sorted(const double *s, int64_t n, double pct)                       // analyser_vad.go:166-173
{
    if (n <= 0) return 0;
    pct = gomax(0, gomin(100, pct));
    return s[(int64_t)(pct / 100 * (double)(n - 1))];
}
double pct_sorted(const std::vector<double> &s, double pct) { return pct_sorted(s.data(), (int64_t)s.size(), pct); }
double percentile_floor(const double *s, int64_t n, double seed) { return gomax(pct_sorted(s, n, kNoiseFloorPct), seed + kNoiseMarginDB); }
double clamp_split(double split, double floor, double p75)                       // analyser_vad.go:325-331
{
    const double lower = floor + kNoiseMarginDB;
    if (p75 < lower) return lower;
    return gomax(lower, gomin(p75, split));
}
inline bool veto_ok(const jt_interval &s) { return s.spectral[JT_SP_centroid] >= kCentroidMin && s.spectral[JT_SP_centroid] <= kCentroidMax && s.spectral[JT_SP_entropy] < kEntropyMax; }
inline bool is_speech(const jt_interval &s, double split, int axis) { return level_of(s, axis) >= split && veto_ok(s); }

int gap_tolerance(const uint8_t *flags, int64_t n, ns_t hop)                     // analyser_vad.go:407-444
{
    const int64_t fl = intervals_for(kGapTolFloor, hop), ce = intervals_for(kGapTolCeil, hop);
    int64_t first = -1, last = -1;
    for (int64_t i = 0; i < n; i++) if (flags[i]) { if (first < 0) first = i; last = i; }
    if (first < 0) return (int)fl;
    std::vector<double> gaps; int64_t g = 0;
    for (int64_t i = first; i <= last; i++) { if (flags[i]) { if (g > 0) gaps.push_back((double)g); g = 0; } else g++; }
    if (gaps.empty()) return (int)fl;
    go_sort(gaps);
    const int64_t p75 = (int64_t)round(pct_sorted(gaps, 75));
    return (int)std::max(fl, std::min(ce, p75));
}

// buildSpeechRuns (analyser_vad.go:474-556)
std::vector<jt_region> speech_runs(const View &v, double split, double margin, int tol, int axis, ns_t hop)
{
    std::vector<jt_region> runs;
    const int64_t min_iv = intervals_for(kMinSpeech, hop);
    if (v.n < min_iv || min_iv <= 0) return runs;
    const double high = split + margin, low = split - margin;
    ns_t run_start = 0; int64_t speech_count = 0, last_speech = 0; int pending = 0; bool in_run = false;
    auto flush = [&](int64_t end_idx) {
        if (in_run && speech_count >= min_iv) { const ns_t e = v.ts[end_idx] + hop; runs.push_back(jt_region{run_start, e, e - run_start}); }
        in_run = false; speech_count = 0; pending = 0;
    };
    for (int64_t i = 0; i < v.n; i++) {
        const jt_interval &s = v.iv[i];
        const double l = level_of(s, axis); const bool ok = veto_ok(s), sp = l >= split && ok;
        if (!in_run) { if (l >= high && ok) { run_start = v.ts[i]; speech_count = 1; last_speech = i; pending = 0; in_run = true; } continue; }
        if (sp) { speech_count++; last_speech = i; pending = 0; continue; }
        if (l >= split && !ok) { flush(last_speech); continue; }
        if (l < low) { pending++; if (pending > tol) flush(last_speech); }
    }
    if (v.n > 0) flush(last_speech);
    return runs;
}

// ---- region accumulation (analyser_candidates_shared.go:118-160) ----
struct RegionAcc { double rms = 0, peak = -120, tp = -120, sp = -120, spec[JT_SP_COUNT] = {0}, m = 0, s = 0; };
RegionAcc accumulate(const View &v, Span r)
{
    RegionAcc a;
    for (int64_t i = r.lo; i < r.hi; i++) {
        const jt_interval &x = v.iv[i];
        a.rms += x.rms_level; if (x.peak_level > a.peak) a.peak = x.peak_level;
        for (int k = 0; k < JT_SP_COUNT; k++) a.spec[k] += x.spectral[k];
        a.m += x.momentary_lufs; a.s += x.short_term_lufs;
        if (x.true_peak > a.tp) a.tp = x.true_peak;
        if (x.sample_peak > a.sp) a.sp = x.sample_peak;
    }
    return a;
}
jt_region_sample region_sample(const RegionAcc &a, double n)
{
    jt_region_sample s; memset(&s, 0, sizeof(s));
    s.rms_level = a.rms / n; s.peak_level = a.peak; s.crest_factor = a.peak - s.rms_level;
    for (int k = 0; k < JT_SP_COUNT; k++) s.spectral[k] = a.spec[k] / n;
    s.momentary_lufs = a.m / n; s.short_term_lufs = a.s / n; s.true_peak = a.tp; s.sample_peak = a.sp;
    return s;
}

double score_interval_window(const jt_interval *iv, int64_t n)                  // analyser_candidates_shared.go:165-175
{
    if (n <= 0) return 0;
    double s = 0; for (int64_t i = 0; i < n; i++) s += iv[i].rms_level;
    return s / (double)n;
}
double rolloff_score(double r)                                                    // analyser_candidates_speech.go:129-142
{
    if (r >= kRollIdealMin && r <= kRollIdealMax) return 1.0;
    if (r >= kRollOkMin && r < kRollIdealMin) return 0.5 + 0.5 * (r - kRollOkMin) / (kRollIdealMin - kRollOkMin);
    if (r > kRollIdealMax && r <= kRollOkMax) return 0.5 + 0.5 * (kRollOkMax - r) / (kRollOkMax - kRollIdealMax);
    return 0.0;
}
double flux_score(double f)                                                       // analyser_candidates_speech.go:147-165
{
    if (f <= kFluxStable) return 1.0;
    if (f <= kFluxNormal) return 1.0 - (f - kFluxStable) / (kFluxNormal - kFluxStable) * 0.3;
    if (f <= kFluxTransient) return 0.7 - (f - kFluxNormal) / (kFluxTransient - kFluxNormal) * 0.3;
    if (f <= kFluxOk) return 0.4 - (f - kFluxTransient) / (kFluxOk - kFluxTransient) * 0.2;
    return 0.2;
}
double score_speech_window(const jt_interval *iv, int64_t cnt)                   // analyser_candidates_shared.go:194-292
{
    if (cnt <= 0) return 0;
    const double n = (double)cnt;
    double ku = 0, fl = 0, ce = 0, rm = 0, ro = 0, fx = 0;
    for (int64_t i = 0; i < cnt; i++) {
        ku += iv[i].spectral[JT_SP_kurtosis]; fl += iv[i].spectral[JT_SP_flatness]; ce += iv[i].spectral[JT_SP_centroid];
        rm += iv[i].rms_level; ro += iv[i].spectral[JT_SP_rolloff]; fx += iv[i].spectral[JT_SP_flux];
    }
    const double aku = ku / n, afl = fl / n, ace = ce / n, arm = rm / n, aro = ro / n, afx = fx / n;
    double var = 0; int64_t voiced = 0;
    for (int64_t i = 0; i < cnt; i++) { const double d = iv[i].spectral[JT_SP_kurtosis] - aku; var += d * d; }
    var /= n;
    for (int64_t i = 0; i < cnt; i++) if (iv[i].spectral[JT_SP_kurtosis] > kVoicedKurtosis) voiced++;
    const double voicing = gomax(0.0, gomin(((double)voiced / n) / kVoicingDensityThr, 1.0));
    const double s_ku = gomax(0.0, gomin(aku / 7.5, 1.0));
    const double s_fl = gomax(0.0, gomin(1.0 - afl, 1.0));
    double s_ce = 0.0;
    if (ace >= kCentroidMin && ace <= kCentroidMax) {
        const double mid = (kCentroidMin + kCentroidMax) / 2, half = (kCentroidMax - kCentroidMin) / 2;
        s_ce = 1.0 - (fabs(ace - mid) / half) * 0.5;
    }
    const double s_co = gomax(0.0, gomin(1.0 - (var / 100.0), 1.0));
    double s_rm = 0.0;
    if (arm > -30.0) s_rm = gomax(0.0, gomin((arm - (-30.0)) / 18.0, 1.0));
    return s_ku * kWKurt + s_fl * kWFlat + s_ce * kWCent + s_co * kWCons + s_rm * kWRMS + voicing * kWVoic +
           rolloff_score(aro) * kWRoll + flux_score(afx) * kWFlux;
}
double level_variance(const jt_interval *iv, int64_t cnt, int axis)              // analyser_candidates_shared.go:302-321
{
    if (cnt <= 0) return 0;
    const double n = (double)cnt;
    double s = 0; for (int64_t i = 0; i < cnt; i++) s += level_of(iv[i], axis);
    const double mean = s / n;
    double v = 0; for (int64_t i = 0; i < cnt; i++) { const double d = level_of(iv[i], axis) - mean; v += d * d; }
    return v / n;
}

// refineToSubregion (analyser_candidates_shared.go:29-81)
bool refine(const View &v, jt_region &reg, ns_t window, ns_t minimum, double (*score)(const jt_interval *, int64_t), bool higher_wins)
{
    if (reg.duration_ns <= window) return false;
    const Span r = in_range(v, reg.start_ns, reg.end_ns);
    if (r.len() == 0) return false;
    int64_t w = window / kGoldenInterval; const int64_t mn = minimum / kGoldenInterval;
    if (r.len() < mn) return false;
    if (r.len() < w) w = r.len();
    int64_t best = 0; double best_score = score(v.iv + r.lo, w);
    for (int64_t s = 1; s <= r.len() - w; s++) {
        const double sc = score(v.iv + r.lo + s, w);
        if (higher_wins ? sc > best_score : sc < best_score) { best_score = sc; best = s; }
    }
    const ns_t st = v.ts[r.lo + best], du = w * kGoldenInterval;
    reg = jt_region{st, st + du, du};
    return true;
}

// measureSpeechCandidateFromIntervals (analyser_candidates_shared.go:326-365)
bool measure_candidate(const View &v, const jt_region &reg, jt_speech_candidate &out)
{
    const Span r = in_range(v, reg.start_ns, reg.end_ns);
    if (r.len() == 0) return false;
    const double n = (double)r.len();
    memset(&out, 0, sizeof(out));
    out.region = reg;
    out.sample = region_sample(accumulate(v, r), n);
    int64_t voiced = 0;
    for (int64_t i = r.lo; i < r.hi; i++) if (v.iv[i].spectral[JT_SP_kurtosis] > kVoicedKurtosis) voiced++;
    out.voicing_density = (double)voiced / n;
    return true;
}

// scoreSpeechCandidateGrounded and its three terms (analyser_candidates_speech.go:381-433)
double grounded_score(const jt_speech_candidate &m, double floor_db, double level_var)
{
    const double snr = m.sample.rms_level - floor_db;
    double s_snr;
    if (snr <= 0) s_snr = 0.0;
    else if (snr < kMinSNR) s_snr = 0.5 * (snr / kMinSNR);
    else if (snr >= kSNRSaturation) s_snr = 1.0;
    else s_snr = 0.5 + 0.5 * (snr - kMinSNR) / (kSNRSaturation - kMinSNR);
    const double s_dur = m.region.duration_ns >= kGoldenSpeechMin ? 1.0
                         : gomax(0.0, gomin(dur_seconds(m.region.duration_ns) / dur_seconds(kGoldenSpeechMin), 1.0));
    const double tie = gomax(0.0, gomin(1.0 - (level_var / kGroundVarCap), 1.0)) * kGroundTieMax;
    return s_snr * kGroundSNRW + s_dur * kGroundDurW + tie;
}

// findBestSpeechRegion (analyser_candidates_speech.go:222-326)
bool find_best(const View &v, const std::vector<jt_region> &regions, double floor_db, jt_region &best_out, std::vector<jt_speech_candidate> &cands)
{
    cands.clear();
    if (regions.empty()) return false;
    bool have_best = false, have_fb = false; jt_region best{}, fb{}; double best_score = 0, fb_score = 0;
    for (const jt_region &cand : regions) {
        jt_speech_candidate m;
        if (!measure_candidate(v, cand, m)) continue;
        const Span r = in_range(v, cand.start_ns, cand.end_ns);
        m.score = grounded_score(m, floor_db, level_variance(v.iv + r.lo, r.len(), 0));
        cands.push_back(m);
        if (!have_fb || m.score > fb_score) { fb = m.region; fb_score = m.score; have_fb = true; }
        if (m.score >= kMinViableScore && (!have_best || m.score > best_score)) { best = cand; best_score = m.score; have_best = true; }
    }
    if (!have_best && have_fb) { best = fb; have_best = true; }
    if (have_best && best.duration_ns > kGoldenSpeechWindow) {
        const jt_region orig = best;
        jt_region refined = best;
        refine(v, refined, kGoldenSpeechWindow, kGoldenSpeechMin, score_speech_window, true);
        if (refined.start_ns != orig.start_ns || refined.duration_ns != orig.duration_ns) {
            jt_speech_candidate rm;
            if (measure_candidate(v, refined, rm)) {
                const Span r = in_range(v, refined.start_ns, refined.end_ns);
                rm.score = grounded_score(rm, floor_db, level_variance(v.iv + r.lo, r.len(), 0));
                rm.was_refined = 1; rm.original_start_ns = orig.start_ns; rm.original_duration_ns = orig.duration_ns;
                for (jt_speech_candidate &c : cands) if (c.region.start_ns == orig.start_ns) { c = rm; break; }
                best = refined;
            }
        }
    }
    if (have_best) best_out = best;
    return have_best;
}

// pickLowClusterRegion (analyser_vad.go:618-671)
bool low_cluster_region(const View &v, double split, int axis, ns_t hop, jt_region &out)
{
    bool have = false, in_run = false; jt_region best{}; ns_t run_start = 0;
    auto close = [&](int64_t end_idx) {
        if (!in_run) return;
        const ns_t e = v.ts[end_idx] + hop; const jt_region r{run_start, e, e - run_start};
        if (!have || r.duration_ns > best.duration_ns) { best = r; have = true; }
        in_run = false;
    };
    for (int64_t i = 0; i < v.n; i++) {
        if (level_of(v.iv[i], axis) < split) { if (!in_run) { run_start = v.ts[i]; in_run = true; } continue; }
        if (in_run) close(i - 1);
    }
    if (in_run) close(v.n - 1);
    if (!have) return false;
    out = best;
    refine(v, out, kGoldenWindow, kGoldenMin, score_interval_window, false);
    return true;
}

// extractNoiseProfileFromIntervals (analyser_vad.go:567-606)
bool noise_profile(const View &v, const jt_region &reg, jt_noise_profile &p)
{
    const Span r = in_range(v, reg.start_ns, reg.start_ns + reg.duration_ns);
    if (r.len() == 0) return false;
    const RegionAcc a = accumulate(v, r); const double n = (double)r.len(), rms = a.rms / n;
    memset(&p, 0, sizeof(p));
    p.start_ns = reg.start_ns; p.duration_ns = reg.duration_ns;
    p.measured_noise_floor = rms; p.peak_level = a.peak; p.crest_factor = a.peak - rms;
    for (int k = 0; k < JT_SP_COUNT; k++) p.spectral[k] = a.spec[k] / n;
    p.entropy = p.spectral[JT_SP_entropy];
    p.warning = reg.duration_ns < kIdealMin ? 1 : (reg.duration_ns > kIdealMax ? 2 : 0);
    return true;
}

// deriveGateStatistics (analyser_vad.go:217-253)
void gate_statistics(const View &v, double split, int axis, const jt_region *speech, double &voiced_low, double &noise_high, double &sep)
{
    std::vector<double> voiced, noise;
    for (int64_t i = 0; i < v.n; i++) { const double l = level_of(v.iv[i], axis); if (is_floored(l)) continue; if (l < split) noise.push_back(l); }
    if (speech) {
        const Span r = in_range(v, speech->start_ns, speech->end_ns);
        for (int64_t i = r.lo; i < r.hi; i++) if (is_speech(v.iv[i], split, axis)) voiced.push_back(level_of(v.iv[i], axis));
    }
    voiced_low = go_pct(voiced, kGateVoicedLowPct); noise_high = go_pct(noise, kGateNoiseHighPct); sep = voiced_low - noise_high;
}

double floored_fraction(const View &v, int axis)                                  // analyser_vad.go:683-697
{
    double counted = 0, fl = 0;
    for (int64_t i = 0; i < v.n; i++) { const double l = level_of(v.iv[i], axis); counted++; if (std::isnan(l) || l <= kLevelFloorDB) fl++; }
    return counted == 0 ? 0 : fl / counted;
}

// computeSilenceMedians + roomToneScore + estimateNoiseFloorAndThreshold (analyser_noise_seed.go:78-223)
bool estimate_noise_floor(const View &v, double &floor, double &thr)
{
    floor = 0; thr = 0;
    if (v.n < kSilenceMinIntervals) return false;
    std::vector<double> lv((size_t)v.n), fx((size_t)v.n);
    for (int64_t i = 0; i < v.n; i++) { lv[i] = v.iv[i].momentary_lufs; fx[i] = v.iv[i].spectral[JT_SP_flux]; }
    const double l50 = go_kth(lv, v.n / 2), f50 = go_kth(fx, v.n / 2);
    struct Scored { int64_t idx; double level, score; };
    std::vector<Scored> sc((size_t)v.n);
    for (int64_t i = 0; i < v.n; i++) {
        const jt_interval &x = v.iv[i];
        double amp = 1.0;
        if (x.momentary_lufs > l50) { amp = 1.0 - (x.momentary_lufs - l50) / kRoomAmpDecayDB; if (amp < 0) amp = 0; }
        double flx = 1.0;
        if (f50 > 0 && x.spectral[JT_SP_flux] > f50) { const double ratio = x.spectral[JT_SP_flux] / f50; if (ratio > 1) flx = 1.0 / ratio; }
        sc[i] = Scored{i, x.momentary_lufs, kRoomAmpW * amp + kRoomFluxW * flx};
    }
    int64_t cnt = v.n / kSeedTopDiv; cnt = std::max<int64_t>(cnt, kSeedMinCount); cnt = std::min<int64_t>(cnt, v.n);
    // the order is total (index breaks ties), so the first cnt elements of the sorted slice are a well-defined SET, and only
    // their maximum level is read: a selection instead of the sort
    auto before = [](const Scored &a, const Scored &b) {
        int c = go_cmp(b.score, a.score); if (c) return c < 0;
        c = go_cmp(a.level, b.level); if (c) return c < 0;
        return a.idx < b.idx;
    };
    if (cnt < v.n) std::nth_element(sc.begin(), sc.begin() + cnt, sc.end(), before);
    double mx = -120.0; bool seen = false;
    for (int64_t i = 0; i < cnt; i++) { const double l = sc[i].level; if (is_floored(l)) continue; if (!seen || l > mx) { mx = l; seen = true; } }
    if (!seen) return false;
    floor = mx; thr = mx + kSilenceHeadroomDB;
    return true;
}
double adaptive_silence_threshold(double floor)                                   // analyser_noise_seed.go:228-241
{
    double t = floor + kSilenceFallbackHeadroom;
    if (t < kSilenceMinThr) t = kSilenceMinThr;
    if (t > kSilenceMaxThr) t = kSilenceMaxThr;
    return t;
}

int copy_out(const std::string &s, char *buf, size_t cap)
{
    if (!buf || s.size() + 1 > cap) return JT_ERR_BUFFER;
    memcpy(buf, s.c_str(), s.size() + 1);
    return JT_OK;
}
std::string fmt(const char *f, ...) __attribute__((format(printf, 1, 2)));
std::string fmt(const char *f, ...)
{
    char b[512]; va_list ap; va_start(ap, f); vsnprintf(b, sizeof(b), f, ap); va_end(ap); return b;
}

// fmt.Sprintf("%g", v): strconv.FormatFloat(v, 'g', -1, 64) -- shortest digits that round-trip, %e form when the decimal
// exponent is < -4 or >= 6 (the shortest form decides with precision 6), exponent with at least two digits
std::string go_g(double v)
{
    if (std::isnan(v)) return "NaN";
    if (std::isinf(v)) return v > 0 ? "+Inf" : "-Inf";
    if (v == 0) return std::signbit(v) ? "-0" : "0";
    char e[40]; int p = 1;
    for (; p <= 17; p++) { snprintf(e, sizeof(e), "%.*e", p - 1, v); if (strtod(e, nullptr) == v) break; }
    // e = [-]d[.ddd]e[+-]XX ; digits and exponent
    std::string digits; const char *q = e; bool neg = false;
    if (*q == '-') { neg = true; q++; }
    for (; *q && *q != 'e'; q++) if (*q != '.') digits.push_back(*q);
    const int x = atoi(q + 1);
    while (digits.size() > 1 && digits.back() == '0') digits.pop_back();
    const int nd = (int)digits.size();
    std::string out = neg ? "-" : "";
    if (x < -4 || x >= 6) {
        out += digits[0];
        if (nd > 1) { out += '.'; out += digits.substr(1); }
        char xb[16]; snprintf(xb, sizeof(xb), "e%c%02d", x < 0 ? '-' : '+', abs(x));
        return out + xb;
    }
    if (x < 0) { out += "0."; out.append((size_t)(-x - 1), '0'); out += digits; return out; }
    if (nd <= x + 1) { out += digits; out.append((size_t)(x + 1 - nd), '0'); return out; }
    out += digits.substr(0, (size_t)x + 1); out += '.'; out += digits.substr((size_t)x + 1);
    return out;
}

// ---- defaults (filters.go:421-532) ----
void set_str(char *dst, size_t cap, const char *s) { memset(dst, 0, cap); strncpy(dst, s, cap - 1); }
jt_biquad_config default_biquad(double f)
{
    jt_biquad_config b; memset(&b, 0, sizeof(b));
    b.enabled = 1; b.frequency = f; b.poles = 2; b.width = 0.707; b.mix = 1.0; set_str(b.transform, sizeof(b.transform), "tdii");
    return b;
}
void defaults(jt_filter_config &c)
{
    memset(&c, 0, sizeof(c));
    c.downmix_enabled = 1; c.analysis_enabled = 1;
    c.resample_enabled = 1; c.resample_rate = 44100; c.resample_frame_size = 4096; set_str(c.resample_format, sizeof(c.resample_format), "s16");
    c.rumble_highpass = default_biquad(80.0); c.bandlimit_lowpass = default_biquad(20500.0);
    auto &nr = c.noise_reduction;
    nr.enabled = 1; nr.strength = 0.00001; nr.patch_s = 0.0060; nr.research_s = 0.0020; nr.smooth = 3.0;
    nr.afftdn_enabled = 1; nr.afftdn_noise_reduction = 12; set_str(nr.afftdn_noise_type, sizeof(nr.afftdn_noise_type), "w"); nr.afftdn_track_noise = 1;
    auto &g = c.speech_gate;
    g.enabled = 1; g.threshold = 0.01; g.ratio = 2.0; g.attack = 5.0; g.release = 200.0; g.range = db_to_linear(-14.0); g.knee = 3.0; g.makeup = 1.0;
    set_str(g.detection, sizeof(g.detection), "rms");
    auto &k = c.levelling_compressor;
    k.enabled = 1; k.threshold = -18; k.ratio = 3.0; k.attack = 10; k.release = 200; k.makeup = 0; k.knee = 4.0; k.mix = 1.0;
    c.deesser.enabled = 1; c.deesser.intensity = 0.0; c.deesser.amount = 0.50; c.deesser.frequency = 0.80;
    c.adeclick.enabled = 1; c.adeclick.threshold = 1.7; c.adeclick.window = 55.0; c.adeclick.overlap = 50.0; set_str(c.adeclick.method, sizeof(c.adeclick.method), "s");
    c.loudnorm.enabled = 1; c.loudnorm.target_i = -16.0; c.loudnorm.target_tp = -1.0; c.loudnorm.target_lra = 20.0; c.loudnorm.dual_mono = 1; c.loudnorm.linear = 1;
    c.n_filter_order = 0;
}

// ---- filter builders (filters.go:607-960) ----
std::string build_biquad(const jt_biquad_config &b, const char *kw)
{
    if (!b.enabled) return "";
    const int poles = b.poles < 1 ? 2 : b.poles;
    const double width = b.width <= 0 ? 0.707 : b.width;
    std::string s = fmt("%s=f=%.0f:poles=%d:width_type=q:width=%.3f:normalize=1", kw, b.frequency, poles, width);
    if (b.transform[0]) s += fmt(":a=%s", b.transform);
    if (b.mix > 0 && b.mix < 1.0) s += fmt(":m=%.2f", b.mix);
    return s;
}
std::string build_afftdn(const jt_filter_config &c)
{
    const auto &n = c.noise_reduction;
    if (!n.afftdn_enabled) return "";
    const int tn = n.afftdn_track_noise ? 1 : 0;
    std::string s;
    if (!strcmp(n.afftdn_noise_type, "custom") && n.afftdn_band_noise[0])
        s = "afftdn=nr=" + go_g(n.afftdn_noise_reduction) + ":nt=custom:bn=" + n.afftdn_band_noise + fmt(":tn=%d", tn);
    else
        s = "afftdn=nr=" + go_g(n.afftdn_noise_reduction) + ":nt=" + n.afftdn_noise_type + fmt(":tn=%d", tn);
    if (n.afftdn_noise_floor < 0) s += ":nf=" + go_g(n.afftdn_noise_floor);
    return s;
}
std::string build_one(const jt_filter_config &c, int id)
{
    switch (id) {
    case JT_FILTER_DOWNMIX: return c.downmix_enabled ? "aformat=channel_layouts=mono" : "";
    case JT_FILTER_ANALYSIS:
        if (!c.analysis_enabled) return "";
        return fmt("astats=metadata=1:measure_perchannel=all,aspectralstats=win_size=2048:win_func=hann:measure=all,"
                   "ebur128=metadata=1:peak=sample+true:dualmono=true:target=%.0f", c.loudnorm.target_i);
    case JT_FILTER_RESAMPLE:
        if (!c.resample_enabled) return "";
        return fmt("aformat=sample_rates=%d:channel_layouts=mono:sample_fmts=%s,asetnsamples=n=%d", c.resample_rate, c.resample_format, c.resample_frame_size);
    case JT_FILTER_RUMBLE_HIGHPASS: return build_biquad(c.rumble_highpass, "highpass");
    case JT_FILTER_BANDLIMIT_LOWPASS: return build_biquad(c.bandlimit_lowpass, "lowpass");
    case JT_FILTER_NOISE_REDUCTION: {
        const auto &n = c.noise_reduction;
        if (!n.enabled) return "";
        std::string s = fmt("anlmdn=s=%.5f:p=%.4f:r=%.4f:m=%.0f", n.strength, n.patch_s, n.research_s, n.smooth);
        const std::string a = build_afftdn(c);
        return a.empty() ? s : s + "," + a;
    }
    case JT_FILTER_SPEECH_GATE: {
        const auto &g = c.speech_gate;
        if (!g.enabled) return "";
        return fmt("agate=threshold=%.6f:ratio=%.1f:attack=%.2f:release=%.0f:range=%.4f:knee=%.1f:detection=%s:makeup=%.1f",
                   g.threshold, g.ratio, g.attack, g.release, g.range, g.knee, g.detection[0] ? g.detection : "rms", g.makeup);
    }
    case JT_FILTER_LEVELLING_COMPRESSOR: {
        const auto &k = c.levelling_compressor;
        if (!k.enabled) return "";
        return fmt("acompressor=threshold=%.6f:ratio=%.1f:attack=%.0f:release=%.0f:makeup=%.2f:knee=%.1f:detection=rms:mix=%.2f",
                   db_to_linear(k.threshold), k.ratio, k.attack, k.release, db_to_linear(k.makeup), k.knee, k.mix);
    }
    case JT_FILTER_DEESSER:
        if (!c.deesser.enabled || c.deesser.intensity <= 0) return "";
        return fmt("deesser=i=%.2f:m=%.2f:f=%.2f", c.deesser.intensity, c.deesser.amount, c.deesser.frequency);
    }
    return "";
}
const int kPass2Order[] = { JT_FILTER_DOWNMIX, JT_FILTER_RUMBLE_HIGHPASS, JT_FILTER_BANDLIMIT_LOWPASS, JT_FILTER_NOISE_REDUCTION,
                            JT_FILTER_SPEECH_GATE, JT_FILTER_LEVELLING_COMPRESSOR, JT_FILTER_DEESSER, JT_FILTER_ANALYSIS, JT_FILTER_RESAMPLE };
std::string build_spec(const jt_filter_config &c)
{
    const int *order = kPass2Order; int n = (int)(sizeof(kPass2Order) / sizeof(int));
    if (c.n_filter_order > 0) { order = c.filter_order; n = std::min<int>(c.n_filter_order, 12); }
    std::string s;
    for (int i = 0; i < n; i++) { const std::string f = build_one(c, order[i]); if (f.empty()) continue; if (!s.empty()) s += ","; s += f; }
    return s;
}

// ---- AdaptConfig pieces (adaptive*.go) ----
std::string band_noise(const double *bands, int n)                               // adaptive.go:83-111
{
    if (n <= 0) return "";
    double sum = 0; int fin = 0;
    for (int i = 0; i < n; i++) if (is_finite(bands[i])) { sum += bands[i]; fin++; }
    if (!fin) return "";
    const double mean = sum / (double)fin;
    std::string s;
    for (int i = 0; i < n; i++) {
        if (i) s += "|";
        if (!is_finite(bands[i])) { s += "0.0"; continue; }
        s += fmt("%.1f", gomax(-24.0, gomin(24.0, bands[i] - mean)));
    }
    return s;
}
double gate_threshold(double voiced_low, double sep, int *narrow)                // adaptive_speech_gate.go:258-270
{
    double t = voiced_low - 6.0;
    if (narrow) *narrow = sep < (6.0 + 6.0);
    t = gomax(-80.0, gomin(t, -25.0));
    return db_to_linear(t);
}
double gate_threshold_no_profile(double floor, double peak, double crest, double ratio, double gap)   // adaptive_speech_gate.go:216-234
{
    double t;
    if (crest > 20.0 && peak != 0 && gap < 25.0) t = peak + 3.0;
    else t = gomax(floor + 12.0 / (1.0 - 1.0 / ratio), -40.0);
    t = gomax(-80.0, gomin(t, -25.0));
    return db_to_linear(t);
}

struct Meas { double input_i, input_lra, rms_level, peak_level; };
Meas go_meas(const jt_measurements *m)
{
    // Dynamics.* stay at Go's zero value unless astats reported (assignAstatsMeasurements analyser.go:417-443)
    Meas g; g.input_i = m->input_i; g.input_lra = m->input_lra;
    const bool found = !std::isnan(m->astats[JT_AS_Dynamic_range]);
    // (astats writes all its keys or none, so a NaN next to a reported Dynamic_range is a parsed "nan", not an absent key)
    g.rms_level = found ? m->astats[JT_AS_RMS_level] : 0.0;
    g.peak_level = found ? m->astats[JT_AS_Peak_level] : 0.0;
    return g;
}

// sanitizeConfig (adaptive.go:175-232)
void sanitize_config(jt_filter_config &c)
{
    auto &nr = c.noise_reduction; auto &gt = c.speech_gate; auto &k = c.levelling_compressor;
    jt_filter_config def; defaults(def);
    auto san_bq = [](jt_biquad_config &b, double f) { b.frequency = sanitize(b.frequency, f); b.width = sanitize(b.width, 0.707); b.mix = sanitize(b.mix, 1.0); };
    san_bq(c.rumble_highpass, 80.0); san_bq(c.bandlimit_lowpass, 20500.0);
    nr.strength = sanitize(nr.strength, def.noise_reduction.strength); nr.patch_s = sanitize(nr.patch_s, def.noise_reduction.patch_s);
    nr.research_s = sanitize(nr.research_s, def.noise_reduction.research_s); nr.smooth = sanitize(nr.smooth, def.noise_reduction.smooth);
    nr.afftdn_noise_reduction = sanitize(nr.afftdn_noise_reduction, def.noise_reduction.afftdn_noise_reduction);
    nr.afftdn_noise_floor = sanitize(nr.afftdn_noise_floor, def.noise_reduction.afftdn_noise_floor);
    if (!strcmp(nr.afftdn_noise_type, "custom") && !nr.afftdn_band_noise[0]) set_str(nr.afftdn_noise_type, sizeof(nr.afftdn_noise_type), "w");
    if (!is_finite(gt.threshold) || gt.threshold <= 0) gt.threshold = 0.01;
    gt.ratio = sanitize(gt.ratio, def.speech_gate.ratio); gt.attack = sanitize(gt.attack, def.speech_gate.attack);
    gt.release = sanitize(gt.release, def.speech_gate.release); gt.range = sanitize(gt.range, def.speech_gate.range);
    gt.knee = sanitize(gt.knee, def.speech_gate.knee); gt.makeup = sanitize(gt.makeup, def.speech_gate.makeup);
    k.ratio = sanitize(k.ratio, 3.0); k.threshold = sanitize(k.threshold, -18.0); k.attack = sanitize(k.attack, 10); k.release = sanitize(k.release, 200);
    k.makeup = sanitize(k.makeup, 0); k.knee = sanitize(k.knee, 4.0); k.mix = sanitize(k.mix, 1.0);
    c.deesser.intensity = sanitize(c.deesser.intensity, 0.0); c.deesser.amount = sanitize(c.deesser.amount, 0.50); c.deesser.frequency = sanitize(c.deesser.frequency, 0.80);
}

void adapt(const jt_filter_config &base, const jt_measurements *m, const jt_voice_activity *va, jt_filter_config &c, jt_adapt_diagnostics &d)
{
    c = base; memset(&d, 0, sizeof(d));
    const Meas g = go_meas(m);
    // tuneBandlimitLowPass (adaptive_bandlimit_lowpass.go:23-31)
    c.bandlimit_lowpass.enabled = 1; c.bandlimit_lowpass.frequency = 20500.0; c.bandlimit_lowpass.poles = 2; c.bandlimit_lowpass.mix = 1.0;
    set_str(d.bandlimit_lp_reason, sizeof(d.bandlimit_lp_reason), "20.5 kHz band-limit (always on)");
    // tuneNoiseReduction (adaptive.go:135-172)
    auto &nr = c.noise_reduction;
    if (va->voice_activated) {
        nr.afftdn_enabled = 0; d.afftdn_enabled = 0; set_str(d.afftdn_disable_reason, sizeof(d.afftdn_disable_reason), "voice_activated");
    } else {
        d.afftdn_enabled = nr.afftdn_enabled;
        if (va->floor != 0) {
            const double fl = gomax(-80.0, gomin(-20.0, va->floor));
            nr.afftdn_noise_floor = fl; nr.afftdn_track_noise = 0; d.afftdn_noise_floor_db = fl;
            set_str(nr.afftdn_noise_type, sizeof(nr.afftdn_noise_type), "w");
            const bool custom = va->has_noise_profile && va->noise_profile.bands_measured && !(va->gate_separation_db < 12.0) &&
                                va->noise_profile.spectral[JT_SP_flatness] >= 0.45;             // useCustomAfftdnProfile adaptive.go:117-126
            if (custom) {
                const std::string bn = band_noise(va->noise_profile.band_noise, std::min<int>(va->noise_profile.n_band_noise, JT_AFFTDN_BANDS));
                if (!bn.empty()) { set_str(nr.afftdn_noise_type, sizeof(nr.afftdn_noise_type), "custom"); set_str(nr.afftdn_band_noise, sizeof(nr.afftdn_band_noise), bn.c_str()); }
            }
            set_str(d.afftdn_noise_type, sizeof(d.afftdn_noise_type), nr.afftdn_noise_type);
        }
    }
    // tuneSpeechGate (adaptive_speech_gate.go:95-181)
    auto &gt = c.speech_gate;
    double crest = 15.0, peak = 0;
    if (va->has_noise_profile) { crest = va->noise_profile.crest_factor; peak = va->noise_profile.peak_level; }
    double gap = c.loudnorm.target_i - g.input_i; if (gap < 0) gap = 0;
    gt.ratio = g.input_lra > 15.0 ? 1.5 : 2.0;
    int narrow = 0;
    if (va->has_speech_profile) {
        gt.threshold = gate_threshold(va->voiced_low_percentile, va->gate_separation_db, &narrow);
        const double actual = linear_to_db(gt.threshold);
        d.speech_gate_narrow_gap = narrow; d.speech_gate_quiet_speech_estimate = va->voiced_low_percentile;
        d.speech_gate_speech_separation = va->gate_separation_db; d.speech_gate_threshold_unclamped = va->voiced_low_percentile - 6.0;
        d.speech_gate_speech_headroom = va->voiced_low_percentile - actual;
        set_str(d.speech_gate_clamp_reason, sizeof(d.speech_gate_clamp_reason), narrow ? "narrow_gap" : "none");
    } else {
        gt.threshold = gate_threshold_no_profile(va->floor, peak, crest, gt.ratio, gap);
    }
    gt.attack = 5.0; gt.release = 200.0;
    const double depth = narrow ? 8.0 : 14.0;
    gt.range = db_to_linear(-depth); d.speech_gate_depth_db = depth;
    gt.knee = 3.0; set_str(gt.detection, sizeof(gt.detection), "rms");
    // tuneDeesser (adaptive_deesser.go:45-67)
    if (!va->has_speech_profile || !va->speech_profile.bands_measured) c.deesser.intensity = 0.0;
    else {
        const double ex = va->speech_profile.sib_band_rms - va->speech_profile.body_band_rms;
        if (ex < -6.0) c.deesser.intensity = 0.0;
        else if (ex < -3.0) c.deesser.intensity = (ex - (-6.0)) / (-3.0 - (-6.0)) * 0.6;
        else if (ex < 0.0) c.deesser.intensity = 0.6 + (ex - (-3.0)) / (0.0 - (-3.0)) * (0.85 - 0.6);
        else c.deesser.intensity = 0.85;
    }
    // tuneLevellingCompressor (adaptive_levelling_compressor.go:55-101)
    auto &k = c.levelling_compressor;
    k.ratio = 3.0; k.attack = 10.0; k.release = 200.0; k.knee = 4.0; k.mix = 1.0; k.makeup = 0.0;
    if (va->has_speech_profile) {
        double rms = va->speech_profile.sample.rms_level;
        if (g.rms_level < 0 && !(std::isinf(g.rms_level) && g.rms_level < 0)) rms = gomax(rms, g.rms_level);
        k.threshold = gomax(-45.0, gomin(rms + 9.0, -6.0));
    } else if (!is_finite(g.peak_level)) k.threshold = -18.0;
    else k.threshold = gomax(-45.0, gomin(g.peak_level - 20.0, -6.0));
    sanitize_config(c);
}

}  // namespace

// =====================================================================================================================
// C ABI
// =====================================================================================================================
// detectVoiceActivity (analyser_vad.go:728-783); axis = momentary LUFS, hop = 250 ms.  Fills everything but the seed fields.
static void detect(const View &v, double seed, jt_voice_activity &va, std::vector<jt_region> &runs, std::vector<jt_speech_candidate> &cands)
{
    const int axis = 0; const ns_t hop = kIntervalHop;
    const Hist h = build_hist(v, axis, 1.0);
    std::vector<double> levels; levels.reserve((size_t)v.n);           // vad_levels without the sort: two order statistics are read
    for (int64_t i = 0; i < v.n; i++) { const double l = level_of(v.iv[i], axis); if (!is_floored(l)) levels.push_back(l); }
    const double p75 = go_pct(levels, 75);
    const double split = clamp_split(otsu(h), seed, p75);
    const double floor = gomax(go_pct(levels, kNoiseFloorPct), seed + kNoiseMarginDB);        // percentile_floor
    std::vector<uint8_t> flags((size_t)v.n);
    for (int64_t i = 0; i < v.n; i++) flags[i] = is_speech(v.iv[i], split, axis);
    const double margin = hysteresis_margin(h, split);
    const int tol = gap_tolerance(flags.data(), v.n, hop);
    runs = speech_runs(v, split, margin, tol, axis, hop);
    va.split = split; va.margin = margin; va.gap_tolerance = tol; va.n_speech_regions = (int64_t)runs.size();

    jt_region nreg;
    if (low_cluster_region(v, split, axis, hop, nreg) && noise_profile(v, nreg, va.noise_profile)) {
        va.noise_profile.measured_noise_floor = floor;
        va.has_noise_profile = 1; va.noise_region = nreg;
        const Span r = in_range(v, nreg.start_ns, nreg.start_ns + nreg.duration_ns);      // setVADRoomToneSample
        if (r.len() > 0) { va.room_tone_sample = region_sample(accumulate(v, r), (double)r.len()); va.has_room_tone_sample = 1; }
    }
    jt_region best;
    const bool elected = find_best(v, runs, va.has_noise_profile ? va.noise_profile.measured_noise_floor : -INFINITY, best, cands);
    va.n_candidates = (int64_t)cands.size();
    va.speech_profile_index = -1;
    if (elected)
        for (size_t i = 0; i < cands.size(); i++) if (cands[i].region.start_ns == best.start_ns) { va.speech_profile = cands[i]; va.has_speech_profile = 1; va.speech_profile_index = (int32_t)i; break; }
    gate_statistics(v, split, axis, va.has_speech_profile ? &va.speech_profile.region : nullptr,
                    va.voiced_low_percentile, va.noise_high_percentile, va.gate_separation_db);
    va.floor = floor; va.floor_source = JT_FLOOR_VAD_PERCENTILE;
    va.floored_fraction = floored_fraction(v, axis);
    va.voice_activated = va.floored_fraction >= kVoiceActivatedFraction;
}

static int emit(const jt_voice_activity &va, const std::vector<jt_region> &runs, const std::vector<jt_speech_candidate> &cands, jt_voice_activity *out,
                jt_region *regions_out, int64_t regions_cap, jt_speech_candidate *cands_out, int64_t cands_cap)
{
    if (regions_out) { if ((int64_t)runs.size() > regions_cap) return JT_ERR_BUFFER; for (size_t i = 0; i < runs.size(); i++) regions_out[i] = runs[i]; }
    if (cands_out) { if ((int64_t)cands.size() > cands_cap) return JT_ERR_BUFFER; for (size_t i = 0; i < cands.size(); i++) cands_out[i] = cands[i]; }
    *out = va;
    return JT_OK;
}

extern "C" int jt_vad_detect(const jt_interval *intervals, int64_t n, double noise_floor_seed, jt_voice_activity *out,
                             jt_region *regions_out, int64_t regions_cap, jt_speech_candidate *cands_out, int64_t cands_cap)
{
    if (!out || (n > 0 && !intervals) || n < 0) return JT_ERR_INVALID_ARG;
    const View v(intervals, n);
    jt_voice_activity va; memset(&va, 0, sizeof(va));
    std::vector<jt_region> runs; std::vector<jt_speech_candidate> cands;
    va.floor_prescan = noise_floor_seed;
    detect(v, noise_floor_seed, va, runs, cands);
    return emit(va, runs, cands, out, regions_out, regions_cap, cands_out, cands_cap);
}

extern "C" int jt_detect_voice_activity(const jt_measurements *m, const jt_interval *intervals, int64_t n, jt_voice_activity *out,
                                        jt_region *regions_out, int64_t regions_cap, jt_speech_candidate *cands_out, int64_t cands_cap)
{
    if (!m || !out || (n > 0 && !intervals) || n < 0) return JT_ERR_INVALID_ARG;
    const View v(intervals, n);
    jt_voice_activity va; memset(&va, 0, sizeof(va));
    // buildInputMeasurements: pre-scan seed (analyser.go:374-395)
    double seed, thr;
    if (!estimate_noise_floor(v, seed, thr)) { seed = kLevelFloorDB; thr = adaptive_silence_threshold(kLevelFloorDB); }
    va.floor_prescan = seed; va.room_tone_detect_level = thr;
    std::vector<jt_region> runs; std::vector<jt_speech_candidate> cands;
    detect(v, seed, va, runs, cands);
    jt_vad_assign_astats(m, &va);
    return emit(va, runs, cands, out, regions_out, regions_cap, cands_out, cands_cap);
}

// The two values of jt_detect_voice_activity that read astats' whole-file statistics (nothing in the detector itself does): the
// adaptive driver runs the detector while Pass 1's astats is still on the GPU and calls this once its values are in `m`.
extern "C" void jt_vad_assign_astats(const jt_measurements *m, jt_voice_activity *va)
{
    if (!m || !va) return;
    const Meas g = go_meas(m);
    const bool astats_found = !std::isnan(m->astats[JT_AS_Dynamic_range]);
    va->floor_astats = astats_found && !std::isnan(m->astats[JT_AS_Noise_floor]) ? m->astats[JT_AS_Noise_floor] : 0.0;
    // assignInputMeasurementSuggestions (analyser.go:515-531)
    if (g.rms_level != 0 && va->floor != 0) va->reduction_headroom = gomax(0, gomin(60, g.rms_level - va->floor));
    else va->reduction_headroom = m->input_i > -20.0 ? 40.0 : (m->input_i > -30.0 ? 25.0 : 15.0);
}

extern "C" void jt_band_plan(double lo[17], double hi[17])
{
    static const double c[JT_AFFTDN_BANDS] = {80, 125, 195, 290, 440, 660, 1000, 1500, 2250, 3350, 5000, 7500, 11200, 16000, 24000};
    lo[0] = 1000.0; hi[0] = 3000.0; lo[1] = 6000.0; hi[1] = 9000.0;
    const int last = JT_AFFTDN_BANDS - 1;
    for (int i = 0; i <= last; i++) {
        lo[2 + i] = i == 0 ? c[0] / sqrt(c[1] / c[0]) : sqrt(c[i - 1] * c[i]);
        hi[2 + i] = i == last ? c[last] * sqrt(c[last] / c[last - 1]) : sqrt(c[i] * c[i + 1]);
    }
}

extern "C" int jt_apply_band_rms(jt_voice_activity *va, const double *srms, const int32_t *sfound, const double *nrms, const int32_t *nfound)
{
    if (!va) return JT_ERR_INVALID_ARG;
    if (srms && sfound && va->has_speech_profile && va->speech_profile.region.duration_ns > 0) {       // measureSpeechBands
        if (sfound[0]) va->speech_profile.body_band_rms = srms[0];
        if (sfound[1]) va->speech_profile.sib_band_rms = srms[1];
        va->speech_profile.bands_measured = sfound[0] && sfound[1];
    }
    if (nrms && nfound && va->has_noise_profile && va->noise_profile.duration_ns > 0) {                 // measureNoiseBands
        int fin = 0;
        for (int i = 0; i < JT_AFFTDN_BANDS; i++) {
            va->noise_profile.band_noise[i] = nfound[i] ? nrms[i] : 0.0;       // an unmeasured band keeps Go's zero value
            if (nfound[i] && is_finite(nrms[i])) fin++;
        }
        va->noise_profile.n_band_noise = JT_AFFTDN_BANDS;
        va->noise_profile.bands_measured = fin >= 10;
    }
    return JT_OK;
}

extern "C" void jt_default_filter_config(jt_filter_config *c) { if (c) defaults(*c); }

extern "C" int jt_adapt_config(const jt_filter_config *base, const jt_measurements *m, const jt_voice_activity *va, jt_filter_config *out, jt_adapt_diagnostics *diag)
{
    if (!m || !va || !out) return JT_ERR_INVALID_ARG;
    jt_filter_config b; if (base) b = *base; else defaults(b);
    jt_filter_config c; jt_adapt_diagnostics d;
    adapt(b, m, va, c, d);
    *out = c; if (diag) *diag = d;
    return JT_OK;
}
extern "C" void jt_sanitize_config(jt_filter_config *c) { if (c) sanitize_config(*c); }
extern "C" int jt_build_filter_spec(const jt_filter_config *c, char *buf, size_t cap) { return c ? copy_out(build_spec(*c), buf, cap) : copy_out("", buf, cap); }
extern "C" int jt_build_filter(const jt_filter_config *c, int id, char *buf, size_t cap) { if (!c) return JT_ERR_INVALID_ARG; return copy_out(build_one(*c, id), buf, cap); }
extern "C" int jt_build_adeclick_filter(const jt_filter_config *c, char *buf, size_t cap)
{
    if (!c) return JT_ERR_INVALID_ARG;
    if (!c->adeclick.enabled) return copy_out("", buf, cap);
    std::string s = fmt("adeclick=t=%.1f:w=%.0f:o=%.0f", c->adeclick.threshold, c->adeclick.window, c->adeclick.overlap);
    if (c->adeclick.method[0]) s += std::string(":m=") + c->adeclick.method;
    return copy_out(s, buf, cap);
}
extern "C" int jt_go_format_g(double v, char *buf, size_t cap) { return copy_out(go_g(v), buf, cap); }

// ---- stages ----
extern "C" int64_t jt_vad_intervals_for_duration(int64_t d, int64_t hop) { return intervals_for(d, hop); }
extern "C" int jt_vad_histogram(const jt_interval *iv, int64_t n, int axis, double bw, int32_t *bins, int64_t cap, int64_t *nb, double *mn, double *mx, int64_t *count)
{
    const View v(iv, n); const Hist h = build_hist(v, axis, bw);
    if (nb) *nb = (int64_t)h.bins.size(); if (mn) *mn = h.lo; if (mx) *mx = h.hi; if (count) *count = h.count;
    if (bins) { if ((int64_t)h.bins.size() > cap) return JT_ERR_BUFFER; for (size_t i = 0; i < h.bins.size(); i++) bins[i] = h.bins[i]; }
    return JT_OK;
}
static Hist hist_from(const int32_t *bins, int64_t nb, double bw, double mn, double mx)
{
    Hist h; h.bw = bw; h.lo = mn; h.hi = mx;
    for (int64_t i = 0; i < nb; i++) { h.bins.push_back(bins[i]); h.count += bins[i]; }
    return h;
}
extern "C" double jt_vad_otsu_split(const int32_t *bins, int64_t nb, double bw, double mn, double mx) { return otsu(hist_from(bins, nb, bw, mn, mx)); }
extern "C" double jt_vad_hysteresis_margin(const int32_t *bins, int64_t nb, double bw, double mn, double split) { return hysteresis_margin(hist_from(bins, nb, bw, mn, mn), split); }
extern "C" double jt_vad_percentile_of_sorted(const double *s, int64_t n, double pct) { return pct_sorted(s, n, pct); }
extern "C" double jt_vad_percentile_floor(const double *s, int64_t n, double seed) { return percentile_floor(s, n, seed); }
extern "C" double jt_vad_clamp_split(double split, double floor, double p75) { return clamp_split(split, floor, p75); }
extern "C" double jt_vad_floored_fraction(const jt_interval *iv, int64_t n, int axis) { return floored_fraction(View(iv, n), axis); }
extern "C" int jt_vad_is_speech_interval(const jt_interval *iv, double split, int axis) { return iv ? is_speech(*iv, split, axis) : 0; }
extern "C" int jt_vad_gap_tolerance(const uint8_t *flags, int64_t n, int64_t hop) { return gap_tolerance(flags, n, hop); }
extern "C" int64_t jt_vad_build_speech_runs(const jt_interval *iv, int64_t n, double split, double margin, int tol, int axis, int64_t hop, jt_region *runs, int64_t cap)
{
    const std::vector<jt_region> r = speech_runs(View(iv, n), split, margin, tol, axis, hop);
    if (runs) { if ((int64_t)r.size() > cap) return JT_ERR_BUFFER; for (size_t i = 0; i < r.size(); i++) runs[i] = r[i]; }
    return (int64_t)r.size();
}
extern "C" int jt_vad_pick_low_cluster_region(const jt_interval *iv, int64_t n, double split, int axis, int64_t hop, jt_region *out)
{
    jt_region r; if (!low_cluster_region(View(iv, n), split, axis, hop, r)) return 0;
    if (out) *out = r; return 1;
}
extern "C" int jt_vad_gate_statistics(const jt_interval *iv, int64_t n, double split, int axis, const jt_region *sp, double *vl, double *nh, double *sep)
{
    double a, b, c; gate_statistics(View(iv, n), split, axis, sp, a, b, c);
    if (vl) *vl = a; if (nh) *nh = b; if (sep) *sep = c; return JT_OK;
}
extern "C" int jt_vad_estimate_noise_floor(const jt_interval *iv, int64_t n, double *floor, double *thr)
{
    double f, t; const bool ok = estimate_noise_floor(View(iv, n), f, t);
    if (floor) *floor = f; if (thr) *thr = t; return ok;
}
extern "C" int jt_vad_noise_profile(const jt_interval *iv, int64_t n, const jt_region *reg, jt_noise_profile *out)
{
    if (!reg) return 0;
    jt_noise_profile p; if (!noise_profile(View(iv, n), *reg, p)) return 0;
    if (out) *out = p; return 1;
}
extern "C" int64_t jt_vad_intervals_in_range(const jt_interval *iv, int64_t n, int64_t start, int64_t end, int64_t *first)
{
    const Span s = in_range(View(iv, n), start, end);
    if (first) *first = s.lo; return s.len();
}
extern "C" double jt_vad_score_interval_window(const jt_interval *iv, int64_t n) { return score_interval_window(iv, n); }
extern "C" double jt_vad_score_speech_window(const jt_interval *iv, int64_t n) { return score_speech_window(iv, n); }
extern "C" double jt_vad_level_variance(const jt_interval *iv, int64_t n, int axis) { return level_variance(iv, n, axis); }
extern "C" double jt_vad_score_candidate_grounded(const jt_speech_candidate *c, double floor_db, double var) { return c ? grounded_score(*c, floor_db, var) : 0.0; }
extern "C" int jt_vad_measure_candidate(const jt_interval *iv, int64_t n, const jt_region *reg, jt_speech_candidate *out)
{
    if (!reg) return 0;
    jt_speech_candidate c; if (!measure_candidate(View(iv, n), *reg, c)) return 0;
    if (out) *out = c; return 1;
}
extern "C" int jt_vad_refine_speech_region(const jt_interval *iv, int64_t n, const jt_region *cand, jt_region *out)
{
    if (!cand) return 0;
    jt_region r = *cand; const bool ok = refine(View(iv, n), r, kGoldenSpeechWindow, kGoldenSpeechMin, score_speech_window, true);
    if (out) *out = r; return ok;
}
extern "C" int jt_vad_find_best_speech_region(const jt_interval *iv, int64_t n, const jt_region *regions, int64_t nr, double floor_db,
                                              jt_region *best, jt_speech_candidate *cands, int64_t cap, int64_t *n_cands)
{
    std::vector<jt_region> rg(regions, regions + (regions ? nr : 0)); std::vector<jt_speech_candidate> cs; jt_region b{};
    const bool ok = find_best(View(iv, n), rg, floor_db, b, cs);
    if (n_cands) *n_cands = (int64_t)cs.size();
    if (cands) { if ((int64_t)cs.size() > cap) return JT_ERR_BUFFER; for (size_t i = 0; i < cs.size(); i++) cands[i] = cs[i]; }
    if (ok && best) *best = b;
    return ok;
}
extern "C" double jt_adapt_gate_threshold(double vl, double sep, int *narrow) { return gate_threshold(vl, sep, narrow); }
extern "C" double jt_adapt_gate_threshold_no_profile(double floor, double peak, double crest, double ratio, double gap) { return gate_threshold_no_profile(floor, peak, crest, ratio, gap); }
extern "C" int jt_adapt_band_noise(const double *bands, int n, char *buf, size_t cap) { return copy_out(band_noise(bands, n), buf, cap); }

