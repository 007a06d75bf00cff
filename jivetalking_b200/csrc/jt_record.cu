// jt_record.cu -- the numbers of the reference's per-file run record: the `loudness`, `dynamics`, `spectral`
// and `noise` domains of RunRecord (internal/processor/runrecord.go:24-100) rendered as the JSON text MarshalRunRecord writes
// (runrecord.go:425-433): the record is reflected into a generic tree, so every object's keys come out SORTED, floats are
// encoding/json's shortest round-trip form ('e' notation below 1e-6 / from 1e21, exponent without padding), non-finite
// values become null, and the indent is two spaces.  Tags: InputLoudnessMetrics / OutputLoudnessMetrics / DynamicsMetrics /
// NoiseMetrics (analyser.go:140-199, 262-264), SpectralMetrics (analyser_metrics.go:696-710).  Host-only, no jt_ctx.
#include "../../include/jtdsp.h"
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

namespace {
// strconv.AppendFloat(b, f, fmt, -1, 64) as encoding/json's floatEncoder uses it
std::string go_json_float(double v)
{
    if (!std::isfinite(v)) return "null";
    if (v == 0) return std::signbit(v) ? "-0" : "0";
    char e[48]; int p = 1;
    for (; p <= 17; p++) { snprintf(e, sizeof(e), "%.*e", p - 1, v); if (strtod(e, nullptr) == v) break; }
    std::string digits; const char *q = e; bool neg = false;
    if (*q == '-') { neg = true; q++; }
    for (; *q && *q != 'e'; q++) if (*q != '.') digits.push_back(*q);
    const int x = atoi(q + 1);
    while (digits.size() > 1 && digits.back() == '0') digits.pop_back();
    const int nd = (int)digits.size();
    const double a = fabs(v);
    std::string out = neg ? "-" : "";
    if (a < 1e-6 || a >= 1e21) {
        out += digits[0];
        if (nd > 1) { out += '.'; out += digits.substr(1); }
        char xb[16]; snprintf(xb, sizeof(xb), "e%c%d", x < 0 ? '-' : '+', abs(x));        // "e-09" is cleaned to "e-9"
        return out + xb;
    }
    if (x < 0) { out += "0."; out.append((size_t)(-x - 1), '0'); out += digits; return out; }
    if (nd <= x + 1) { out += digits; out.append((size_t)(x + 1 - nd), '0'); return out; }
    out += digits.substr(0, (size_t)x + 1); out += '.'; out += digits.substr((size_t)x + 1);
    return out;
}
std::string go_json_string(const char *s)
{
    std::string o = "\"";
    for (const unsigned char *p = (const unsigned char *)s; *p; p++) {
        if (*p == '"' || *p == '\\') { o += '\\'; o += (char)*p; }
        else if (*p == '\n') o += "\\n"; else if (*p == '\r') o += "\\r"; else if (*p == '\t') o += "\\t";
        else if (*p < 0x20 || *p == '<' || *p == '>' || *p == '&') { char b[8]; snprintf(b, sizeof(b), "\\u%04x", *p); o += b; }     // HTML-safe escaping is encoding/json's default
        else o += (char)*p;
    }
    return o + "\"";
}

struct Node {      // an ordered-by-key object or a leaf already rendered
    std::string leaf; std::map<std::string, Node> kids; bool is_obj = false;
    static Node num(double v) { Node n; n.leaf = go_json_float(v); return n; }
    static Node integer(long long v) { Node n; n.leaf = std::to_string(v); return n; }
    static Node boolean(bool v) { Node n; n.leaf = v ? "true" : "false"; return n; }
    static Node str(const char *s) { Node n; n.leaf = go_json_string(s); return n; }
    static Node obj() { Node n; n.is_obj = true; return n; }
    Node &operator[](const char *k) { is_obj = true; return kids[k]; }
};
void render(const Node &n, int depth, std::string &out)
{
    if (!n.is_obj) { out += n.leaf; return; }
    if (n.kids.empty()) { out += "{}"; return; }
    out += "{\n";
    size_t i = 0;
    for (const auto &kv : n.kids) {
        out.append((size_t)(depth + 1) * 2, ' ');
        out += go_json_string(kv.first.c_str()); out += ": ";
        render(kv.second, depth + 1, out);
        if (++i < n.kids.size()) out += ',';
        out += '\n';
    }
    out.append((size_t)depth * 2, ' ');
    out += '}';
}

double ratio_db(double r) { return r <= 0 ? -120.0 : 20 * log10(r); }                     // linearRatioToDB
double sample_dbfs(double s)                                                               // linearSampleToDBFS (analyser_metrics.go:672-692)
{
    double a = fabs(s);
    if (a <= 0) return -120.0;
    if (a > 1.0) a /= 32768.0;
    if (a > 1.0) a = 1.0;
    return 20 * log10(a);
}
// Go zero value for a key that was absent from the metadata (the accumulators start at 0)
double z(double v) { return std::isnan(v) ? 0.0 : v; }

Node loudness_stage(const jt_measurements &m, bool input, double target_i)
{
    Node n = Node::obj();
    n["momentary_lufs"] = Node::num(m.last_m); n["short_term_lufs"] = Node::num(m.last_s); n["sample_peak_dbfs"] = Node::num(m.input_sp);
    n["integrated_lufs"] = Node::num(m.input_i); n["true_peak_dbtp"] = Node::num(m.input_tp); n["lra_lu"] = Node::num(m.input_lra);
    if (input) {            // analyser.go:395 InputThresh = I - 10; TargetOffset = config.TargetI - InputI
        n["thresh_lufs"] = Node::num(m.input_i - 10.0); n["target_offset_db"] = Node::num(target_i - m.input_i);
    } else {                // finalizeOutputMeasurements (analyser_metrics.go:1034-1038): the missing threshold key falls back to I - 10
        n["thresh_lufs"] = Node::num(m.input_i != 0.0 ? m.input_i - 10.0 : 0.0); n["target_offset_db"] = Node::num(0.0);
    }
    return n;
}
Node dynamics_stage(const jt_measurements &m)
{
    const double *a = m.astats;
    Node n = Node::obj();
    n["dynamic_range_db"] = Node::num(z(a[JT_AS_Dynamic_range])); n["rms_level_dbfs"] = Node::num(z(a[JT_AS_RMS_level]));
    n["peak_level_dbfs"] = Node::num(z(a[JT_AS_Peak_level])); n["rms_trough_dbfs"] = Node::num(z(a[JT_AS_RMS_trough]));
    n["rms_peak_dbfs"] = Node::num(z(a[JT_AS_RMS_peak])); n["dc_offset"] = Node::num(z(a[JT_AS_DC_offset]));
    n["flat_factor"] = Node::num(z(a[JT_AS_Flat_factor]));
    n["crest_factor_astats_db"] = Node::num(std::isnan(a[JT_AS_Crest_factor]) ? 0.0 : ratio_db(a[JT_AS_Crest_factor]));
    n["zero_crossings_rate"] = Node::num(z(a[JT_AS_Zero_crossings_rate])); n["zero_crossings_count"] = Node::num(z(a[JT_AS_Zero_crossings]));
    n["max_difference"] = Node::num(z(a[JT_AS_Max_difference])); n["min_difference"] = Node::num(z(a[JT_AS_Min_difference]));
    n["mean_difference"] = Node::num(z(a[JT_AS_Mean_difference])); n["rms_difference"] = Node::num(z(a[JT_AS_RMS_difference]));
    n["entropy"] = Node::num(z(a[JT_AS_Entropy]));
    n["min_level_dbfs"] = Node::num(std::isnan(a[JT_AS_Min_level]) ? 0.0 : sample_dbfs(a[JT_AS_Min_level]));
    n["max_level_dbfs"] = Node::num(std::isnan(a[JT_AS_Max_level]) ? 0.0 : sample_dbfs(a[JT_AS_Max_level]));
    n["noise_floor_count"] = Node::num(z(a[JT_AS_Noise_floor_count])); n["bit_depth"] = Node::num(z(a[JT_AS_Bit_depth]));
    n["number_of_samples"] = Node::num(z(a[JT_AS_Number_of_samples]));
    return n;
}
Node spectral_stage(const jt_measurements &m)
{
    static const char *keys[JT_SP_COUNT] = {"mean", "variance", "centroid_hz", "spread_hz", "skewness", "kurtosis", "entropy", "flatness",
                                            "crest", "flux", "slope", "decrease", "rolloff_hz"};
    Node n = Node::obj();
    for (int k = 0; k < JT_SP_COUNT; k++) n[keys[k]] = Node::num(m.spectral_mean[k]);
    return n;
}
}   // namespace

extern "C" int jt_run_record_json(const jt_process_result *res, const jt_analysis *analysis, const jt_run_info *run, double target_i,
                                  char *buf, size_t cap, size_t *needed)
{
    if (!res && !analysis) return JT_ERR_INVALID_ARG;
    const bool processed = res != nullptr;                         // analysis-only records drop the filtered / final stages (omitempty)
    const jt_measurements &in = processed ? res->input : analysis->measurements;
    Node root = Node::obj();
    root["schema_version"] = Node::integer(1);
    if (run) {
        Node r = Node::obj();
        r["input_file"] = Node::str(run->input_file ? run->input_file : ""); r["version"] = Node::str(run->version ? run->version : "");
        r["executable"] = Node::str(run->executable ? run->executable : ""); r["processed_at"] = Node::str(run->processed_at ? run->processed_at : "");
        r["duration_s"] = Node::num(run->duration_s); r["sample_rate_hz"] = Node::integer(run->sample_rate_hz); r["channels"] = Node::integer(run->channels);
        root["run"] = r;
    }
    Node lst = Node::obj(), dst = Node::obj(), sst = Node::obj();
    lst["input"] = loudness_stage(in, true, target_i); dst["input"] = dynamics_stage(in); sst["input"] = spectral_stage(in);
    if (processed) {
        lst["filtered"] = loudness_stage(res->filtered, false, target_i); dst["filtered"] = dynamics_stage(res->filtered); sst["filtered"] = spectral_stage(res->filtered);
        lst["final"] = loudness_stage(res->final, false, target_i); dst["final"] = dynamics_stage(res->final); sst["final"] = spectral_stage(res->final);
    }
    Node loud = Node::obj(); loud["target_i_lufs"] = Node::num(target_i); loud["stages"] = lst; root["loudness"] = loud;
    Node dyn = Node::obj(); dyn["stages"] = dst; root["dynamics"] = dyn;
    Node spc = Node::obj(); spc["stages"] = sst; root["spectral"] = spc;
    if (analysis) {
        const jt_voice_activity &va = analysis->voice_activity;
        static const char *src[] = {"astats", "rms_estimate", "ebur128_estimate", "vad_percentile"};
        Node nz = Node::obj();
        nz["floor_dbfs"] = Node::num(va.floor); nz["floor_source"] = Node::str(va.floor_source >= 0 && va.floor_source < 4 ? src[va.floor_source] : "");
        nz["floor_prescan_dbfs"] = Node::num(va.floor_prescan); nz["floor_astats_dbfs"] = Node::num(va.floor_astats);
        nz["room_tone_detect_level_dbfs"] = Node::num(va.room_tone_detect_level); nz["voice_activated"] = Node::boolean(va.voice_activated != 0);
        nz["floored_fraction"] = Node::num(va.floored_fraction); nz["reduction_headroom_db"] = Node::num(va.reduction_headroom);
        root["noise"] = nz;
    }
    std::string out;
    render(root, 0, out);
    if (needed) *needed = out.size() + 1;
    if (!buf || out.size() + 1 > cap) return JT_ERR_BUFFER;
    memcpy(buf, out.c_str(), out.size() + 1);
    return JT_OK;
}

extern "C" int jt_go_json_float(double v, char *buf, size_t cap)
{
    const std::string s = go_json_float(v);
    if (!buf || s.size() + 1 > cap) return JT_ERR_BUFFER;
    memcpy(buf, s.c_str(), s.size() + 1);
    return JT_OK;
}
