"""jt_run_record_json: the loudness / dynamics / spectral / noise / run blocks of the reference's per-file run
record, rendered as MarshalRunRecord writes them.  Replays the shape checks of internal/processor/runrecord_test.go:76-186
(TestRunRecord_FullShape, _AnalysisOnlyDropsProcessingBlocks, _NonFiniteFloatSerialisesAsNull) and the tag list of
runrecord_tags_test.go:100-116 through the C ABI, plus encoding/json's float text.  Host-only."""
import ctypes as C
import json
import math

from jivetalking_b200 import adapt as A
from jivetalking_b200 import gpudsp


class RunInfo(C.Structure):
    _fields_ = [("input_file", C.c_char_p), ("version", C.c_char_p), ("executable", C.c_char_p), ("processed_at", C.c_char_p),
                ("duration_s", C.c_double), ("sample_rate_hz", C.c_int32), ("channels", C.c_int32)]


def record(res, an, run, target=-16.0):
    L = gpudsp.lib()
    L.jt_run_record_json.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_char_p, C.c_size_t, C.POINTER(C.c_size_t)]
    need = C.c_size_t(0)
    assert L.jt_run_record_json(C.byref(res) if res is not None else None, C.byref(an) if an is not None else None,
                                C.byref(run) if run is not None else None, target, None, 0, C.byref(need)) == -7
    buf = C.create_string_buffer(need.value)
    assert L.jt_run_record_json(C.byref(res) if res is not None else None, C.byref(an) if an is not None else None,
                                C.byref(run) if run is not None else None, target, buf, len(buf), None) == 0
    return buf.value.decode()


def populated():
    res = gpudsp.ProcessResult()
    for i, m in enumerate((res.input, res.filtered, res.final)):
        m.input_i, m.input_tp, m.input_sp, m.input_lra, m.last_m, m.last_s = -23.5 + 3 * i, -3.2 - i, -3.9 - i, 7.25, -24.125, -23.75
        for k in range(gpudsp.AS_COUNT):
            m.astats[k] = 0.001 * (k + 1)
        m.astats[gpudsp.AS_NAMES.index("Crest_factor")] = 4.0
        m.astats[gpudsp.AS_NAMES.index("Min_level")] = -0.5
        m.astats[gpudsp.AS_NAMES.index("Max_level")] = 12000.0          # an integer-format link: normalised by 32768
        for k in range(gpudsp.SP_COUNT):
            m.spectral_mean[k] = 100.0 * (k + 1) + 0.5
    an = A.Analysis()
    an.measurements = res.input
    an.voice_activity.floor, an.voice_activity.floor_source, an.voice_activity.voice_activated = -58.25, 3, 1
    run = RunInfo(b"episode-LUFS-16-processed.flac", b"dev", b"/usr/bin/jivetalking", b"2026-01-02T03:04:05Z", 3600.5, 48000, 1)
    return res, an, run


KEYS = ["integrated_lufs", "true_peak_dbtp", "lra_lu", "thresh_lufs", "target_offset_db", "momentary_lufs", "short_term_lufs",
        "sample_peak_dbfs", "rms_level_dbfs", "peak_level_dbfs", "dynamic_range_db", "crest_factor_astats_db", "rms_trough_dbfs",
        "rms_peak_dbfs", "dc_offset", "flat_factor", "zero_crossings_rate", "zero_crossings_count", "min_level_dbfs",
        "max_level_dbfs", "bit_depth", "number_of_samples", "noise_floor_count", "entropy", "floor_dbfs", "floor_source",
        "floor_prescan_dbfs", "floor_astats_dbfs", "reduction_headroom_db", "room_tone_detect_level_dbfs", "voice_activated",
        "centroid_hz", "spread_hz", "rolloff_hz"]                        # runrecord_tags_test.go:100-116


def all_keys(t, out):
    if isinstance(t, dict):
        for k, v in t.items():
            out.add(k)
            all_keys(v, out)
    return out


def test_full_shape():
    res, an, run = populated()
    text = record(res, an, run)
    tree = json.loads(text)
    assert tree["schema_version"] == 1 and isinstance(tree["schema_version"], int)
    for dom in ("loudness", "dynamics", "spectral"):
        assert set(tree[dom]["stages"]) == {"input", "filtered", "final"}
    assert tree["run"]["sample_rate_hz"] == 48000 and tree["run"]["channels"] == 1 and tree["run"]["input_file"] == "episode-LUFS-16-processed.flac"
    assert not [k for k in KEYS if k not in all_keys(tree, set())]
    lo = tree["loudness"]
    assert lo["target_i_lufs"] == -16 and lo["stages"]["input"]["integrated_lufs"] == -23.5 and lo["stages"]["final"]["integrated_lufs"] == -17.5
    assert lo["stages"]["input"]["thresh_lufs"] == -33.5 and lo["stages"]["input"]["target_offset_db"] == 7.5      # I - 10, target - I
    assert lo["stages"]["filtered"]["thresh_lufs"] == -30.5 and lo["stages"]["filtered"]["target_offset_db"] == 0
    dy = tree["dynamics"]["stages"]["input"]
    assert abs(dy["crest_factor_astats_db"] - 20 * math.log10(4.0)) < 1e-12          # linear ratio -> dB
    assert abs(dy["min_level_dbfs"] - 20 * math.log10(0.5)) < 1e-12 and abs(dy["max_level_dbfs"] - 20 * math.log10(12000 / 32768)) < 1e-12
    assert tree["noise"]["floor_source"] == "vad_percentile" and tree["noise"]["voice_activated"] is True
    # MarshalRunRecord marshals a generic tree: sorted keys at every level, two-space indent
    def sorted_everywhere(t):
        return not isinstance(t, dict) or (list(t) == sorted(t) and all(sorted_everywhere(v) for v in t.values()))
    assert sorted_everywhere(json.loads(text, object_pairs_hook=dict))
    assert text.startswith('{\n  "dynamics": {\n    "stages": {\n      "filtered": {\n        "bit_depth": ')
    assert text == json.dumps(tree, indent=2, sort_keys=True, separators=(",", ": ")).replace("3600.5", "3600.5")     # same layout as Go's MarshalIndent


def test_analysis_only_drops_processing_stages():
    res, an, run = populated()
    tree = json.loads(record(None, an, None))
    assert set(tree["loudness"]["stages"]) == {"input"} and "run" not in tree and "filters" not in tree and "normalisation" not in tree
    assert tree["loudness"]["stages"]["input"]["integrated_lufs"] == -23.5


def test_non_finite_floats_become_null():
    res, an, run = populated()
    res.input.astats[gpudsp.AS_NAMES.index("RMS_level")] = float("-inf")
    res.input.input_tp = float("inf")
    res.final.spectral_mean[2] = float("nan")
    tree = json.loads(record(res, None, None))
    assert tree["dynamics"]["stages"]["input"]["rms_level_dbfs"] is None and tree["loudness"]["stages"]["input"]["true_peak_dbtp"] is None
    assert tree["spectral"]["stages"]["final"]["centroid_hz"] is None and "noise" not in tree


def test_float_text_is_encoding_json():
    L = gpudsp.lib()
    L.jt_go_json_float.argtypes = [C.c_double, C.c_char_p, C.c_size_t]
    def f(v):
        b = C.create_string_buffer(64)
        assert L.jt_go_json_float(v, b, 64) == 0
        return b.value.decode()
    cases = {0.1: "0.1", -16.0: "-16", 1e-7: "1e-7", 1.5e-10: "1.5e-10", 1e-6: "0.000001", 9.999e-7: "9.999e-7", 1e20: "100000000000000000000",
             1e21: "1e+21", 123456789.0: "123456789", -23.456: "-23.456", 5e-324: "5e-324", 0.30000000000000004: "0.30000000000000004",
             float("nan"): "null", float("inf"): "null", 0.0: "0"}
    for v, want in cases.items():
        assert f(v) == want, (v, f(v), want)
