"""ctypes binding of the Pass 1 -> Pass 2 host logic of libjtdsp.so (include/jtdsp.h, csrc/jt_adapt.cu): the voice-
activity detector, region election, AdaptConfig and BuildFilterSpec of the reference
(internal/processor/analyser_vad.go, analyser_candidates_*.go, analyser_noise_seed.go, adaptive*.go, filters.go).

Function names follow the reference's (snake-cased) so the parity tests read like its own table-driven tests.  All of
it is host code inside the shared library; this module only marshals arguments (it never imports oracle/).
"""
import ctypes as C
import math

from . import gpudsp
from .gpudsp import Interval, Measurements, ProcessResult, SP_COUNT, SP_NAMES, AS_NAMES, AS_COUNT

NS_MS, NS_S = 1_000_000, 1_000_000_000
HOP_NS = 250 * NS_MS
AXIS_MOMENTARY, AXIS_RMS = 0, 1
AFFTDN_BANDS = 15
(FILTER_DOWNMIX, FILTER_ANALYSIS, FILTER_RESAMPLE, FILTER_RUMBLE_HIGHPASS, FILTER_BANDLIMIT_LOWPASS, FILTER_SPEECH_GATE,
 FILTER_NOISE_REDUCTION, FILTER_LEVELLING_COMPRESSOR, FILTER_DEESSER) = range(1, 10)
PASS1_ORDER = [FILTER_DOWNMIX, FILTER_ANALYSIS]
PASS2_ORDER = [FILTER_DOWNMIX, FILTER_RUMBLE_HIGHPASS, FILTER_BANDLIMIT_LOWPASS, FILTER_NOISE_REDUCTION, FILTER_SPEECH_GATE,
               FILTER_LEVELLING_COMPRESSOR, FILTER_DEESSER, FILTER_ANALYSIS, FILTER_RESAMPLE]

_D, _I64, _I32, _INT, _P = C.c_double, C.c_int64, C.c_int32, C.c_int, C.c_void_p


class Region(C.Structure):
    _fields_ = [("start_ns", _I64), ("end_ns", _I64), ("duration_ns", _I64)]

    @classmethod
    def of(cls, start_ns, end_ns):
        return cls(int(start_ns), int(end_ns), int(end_ns) - int(start_ns))

    def __repr__(self):
        return f"Region({self.start_ns / 1e9:.3f}s..{self.end_ns / 1e9:.3f}s)"


class RegionSample(C.Structure):
    _fields_ = [("rms_level", _D), ("peak_level", _D), ("crest_factor", _D), ("spectral", _D * SP_COUNT),
                ("momentary_lufs", _D), ("short_term_lufs", _D), ("true_peak", _D), ("sample_peak", _D)]


class SpeechCandidate(C.Structure):
    _fields_ = [("region", Region), ("sample", RegionSample), ("voicing_density", _D), ("body_band_rms", _D),
                ("sib_band_rms", _D), ("score", _D), ("original_start_ns", _I64), ("original_duration_ns", _I64),
                ("bands_measured", _I32), ("was_refined", _I32)]


class NoiseProfile(C.Structure):
    _fields_ = [("start_ns", _I64), ("duration_ns", _I64), ("measured_noise_floor", _D), ("peak_level", _D),
                ("crest_factor", _D), ("entropy", _D), ("spectral", _D * SP_COUNT), ("band_noise", _D * AFFTDN_BANDS),
                ("bands_measured", _I32), ("n_band_noise", _I32), ("warning", _I32), ("reserved", _I32)]


class VoiceActivity(C.Structure):
    _fields_ = [("floor", _D), ("floor_prescan", _D), ("floor_astats", _D), ("room_tone_detect_level", _D),
                ("floored_fraction", _D), ("reduction_headroom", _D), ("split", _D), ("margin", _D),
                ("voiced_low_percentile", _D), ("noise_high_percentile", _D), ("gate_separation_db", _D),
                ("floor_source", _I32), ("voice_activated", _I32), ("gap_tolerance", _I32), ("reserved", _I32),
                ("has_noise_profile", _I32), ("has_room_tone_sample", _I32), ("has_speech_profile", _I32),
                ("speech_profile_index", _I32), ("n_speech_regions", _I64), ("n_candidates", _I64),
                ("noise_region", Region), ("noise_profile", NoiseProfile), ("room_tone_sample", RegionSample),
                ("speech_profile", SpeechCandidate)]


class BiquadConfig(C.Structure):
    _fields_ = [("enabled", _I32), ("poles", _I32), ("frequency", _D), ("width", _D), ("mix", _D), ("transform", C.c_char * 8)]


class _NoiseReduction(C.Structure):
    _fields_ = [("enabled", _I32), ("afftdn_enabled", _I32), ("afftdn_track_noise", _I32), ("reserved", _I32),
                ("strength", _D), ("patch_s", _D), ("research_s", _D), ("smooth", _D), ("afftdn_noise_reduction", _D),
                ("afftdn_noise_floor", _D), ("afftdn_noise_type", C.c_char * 8), ("afftdn_band_noise", C.c_char * 136)]


class _SpeechGate(C.Structure):
    _fields_ = [("enabled", _I32), ("reserved", _I32), ("threshold", _D), ("ratio", _D), ("attack", _D), ("release", _D),
                ("range", _D), ("knee", _D), ("makeup", _D), ("detection", C.c_char * 8)]


class _Compressor(C.Structure):
    _fields_ = [("enabled", _I32), ("reserved", _I32), ("threshold", _D), ("ratio", _D), ("attack", _D), ("release", _D),
                ("makeup", _D), ("knee", _D), ("mix", _D)]


class _Deesser(C.Structure):
    _fields_ = [("enabled", _I32), ("reserved", _I32), ("intensity", _D), ("amount", _D), ("frequency", _D)]


class _Adeclick(C.Structure):
    _fields_ = [("enabled", _I32), ("reserved", _I32), ("threshold", _D), ("window", _D), ("overlap", _D), ("method", C.c_char * 8)]


class _Loudnorm(C.Structure):
    _fields_ = [("enabled", _I32), ("dual_mono", _I32), ("linear", _I32), ("reserved", _I32), ("target_i", _D),
                ("target_tp", _D), ("target_lra", _D)]


class FilterConfig(C.Structure):
    _fields_ = [("downmix_enabled", _I32), ("analysis_enabled", _I32), ("resample_enabled", _I32), ("resample_rate", _I32),
                ("resample_frame_size", _I32), ("reserved0", _I32), ("resample_format", C.c_char * 8),
                ("rumble_highpass", BiquadConfig), ("bandlimit_lowpass", BiquadConfig),
                ("noise_reduction", _NoiseReduction), ("speech_gate", _SpeechGate), ("levelling_compressor", _Compressor),
                ("deesser", _Deesser), ("adeclick", _Adeclick), ("loudnorm", _Loudnorm),
                ("n_filter_order", _I32), ("filter_order", _I32 * 12), ("reserved1", _I32)]

    def set_order(self, order):
        self.n_filter_order = len(order)
        for i, f in enumerate(order):
            self.filter_order[i] = f

    def copy(self):
        c = FilterConfig()
        C.memmove(C.byref(c), C.byref(self), C.sizeof(FilterConfig))
        return c


class AdaptDiagnostics(C.Structure):
    _fields_ = [("bandlimit_lp_reason", C.c_char * 40), ("speech_gate_dynamic_range", _D),
                ("speech_gate_quiet_speech_estimate", _D), ("speech_gate_speech_separation", _D),
                ("speech_gate_speech_headroom", _D), ("speech_gate_threshold_unclamped", _D), ("speech_gate_depth_db", _D),
                ("speech_gate_clamp_reason", C.c_char * 16), ("speech_gate_narrow_gap", _I32), ("afftdn_enabled", _I32),
                ("afftdn_noise_floor_db", _D), ("afftdn_disable_reason", C.c_char * 24), ("afftdn_noise_type", C.c_char * 8)]


class OutputRegions(C.Structure):
    _fields_ = [("room_tone", RegionSample), ("speech", RegionSample), ("has_room_tone", _I32), ("has_speech", _I32)]


class Analysis(C.Structure):
    _fields_ = [("measurements", Measurements), ("voice_activity", VoiceActivity), ("config", FilterConfig),
                ("diagnostics", AdaptDiagnostics), ("pass2_spec", C.c_char * 2048),
                ("filtered_regions", OutputRegions), ("final_regions", OutputRegions)]


_bound = False


def _L():
    global _bound
    L = gpudsp.lib()
    if _bound:
        return L
    PI, PR, PC = C.POINTER(Interval), C.POINTER(Region), C.POINTER(SpeechCandidate)
    L.jt_detect_voice_activity.argtypes = [C.POINTER(Measurements), PI, _I64, C.POINTER(VoiceActivity), PR, _I64, PC, _I64]
    L.jt_vad_detect.argtypes = [PI, _I64, _D, C.POINTER(VoiceActivity), PR, _I64, PC, _I64]
    L.jt_apply_band_rms.argtypes = [C.POINTER(VoiceActivity), _P, _P, _P, _P]
    L.jt_band_plan.argtypes = [_P, _P]
    L.jt_band_plan.restype = None
    L.jt_default_filter_config.argtypes = [C.POINTER(FilterConfig)]
    L.jt_default_filter_config.restype = None
    L.jt_adapt_config.argtypes = [C.POINTER(FilterConfig), C.POINTER(Measurements), C.POINTER(VoiceActivity),
                                  C.POINTER(FilterConfig), C.POINTER(AdaptDiagnostics)]
    L.jt_sanitize_config.argtypes = [C.POINTER(FilterConfig)]
    L.jt_sanitize_config.restype = None
    L.jt_build_filter_spec.argtypes = [C.POINTER(FilterConfig), C.c_char_p, C.c_size_t]
    L.jt_build_filter.argtypes = [C.POINTER(FilterConfig), _INT, C.c_char_p, C.c_size_t]
    L.jt_build_adeclick_filter.argtypes = [C.POINTER(FilterConfig), C.c_char_p, C.c_size_t]
    L.jt_go_format_g.argtypes = [_D, C.c_char_p, C.c_size_t]
    L.jt_vad_intervals_for_duration.restype = _I64
    L.jt_vad_intervals_for_duration.argtypes = [_I64, _I64]
    L.jt_vad_histogram.argtypes = [PI, _I64, _INT, _D, _P, _I64, C.POINTER(_I64), C.POINTER(_D), C.POINTER(_D), C.POINTER(_I64)]
    L.jt_vad_otsu_split.restype = _D
    L.jt_vad_otsu_split.argtypes = [_P, _I64, _D, _D, _D]
    L.jt_vad_hysteresis_margin.restype = _D
    L.jt_vad_hysteresis_margin.argtypes = [_P, _I64, _D, _D, _D]
    L.jt_vad_percentile_of_sorted.restype = _D
    L.jt_vad_percentile_of_sorted.argtypes = [_P, _I64, _D]
    L.jt_vad_percentile_floor.restype = _D
    L.jt_vad_percentile_floor.argtypes = [_P, _I64, _D]
    L.jt_vad_clamp_split.restype = _D
    L.jt_vad_clamp_split.argtypes = [_D, _D, _D]
    L.jt_vad_floored_fraction.restype = _D
    L.jt_vad_floored_fraction.argtypes = [PI, _I64, _INT]
    L.jt_vad_is_speech_interval.argtypes = [PI, _D, _INT]
    L.jt_vad_gap_tolerance.argtypes = [_P, _I64, _I64]
    L.jt_vad_build_speech_runs.restype = _I64
    L.jt_vad_build_speech_runs.argtypes = [PI, _I64, _D, _D, _INT, _INT, _I64, PR, _I64]
    L.jt_vad_pick_low_cluster_region.argtypes = [PI, _I64, _D, _INT, _I64, PR]
    L.jt_vad_gate_statistics.argtypes = [PI, _I64, _D, _INT, PR, C.POINTER(_D), C.POINTER(_D), C.POINTER(_D)]
    L.jt_vad_estimate_noise_floor.argtypes = [PI, _I64, C.POINTER(_D), C.POINTER(_D)]
    L.jt_vad_noise_profile.argtypes = [PI, _I64, PR, C.POINTER(NoiseProfile)]
    L.jt_vad_intervals_in_range.restype = _I64
    L.jt_vad_intervals_in_range.argtypes = [PI, _I64, _I64, _I64, C.POINTER(_I64)]
    for f in ("jt_vad_score_interval_window", "jt_vad_score_speech_window"):
        getattr(L, f).restype = _D
        getattr(L, f).argtypes = [PI, _I64]
    L.jt_vad_level_variance.restype = _D
    L.jt_vad_level_variance.argtypes = [PI, _I64, _INT]
    L.jt_vad_score_candidate_grounded.restype = _D
    L.jt_vad_score_candidate_grounded.argtypes = [PC, _D, _D]
    L.jt_vad_measure_candidate.argtypes = [PI, _I64, PR, PC]
    L.jt_vad_refine_speech_region.argtypes = [PI, _I64, PR, PR]
    L.jt_vad_find_best_speech_region.argtypes = [PI, _I64, PR, _I64, _D, PR, PC, _I64, C.POINTER(_I64)]
    L.jt_adapt_gate_threshold.restype = _D
    L.jt_adapt_gate_threshold.argtypes = [_D, _D, C.POINTER(_INT)]
    L.jt_adapt_gate_threshold_no_profile.restype = _D
    L.jt_adapt_gate_threshold_no_profile.argtypes = [_D] * 5
    L.jt_adapt_band_noise.argtypes = [_P, _INT, C.c_char_p, C.c_size_t]
    mo = [_P, _P, _I64, _INT, _INT, _INT, _I64, _I64, C.POINTER(RegionSample), C.POINTER(_I64)]
    L.jt_measure_output_region.argtypes = mo
    L.jt_measure_output_region_dev.argtypes = mo
    L.jt_analyse_adaptive.argtypes = [_P, _P, _I64, _INT, _INT, _INT, _INT, C.POINTER(FilterConfig), C.POINTER(Analysis),
                                      PI, _I64, C.POINTER(_I64)]
    pa = [_P, _P, _I64, _INT, _INT, _INT, C.POINTER(FilterConfig), _P, _I64, C.POINTER(ProcessResult), C.POINTER(Analysis)]
    L.jt_process_audio_adaptive.argtypes = pa
    L.jt_process_audio_adaptive_dev.argtypes = pa
    L.jt_sharded_plan.argtypes = [_I64, _INT, _INT, _INT, C.POINTER(ShardPlan)]
    sh = [_P, _P, _I64, _INT, _INT, _INT, _I64, _INT, _INT, C.POINTER(FilterConfig), _INT, _P, _I64, C.POINTER(_I64), C.POINTER(_I64),
          C.POINTER(ProcessResult), C.POINTER(Analysis), C.POINTER(ShardTiming)]
    L.jt_process_audio_sharded.argtypes = sh
    L.jt_process_audio_sharded_dev.argtypes = sh
    _bound = True
    return L


# ---- interval construction ---------------------------------------------------------------------------------------
def interval(timestamp_ns=0, rms=0.0, peak=0.0, momentary=0.0, short_term=0.0, true_peak=0.0, sample_peak=0.0, found=True, **spectral):
    """IntervalSample literal (analyser_metrics.go:17-32); spectral fields by aspectralstats name."""
    iv = Interval()
    iv.timestamp_s = timestamp_ns * 1e-9
    iv.rms_level, iv.peak_level = rms, peak
    iv.momentary_lufs, iv.short_term_lufs, iv.true_peak, iv.sample_peak = momentary, short_term, true_peak, sample_peak
    iv.spectral_found = 1 if found else 0
    for k, v in spectral.items():
        iv.spectral[SP_NAMES.index(k)] = v
    return iv


def _arr(intervals):
    if isinstance(intervals, C.Array):
        return intervals, len(intervals)
    a = (Interval * max(len(intervals), 1))(*intervals)
    return a, len(intervals)


def _darr(xs):
    return (_D * max(len(xs), 1))(*xs)


def _check(rc):
    if rc < 0:
        raise gpudsp.JtError(rc, "adaptive host call")
    return rc


# ---- the detector's stages ----------------------------------------------------------------------------------------
def intervals_for_duration(d_ns, hop_ns):
    return _L().jt_vad_intervals_for_duration(int(d_ns), int(hop_ns))


class Histogram:
    def __init__(self, bins, bin_width, min_level, max_level, count):
        self.bins, self.bin_width, self.min_level, self.max_level, self.count = bins, bin_width, min_level, max_level, count

    def bin_centre(self, i):
        return self.min_level + (i + 0.5) * self.bin_width


def build_level_histogram(intervals, axis, bin_width):
    a, n = _arr(intervals)
    nb, mn, mx, cnt = _I64(), _D(), _D(), _I64()
    _check(_L().jt_vad_histogram(a, n, axis, bin_width, None, 0, C.byref(nb), C.byref(mn), C.byref(mx), C.byref(cnt)))
    bins = (_I32 * max(nb.value, 1))()
    _check(_L().jt_vad_histogram(a, n, axis, bin_width, bins, nb.value, C.byref(nb), C.byref(mn), C.byref(mx), C.byref(cnt)))
    return Histogram(list(bins[:nb.value]), bin_width if nb.value else 0.0, mn.value, mx.value, cnt.value)


def otsu_split(h):
    b = (_I32 * max(len(h.bins), 1))(*h.bins)
    return _L().jt_vad_otsu_split(b, len(h.bins), h.bin_width, h.min_level, h.max_level)


def hysteresis_margin(h, split):
    b = (_I32 * max(len(h.bins), 1))(*h.bins)
    return _L().jt_vad_hysteresis_margin(b, len(h.bins), h.bin_width, h.min_level, split)


def vad_levels(intervals, axis):
    """Sorted non-floored levels (analyser_vad.go:152-163) -- plain marshalling: the C side has the same helper inside."""
    out = []
    for iv in intervals:
        l = iv.rms_level if axis == AXIS_RMS else iv.momentary_lufs
        if math.isinf(l) or math.isnan(l) or l <= -115.0:
            continue
        out.append(l)
    return sorted(out)


def percentile_of_sorted(sorted_levels, pct):
    return _L().jt_vad_percentile_of_sorted(_darr(sorted_levels), len(sorted_levels), pct)


def percentile_floor(sorted_levels, seed):
    return _L().jt_vad_percentile_floor(_darr(sorted_levels), len(sorted_levels), seed)


def clamp_split(split, noise_floor, p75):
    return _L().jt_vad_clamp_split(split, noise_floor, p75)


def floored_fraction(intervals, axis=AXIS_MOMENTARY):
    a, n = _arr(intervals)
    return _L().jt_vad_floored_fraction(a, n, axis)


def is_speech_interval(iv, split, axis=AXIS_MOMENTARY):
    return bool(_L().jt_vad_is_speech_interval(C.byref(iv), split, axis))


def gap_tolerance_intervals(flags, hop_ns=HOP_NS):
    b = (C.c_uint8 * max(len(flags), 1))(*[1 if f else 0 for f in flags])
    return _L().jt_vad_gap_tolerance(b, len(flags), hop_ns)


def build_speech_runs(intervals, split, margin, tol, axis=AXIS_MOMENTARY, hop_ns=HOP_NS):
    a, n = _arr(intervals)
    cnt = _check(_L().jt_vad_build_speech_runs(a, n, split, margin, tol, axis, hop_ns, None, 0))
    runs = (Region * max(cnt, 1))()
    _check(_L().jt_vad_build_speech_runs(a, n, split, margin, tol, axis, hop_ns, runs, cnt))
    return list(runs[:cnt])


def pick_low_cluster_region(intervals, split, axis=AXIS_MOMENTARY, hop_ns=HOP_NS):
    a, n = _arr(intervals)
    r = Region()
    return r if _L().jt_vad_pick_low_cluster_region(a, n, split, axis, hop_ns, C.byref(r)) else None


def derive_gate_statistics(intervals, split, axis=AXIS_MOMENTARY, speech_region=None):
    a, n = _arr(intervals)
    v, nh, s = _D(), _D(), _D()
    _L().jt_vad_gate_statistics(a, n, split, axis, C.byref(speech_region) if speech_region is not None else None,
                                C.byref(v), C.byref(nh), C.byref(s))
    return v.value, nh.value, s.value


def estimate_noise_floor_and_threshold(intervals):
    a, n = _arr(intervals)
    f, t = _D(), _D()
    ok = _L().jt_vad_estimate_noise_floor(a, n, C.byref(f), C.byref(t))
    return f.value, t.value, bool(ok)


def extract_noise_profile(region, intervals):
    a, n = _arr(intervals)
    p = NoiseProfile()
    if region is None or not _L().jt_vad_noise_profile(a, n, C.byref(region), C.byref(p)):
        return None
    return p


def get_intervals_in_range(intervals, start_ns, end_ns):
    a, n = _arr(intervals)
    first = _I64()
    cnt = _L().jt_vad_intervals_in_range(a, n, int(start_ns), int(end_ns), C.byref(first))
    return first.value, cnt


def score_interval_window(intervals):
    a, n = _arr(intervals)
    return _L().jt_vad_score_interval_window(a, n)


def score_speech_interval_window(intervals):
    a, n = _arr(intervals)
    return _L().jt_vad_score_speech_window(a, n)


def level_variance(intervals, axis=AXIS_MOMENTARY):
    a, n = _arr(intervals)
    return _L().jt_vad_level_variance(a, n, axis)


def score_speech_candidate_grounded(candidate, noise_floor_db, level_var):
    return _L().jt_vad_score_candidate_grounded(C.byref(candidate), noise_floor_db, level_var)


def measure_speech_candidate(region, intervals):
    a, n = _arr(intervals)
    c = SpeechCandidate()
    return c if _L().jt_vad_measure_candidate(a, n, C.byref(region), C.byref(c)) else None


def refine_to_golden_speech_subregion(candidate, intervals):
    a, n = _arr(intervals)
    out = Region()
    _L().jt_vad_refine_speech_region(a, n, C.byref(candidate), C.byref(out))
    return out


def find_best_speech_region(regions, intervals, noise_floor_db=-math.inf):
    """-> (best Region or None, [SpeechCandidate])"""
    a, n = _arr(intervals)
    rg = (Region * max(len(regions), 1))(*regions)
    cands = (SpeechCandidate * max(len(regions), 1))()
    best, nc = Region(), _I64()
    ok = _check(_L().jt_vad_find_best_speech_region(a, n, rg, len(regions), noise_floor_db, C.byref(best), cands, len(regions), C.byref(nc)))
    return (best if ok else None), list(cands[:nc.value])


def detect_voice_activity(measurements, intervals):
    """buildInputMeasurements' seed + detectVoiceActivity -> (VoiceActivity, [Region], [SpeechCandidate])"""
    a, n = _arr(intervals)
    va = VoiceActivity()
    cap = max(n, 1)
    runs, cands = (Region * cap)(), (SpeechCandidate * cap)()
    _check(_L().jt_detect_voice_activity(C.byref(measurements), a, n, C.byref(va), runs, cap, cands, cap))
    return va, list(runs[:va.n_speech_regions]), list(cands[:va.n_candidates])


def vad_detect(intervals, noise_floor_seed):
    """detectVoiceActivity proper (given seed) -> (VoiceActivity, [Region], [SpeechCandidate])"""
    a, n = _arr(intervals)
    va = VoiceActivity()
    cap = max(n, 1)
    runs, cands = (Region * cap)(), (SpeechCandidate * cap)()
    _check(_L().jt_vad_detect(a, n, noise_floor_seed, C.byref(va), runs, cap, cands, cap))
    return va, list(runs[:va.n_speech_regions]), list(cands[:va.n_candidates])


def band_plan():
    lo, hi = (_D * 17)(), (_D * 17)()
    _L().jt_band_plan(lo, hi)
    return list(lo), list(hi)


def apply_band_rms(va, speech=None, noise=None):
    """speech / noise: (rms list, found list)"""
    def pack(p, n):
        if p is None:
            return None, None
        return (_D * n)(*p[0]), (_I32 * n)(*[1 if f else 0 for f in p[1]])
    s, sf = pack(speech, 2)
    nz, nf = pack(noise, AFFTDN_BANDS)
    _check(_L().jt_apply_band_rms(C.byref(va), s, sf, nz, nf))


# ---- config -------------------------------------------------------------------------------------------------------
def default_filter_config():
    c = FilterConfig()
    _L().jt_default_filter_config(C.byref(c))
    return c


def new_measurements(**kw):
    """AudioMeasurements literal: astats keys by name (NaN = not reported, Go's zero value downstream)."""
    m = Measurements()
    for i in range(AS_COUNT):
        m.astats[i] = math.nan
    for k, v in kw.items():
        if k in AS_NAMES:
            m.astats[AS_NAMES.index(k)] = v
        else:
            setattr(m, k, v)
    return m


def adapt_config(measurements, va, base=None):
    out, diag = FilterConfig(), AdaptDiagnostics()
    _check(_L().jt_adapt_config(C.byref(base) if base is not None else None, C.byref(measurements), C.byref(va),
                                C.byref(out), C.byref(diag)))
    return out, diag


def sanitize_config(cfg):
    _L().jt_sanitize_config(C.byref(cfg))
    return cfg


def _str_call(fn, *args):
    b = C.create_string_buffer(4096)
    _check(fn(*args, b, len(b)))
    return b.value.decode()


def build_filter_spec(cfg):
    return _str_call(_L().jt_build_filter_spec, C.byref(cfg) if cfg is not None else None)


def build_filter(cfg, filter_id):
    return _str_call(_L().jt_build_filter, C.byref(cfg), filter_id)


def build_adeclick_filter(cfg):
    return _str_call(_L().jt_build_adeclick_filter, C.byref(cfg))


def go_format_g(v):
    return _str_call(_L().jt_go_format_g, float(v))


def calculate_speech_gate_threshold(voiced_low_percentile, separation):
    narrow = _INT()
    t = _L().jt_adapt_gate_threshold(voiced_low_percentile, separation, C.byref(narrow))
    return t, bool(narrow.value)


def calculate_speech_gate_threshold_no_profile(floor, room_tone_peak, room_tone_crest, ratio, lufs_gap):
    return _L().jt_adapt_gate_threshold_no_profile(floor, room_tone_peak, room_tone_crest, ratio, lufs_gap)


def build_afftdn_band_noise(bands):
    return _str_call(_L().jt_adapt_band_noise, _darr(bands), len(bands))


# ---- device entries (need a gpudsp.Context) -------------------------------------------------------------------------
def analyse_adaptive(ctx, pcm, rate, channels=1, frame_size=4096, base=None):
    """AnalyseAudio + AdaptConfig on the GPU box -> (Analysis, [Interval])"""
    import numpy as np
    pcm = np.ascontiguousarray(pcm)
    n = pcm.size // channels
    cap = int(n / rate / 0.25) + 16
    iv = (Interval * cap)()
    n_iv = _I64()
    out = Analysis()
    ctx._check(_L().jt_analyse_adaptive(ctx._h, pcm.ctypes.data_as(_P), n, rate, channels, gpudsp._FMT_OF_NP[pcm.dtype], frame_size,
                                        C.byref(base) if base is not None else None, C.byref(out), iv, cap, C.byref(n_iv)))
    return out, list(iv[:n_iv.value])


def measure_output_region(ctx, pcm, rate, start_ns, duration_ns, channels=1):
    """measureOutputRegionFromReader -> (RegionSample, frames processed)"""
    import numpy as np
    pcm = np.ascontiguousarray(pcm)
    out, frames = RegionSample(), _I64()
    ctx._check(_L().jt_measure_output_region(ctx._h, pcm.ctypes.data_as(_P), pcm.size // channels, rate, channels,
                                             gpudsp._FMT_OF_NP[pcm.dtype], int(start_ns), int(duration_ns), C.byref(out), C.byref(frames)))
    return out, frames.value


def process_audio_adaptive(ctx, pcm, rate, channels=1, base=None):
    """ProcessAudio with the adaptive Pass-2 spec -> (int16 PCM, ProcessResult, Analysis)"""
    import numpy as np
    pcm = np.ascontiguousarray(pcm)
    n = pcm.size // channels
    cap = int(n * 44100 / rate) + 3 * 4096
    out = np.empty(cap, dtype=np.int16)
    res, an = ProcessResult(), Analysis()
    ctx._check(_L().jt_process_audio_adaptive(ctx._h, pcm.ctypes.data_as(_P), n, rate, channels, gpudsp._FMT_OF_NP[pcm.dtype],
                                              C.byref(base) if base is not None else None, out.ctypes.data_as(_P), cap,
                                              C.byref(res), C.byref(an)))
    return out[:res.n_out], res, an


def process_audio_adaptive_ptr(ctx, in_ptr, n, rate, channels, fmt, out_ptr, out_cap, on_device, base=None):
    """raw-pointer variant for bench.py (torch owns the memory) -> (ProcessResult, Analysis)"""
    res, an = ProcessResult(), Analysis()
    fn = _L().jt_process_audio_adaptive_dev if on_device else _L().jt_process_audio_adaptive
    ctx._check(fn(ctx._h, _P(in_ptr), n, rate, channels, fmt, C.byref(base) if base is not None else None, _P(out_ptr), out_cap,
                  C.byref(res), C.byref(an)))
    return res, an


# ---- ONE stream over several GPUs behind one call per rank (jt_process_audio_sharded, include/jtdsp.h) ------------------------
class ShardPlan(C.Structure):
    _fields_ = [("unit", _I64), ("own_first", _I64), ("owned", _I64), ("local_first", _I64), ("n_local", _I64)]


class ShardTiming(C.Structure):
    _fields_ = [(k, C.c_double) for k in ("upload", "pass1_chunk", "pass1_merge", "adapt", "pass2_chunk", "pass2_merge", "regions", "halo",
                                           "pass3_chunk", "pass3_merge", "pass4_chunk", "pass4_merge", "download", "exchange")] + \
               [("halo_bytes", _I64), ("exchange_calls", C.c_int32), ("reserved", C.c_int32)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_ if k != "reserved"}


def sharded_plan(total_frames, rate, world, rank):
    p = ShardPlan()
    _check(_L().jt_sharded_plan(int(total_frames), int(rate), int(world), int(rank), C.byref(p)))
    return p


def process_audio_sharded(ctx, pcm_local, rate, channels, total_frames, world, rank, exchange=None, adaptive=True, base=None):
    """One rank's call of jt_process_audio_sharded.  pcm_local: this rank's window (sharded_plan), interleaved numpy array;
    exchange(send: bytes) -> bytes of all ranks in rank order (an all-gather).  Returns (owned int16 output, its first sample
    index in the whole output, ProcessResult, Analysis, ShardTiming)."""
    import numpy as np
    pcm_local = np.ascontiguousarray(pcm_local)
    n_local = pcm_local.size // channels
    plan = sharded_plan(total_frames, rate, world, rank)
    cap = int(plan.owned * 44100 / rate) + 4 * 4096 + 2 * 890820
    out = np.empty(cap, dtype=np.int16)
    res, an, tm = ProcessResult(), Analysis(), ShardTiming()
    first, n_out = _I64(), _I64()
    ctx.set_exchange(exchange, world)
    try:
        ctx._check(_L().jt_process_audio_sharded(ctx._h, pcm_local.ctypes.data_as(_P), n_local, rate, channels, gpudsp._FMT_OF_NP[pcm_local.dtype],
                                                 int(total_frames), world, rank, C.byref(base) if base is not None else None, int(bool(adaptive)),
                                                 out.ctypes.data_as(_P), cap, C.byref(first), C.byref(n_out), C.byref(res), C.byref(an), C.byref(tm)))
    finally:
        ctx.set_exchange(None, 1)
    return out[:n_out.value], first.value, res, an, tm


def process_audio_sharded_ptr(ctx, in_ptr, n_local, rate, channels, fmt, total_frames, world, rank, out_ptr, out_cap, on_device, adaptive=True, base=None):
    """raw-pointer variant (bench: torch owns the pinned / device memory); the exchange must already be installed"""
    res, an, tm = ProcessResult(), Analysis(), ShardTiming()
    first, n_out = _I64(), _I64()
    fn = _L().jt_process_audio_sharded_dev if on_device else _L().jt_process_audio_sharded
    ctx._check(fn(ctx._h, _P(in_ptr), n_local, rate, channels, fmt, int(total_frames), world, rank, C.byref(base) if base is not None else None,
                  int(bool(adaptive)), _P(out_ptr), out_cap, C.byref(first), C.byref(n_out), C.byref(res), C.byref(an), C.byref(tm)))
    return first.value, n_out.value, res, an, tm
