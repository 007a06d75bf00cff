"""Test-side composition of the CPU oracle (oracle/) into the sink-frame metadata wire the
reference reads (analyser_metrics.go:432-483), independent of the product's C++ executor.
Frame cadence follows libavfilter's re-framing rules as described in SURVEY.md 8a-bis:
4096-sample decoder frames -> astats stamps (cumulative at frame end) -> aspectralstats
re-frames to 1024-sample hops (props of the frame holding the hop's first sample) ->
ebur128 re-frames to 100 ms (props of the hop frame holding the tick's first sample)."""
import math
import numpy as np
import jt_oracle as O

NAN = float("nan")
AS_NAMES = ["Dynamic_range", "RMS_level", "Peak_level", "RMS_trough", "RMS_peak", "DC_offset", "Flat_factor",
            "Crest_factor", "Zero_crossings_rate", "Zero_crossings", "Max_difference", "Min_difference",
            "Mean_difference", "RMS_difference", "Entropy", "Min_level", "Max_level", "Noise_floor",
            "Noise_floor_count", "Bit_depth", "Number_of_samples"]


def wire(fmt, v):
    """what strconv.ParseFloat reads back from FFmpeg's snprintf"""
    return float(fmt % v)


def downmix(x, channels):
    x = np.ascontiguousarray(x)
    if channels == 1:
        return x
    assert channels == 2
    n = x.size // 2
    out = np.zeros(n, dtype=x.dtype)
    fn = {np.dtype(np.float32): O.lib().orc_downmix_stereo_f32, np.dtype(np.int16): O.lib().orc_downmix_stereo_s16}[x.dtype]
    fn(x.ctypes.data_as(O._P), n, out.ctypes.data_as(O._P))
    return out


def to_f32(x):
    if x.dtype == np.float32:
        return x
    if x.dtype == np.int16:
        return (x.astype(np.float32) * np.float32(1.0 / 32768.0)).astype(np.float32)
    return x.astype(np.float32)


def to_f64(x):
    if x.dtype == np.int16:
        return x.astype(np.float64) * (1.0 / 32768.0)
    return x.astype(np.float64)


def analysis_meta(astats_in, spec_in_f32, r128_in_f64, rate, src_frame_bounds, dualmono=True, true_peak=True,
                  astats_overall_only=False):
    """Sink-frame records of `astats -> aspectralstats -> ebur128` given the signal each filter
    sees and the frame boundaries (ascending end positions) on the astats link."""
    n = len(r128_in_f64)
    T = rate // 10
    rows = O.aspectralstats(spec_in_f32, rate)
    r = O.ebur128(r128_in_f64, rate, dualmono=dualmono, true_peak=true_peak)
    ends = np.asarray(src_frame_bounds, dtype=np.int64)
    starts = np.concatenate(([0], ends[:-1]))

    def src_frame_of(sample):
        return int(np.searchsorted(starts, sample, side="right") - 1)

    recs = []
    nticks_total = (n + T - 1) // T
    last_tick = n // T - 1
    for k in range(nticks_total):
        s0 = k * T
        nb = min(T, n - s0)
        j = s0 // 1024                                   # hop frame holding the first sample
        f = src_frame_of(j * 1024)                       # src frame holding that hop's first sample
        last = s0 + nb - 1
        j2 = last // 1024
        hop_last = min(1024 * (j2 + 1), n) - 1
        ready = int(ends[src_frame_of(hop_last)])
        rec = dict(first_sample=s0, nb_samples=nb, ready=ready, astats_pos=int(ends[f]), hop=j,
                   M=NAN, S=NAN, I=NAN, LRA=NAN, true_peak=NAN, sample_peak=NAN,
                   spectral=[wire("%g", float(v)) for v in rows[j]], astats=None)
        if nb == T and k < r["n_ticks"]:
            rec.update(M=wire("%.3f", r["M"][k]), S=wire("%.3f", r["S"][k]),
                       sample_peak=wire("%.3f", r["sample_peak_cum"][k]))
            if true_peak:
                rec.update(true_peak=wire("%.3f", r["true_peak_cum"][k]))
            if k == last_tick:
                rec.update(I=wire("%.3f", r["I"]), LRA=wire("%.3f", r["LRA"]))
        recs.append(rec)
    if recs:
        a = O.astats(astats_in[: recs[-1]["astats_pos"]], rate)
        recs[-1]["astats"] = {k: wire("%f", a[k]) for k in AS_NAMES if k != "Number_of_samples"}
        recs[-1]["astats"]["Number_of_samples"] = a["nb_samples"]
        recs[-1]["overall_only"] = astats_overall_only
    return recs


def pass1_meta(x, rate, channels=1, frame_size=4096):
    mono = downmix(x, channels)
    n = len(mono)
    ends = [min((f + 1) * frame_size, n) for f in range((n + frame_size - 1) // frame_size)]
    return analysis_meta(mono, to_f32(mono), to_f64(mono), rate, ends)


def _close(a, b, atol, rtol=0.0):
    if isinstance(a, float) and math.isnan(a):
        return isinstance(b, float) and math.isnan(b)
    if math.isinf(a) or math.isinf(b):
        return a == b
    return abs(a - b) <= atol + rtol * abs(b)


ASTATS_TOL = {  # (atol, rtol) on the "%f"-printed values
    "Noise_floor_count": (1e9, 0.0),     # ties on equal window maxima are float-representation dependent
    "Entropy": (2e-6, 0), "Zero_crossings": (0, 0), "Bit_depth": (0, 0), "Number_of_samples": (0, 0),
}


def assert_meta_close(got, exp, spectral_rtol=2e-3, spectral_atol=1e-9):
    """got: list of gpudsp.FrameMeta; exp: list of dicts from analysis_meta()."""
    assert len(got) == len(exp), (len(got), len(exp))
    for i, (g, e) in enumerate(zip(got, exp)):
        assert g.first_sample == e["first_sample"] and g.nb_samples == e["nb_samples"], (i, g.first_sample, e)
        for name, gv, ev in (("M", g.r128_M, e["M"]), ("S", g.r128_S, e["S"]), ("I", g.r128_I, e["I"]),
                             ("LRA", g.r128_LRA, e["LRA"])):
            assert _close(gv, ev, 0.0011), (i, name, gv, ev)
        for name, gv, ev in (("true_peak", g.r128_true_peak, e["true_peak"]),
                             ("sample_peak", g.r128_sample_peak, e["sample_peak"])):
            assert _close(gv, ev, 0.0011), (i, name, gv, ev)
        for k in range(13):
            assert _close(g.spectral[k], e["spectral"][k], spectral_atol, spectral_rtol), (i, "spectral", k, g.spectral[k], e["spectral"][k])
        if e["astats"] is None:
            assert all(math.isnan(g.astats[k]) for k in range(len(AS_NAMES))), (i, "unexpected astats")
        else:
            if not e.get("overall_only"):
                for k, name in enumerate(AS_NAMES):
                    atol, rtol = ASTATS_TOL.get(name, (2e-6, 1e-9))
                    assert _close(g.astats[k], e["astats"][name], atol, rtol), (i, name, g.astats[k], e["astats"][name])
            assert _close(g.astats_overall_RMS_level, e["astats"]["RMS_level"], 2e-6), (i, "overall rms")
            assert _close(g.astats_overall_Peak_level, e["astats"]["Peak_level"], 2e-6), (i, "overall peak")
