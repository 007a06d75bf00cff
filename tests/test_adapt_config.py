"""AdaptConfig, sanitizeConfig and BuildFilterSpec (csrc/jt_adapt.cu) against the reference's golden spec strings and
table-driven cases: /root/reference/internal/processor/adaptive_test.go and filters_test.go (cited per test).  Host-only.

The golden strings (adaptive_test.go:109-124, filters_test.go:299-311) are the reference's own known answers for the
a6 row of SURVEY 8a: they pin the printf rounding the DSP sees."""
import math

import pytest

from jivetalking_b200 import adapt as A
from jivetalking_b200 import gpudsp


def lin_to_db(x):
    return 20 * math.log10(x) if x > 0 else -1000.0


def go_test_base_config():                                   # newTestBaseConfig filters_test.go:14-72
    c = A.FilterConfig()
    c.resample_rate, c.resample_frame_size, c.resample_format = 44100, 4096, b"s16"
    hp, lp = c.rumble_highpass, c.bandlimit_lowpass
    hp.frequency, hp.poles, hp.width, hp.mix, hp.transform = 80.0, 2, 0.707, 1.0, b"tdii"
    lp.frequency, lp.poles, lp.width, lp.mix = 16000.0, 2, 0.707, 1.0
    n = c.noise_reduction
    n.strength, n.patch_s, n.research_s, n.smooth = 0.00001, 0.006, 0.0058, 11.0
    n.afftdn_enabled, n.afftdn_noise_reduction, n.afftdn_noise_type, n.afftdn_track_noise = 1, 12, b"w", 1
    g = c.speech_gate
    g.threshold, g.ratio, g.attack, g.release, g.range, g.knee, g.makeup, g.detection = 0.01, 2.0, 20, 250, 0.0625, 2.828, 1.0, b"rms"
    k = c.levelling_compressor
    k.threshold, k.ratio, k.attack, k.release, k.makeup, k.knee, k.mix = -20, 2.5, 15, 80, 0, 2.5, 1.0
    c.deesser.intensity, c.deesser.amount, c.deesser.frequency = 0.5, 0.5, 0.5
    a = c.adeclick
    a.enabled, a.threshold, a.window, a.overlap, a.method = 1, 2.0, 55.0, 50.0, b"s"
    l = c.loudnorm
    l.enabled, l.target_i, l.target_tp, l.target_lra, l.dual_mono, l.linear = 1, -16.0, -1.5, 11.0, 1, 1
    return c


def order_independence_seed():                            # adaptive_test.go:156-165
    c = go_test_base_config()
    c.rumble_highpass.enabled = c.bandlimit_lowpass.enabled = c.noise_reduction.enabled = 1
    c.speech_gate.enabled = c.levelling_compressor.enabled = 1
    c.loudnorm.target_tp = -2.0
    return c


def go_measurements(floor=0.0, voice_activated=False, input_i=0.0, input_lra=0.0, peak=0.0, rms=0.0, noise_profile=None,
                    speech_profile=None, voiced_low=0.0, noise_high=0.0, separation=0.0):
    """An AudioMeasurements literal of the reference's tests as (jt_measurements, jt_voice_activity)."""
    m = A.new_measurements(input_i=input_i, input_lra=input_lra, Dynamic_range=60.0, RMS_level=rms, Peak_level=peak)
    va = A.VoiceActivity()
    va.floor, va.voice_activated = floor, 1 if voice_activated else 0
    va.voiced_low_percentile, va.noise_high_percentile, va.gate_separation_db = voiced_low, noise_high, separation
    if noise_profile is not None:
        va.has_noise_profile = 1
        p = va.noise_profile
        p.peak_level, p.crest_factor = noise_profile.get("peak", 0.0), noise_profile.get("crest", 0.0)
        p.spectral[A.SP_NAMES.index("flatness")] = noise_profile.get("flatness", 0.0)
        p.bands_measured = 1 if noise_profile.get("bands_measured") else 0
        bands = noise_profile.get("bands", [])
        p.n_band_noise = len(bands)
        for i, b in enumerate(bands):
            p.band_noise[i] = b
    if speech_profile is not None:
        va.has_speech_profile = 1
        s = va.speech_profile
        s.sample.rms_level = speech_profile.get("rms", 0.0)
        s.body_band_rms, s.sib_band_rms = speech_profile.get("body", 0.0), speech_profile.get("sib", 0.0)
        s.bands_measured = 1 if speech_profile.get("bands_measured") else 0
    return m, va


WARM = dict(floor=-58.0, input_i=-42.1, input_lra=6.0, peak=-10.0)                                       # adaptive_test.go:167-182
BRIGHT = dict(floor=-60.0, input_i=-20.0, input_lra=12.0, peak=-6.0, rms=-30.0, noise_profile=dict(peak=-45.0, crest=15.0),
              voiced_low=-34.0, noise_high=-55.0, separation=21.0, speech_profile=dict(rms=-24.0))       # adaptive_test.go:184-224


@pytest.mark.parametrize("meas,want", [
    (WARM, "highpass=f=80:poles=2:width_type=q:width=0.707:normalize=1:a=tdii,"
           "lowpass=f=20500:poles=2:width_type=q:width=0.707:normalize=1,"
           "anlmdn=s=0.00001:p=0.0060:r=0.0058:m=11,"
           "afftdn=nr=12:nt=w:tn=0:nf=-58,"
           "agate=threshold=0.019953:ratio=2.0:attack=5.00:release=200:range=0.1995:knee=3.0:detection=rms:makeup=1.0,"
           "acompressor=threshold=0.031623:ratio=3.0:attack=10:release=200:makeup=1.00:knee=4.0:detection=rms:mix=1.00"),
    (BRIGHT, "highpass=f=80:poles=2:width_type=q:width=0.707:normalize=1:a=tdii,"
             "lowpass=f=20500:poles=2:width_type=q:width=0.707:normalize=1,"
             "anlmdn=s=0.00001:p=0.0060:r=0.0058:m=11,"
             "afftdn=nr=12:nt=w:tn=0:nf=-60,"
             "agate=threshold=0.010000:ratio=2.0:attack=5.00:release=200:range=0.1995:knee=3.0:detection=rms:makeup=1.0,"
             "acompressor=threshold=0.177828:ratio=3.0:attack=10:release=200:makeup=1.00:knee=4.0:detection=rms:mix=1.00")])
def test_adapt_config_filter_spec_golden(meas, want):      # adaptive_test.go:100-145 (the reference's golden strings)
    m, va = go_measurements(**meas)
    cfg, diag = A.adapt_config(m, va, order_independence_seed())
    assert A.build_filter_spec(cfg) == want


def test_adapt_config_returns_effective_config():          # adaptive_test.go:10-72
    m, va = go_measurements(floor=-60.0, input_i=-28.0, input_lra=9.0, peak=-8.0, noise_profile=dict())
    base = go_test_base_config()
    base.set_order([A.FILTER_DEESSER, A.FILTER_ANALYSIS])
    base.rumble_highpass.enabled, base.rumble_highpass.frequency = 1, 95.0
    base.loudnorm.target_i = -18.0
    cfg, diag = A.adapt_config(m, va, base)
    assert list(base.filter_order[:2]) == [A.FILTER_DEESSER, A.FILTER_ANALYSIS] and base.rumble_highpass.frequency == 95.0
    assert base.bandlimit_lowpass.frequency == 16000.0     # the seed is not mutated
    assert list(cfg.filter_order[:cfg.n_filter_order]) == [A.FILTER_DEESSER, A.FILTER_ANALYSIS]
    assert cfg.rumble_highpass.frequency == 95.0
    assert diag.bandlimit_lp_reason == b"20.5 kHz band-limit (always on)"


def test_adapt_config_order_independence():                # adaptive_test.go:74-98
    seed = order_independence_seed()
    A.adapt_config(*go_measurements(**WARM), seed)
    after_a, da = A.adapt_config(*go_measurements(**BRIGHT), seed)
    alone, db = A.adapt_config(*go_measurements(**BRIGHT), order_independence_seed())
    assert bytes(after_a) == bytes(alone) and bytes(da) == bytes(db)


def test_tune_bandlimit_lowpass():                         # adaptive_test.go:277-364
    cfg, diag = A.adapt_config(*go_measurements(), go_test_base_config())
    lp = cfg.bandlimit_lowpass
    assert lp.enabled and lp.frequency == 20500.0 and lp.poles == 2 and lp.mix == 1.0


@pytest.mark.parametrize("profile,want,tol", [
    (None, 0.0, 0.0),
    (dict(body=-20.0, sib=-40.0, bands_measured=True), 0.0, 0.0),
    (dict(body=-20.0, sib=-26.0, bands_measured=True), 0.0, 0.0),
    (dict(body=-20.0, sib=-24.5, bands_measured=True), 0.30, 0.001),
    (dict(body=-20.0, sib=-23.0, bands_measured=True), 0.6, 0.001),
    (dict(body=-20.0, sib=-21.5, bands_measured=True), 0.725, 0.001),
    (dict(body=-20.0, sib=-20.0, bands_measured=True), 0.85, 0.001),
    (dict(body=-20.0, sib=-16.0, bands_measured=True), 0.85, 0.001),
    (dict(body=0.0, sib=0.0, bands_measured=False), 0.0, 0.0)])
def test_tune_deesser(profile, want, tol):                 # adaptive_test.go:387-522
    cfg, _ = A.adapt_config(*go_measurements(speech_profile=profile), go_test_base_config())
    assert abs(cfg.deesser.intensity - want) <= tol


@pytest.mark.parametrize("floor,peak,crest,lra,want,tol", [(-75.0, -70.0, 10.0, 8.0, -40.0, 1.0), (-55.0, -50.0, 10.0, 12.0, -31.0, 1.0),
                                                           (-42.0, -38.0, 10.0, 8.0, -25.0, 1.0), (-55.0, -48.0, 25.0, 12.0, -45.0, 1.0),
                                                           (-20.0, -15.0, 25.0, 8.0, -25.0, 0.5)])
def test_speech_gate_threshold_no_profile(floor, peak, crest, lra, want, tol):     # adaptive_test.go:542-631
    cfg, _ = A.adapt_config(*go_measurements(floor=floor, input_lra=lra, noise_profile=dict(peak=peak, crest=crest)), go_test_base_config())
    assert abs(lin_to_db(cfg.speech_gate.threshold) - want) <= tol


@pytest.mark.parametrize("lra,want", [(18.0, 1.5), (12.0, 2.0), (6.0, 2.0), (15.0, 2.0)])
def test_speech_gate_ratio(lra, want):                     # adaptive_test.go:633-665
    cfg, _ = A.adapt_config(*go_measurements(floor=-55.0, input_lra=lra), go_test_base_config())
    assert cfg.speech_gate.ratio == want


def _fixed_gate(cfg):                                      # assertFixedGateParams adaptive_test.go:1134-1148
    g = cfg.speech_gate
    return g.attack == 5.0 and g.release == 200.0 and g.knee == 3.0 and g.detection == b"rms"


@pytest.mark.parametrize("sep,depth", [(21.0, 14.0), (8.0, 8.0), (11.9, 8.0), (12.0, 14.0)])
def test_speech_gate_range(sep, depth):                    # adaptive_test.go:763-816
    cfg, diag = A.adapt_config(*go_measurements(speech_profile=dict(rms=-20.0), voiced_low=-34.0, noise_high=-34.0 - sep, separation=sep),
                               go_test_base_config())
    assert abs(-lin_to_db(cfg.speech_gate.range) - depth) < 0.5 and cfg.speech_gate.range > 0
    assert diag.speech_gate_depth_db == depth and _fixed_gate(cfg)


def test_speech_gate_nil_noise_profile():                  # adaptive_test.go:818-839
    cfg, _ = A.adapt_config(*go_measurements(floor=-55.0, input_lra=12.0), go_test_base_config())
    assert -70 <= lin_to_db(cfg.speech_gate.threshold) <= -25 and cfg.speech_gate.detection == b"rms"
    assert abs(lin_to_db(cfg.speech_gate.threshold) - (-31.0)) < 1e-9            # -55 + 12 / (1 - 1/2)


def test_speech_gate_diagnostics():                        # adaptive_test.go:884-953
    cfg, d = A.adapt_config(*go_measurements(floor=-70.0, input_i=-48.0, input_lra=6.0, noise_profile=dict(peak=-65.0, crest=12.0),
                                             voiced_low=-35.0, noise_high=-62.0, separation=27.0, speech_profile=dict(rms=-35.0)),
                            go_test_base_config())
    assert d.speech_gate_depth_db == 14.0 and not d.speech_gate_narrow_gap
    assert d.speech_gate_quiet_speech_estimate == -35.0 and d.speech_gate_speech_separation == 27.0
    assert d.speech_gate_threshold_unclamped == -41.0 and d.speech_gate_clamp_reason == b"none"
    assert abs(d.speech_gate_speech_headroom - 6.0) < 1e-9
    cfg, d = A.adapt_config(*go_measurements(floor=-55.0, input_i=-20.0, input_lra=16.0), go_test_base_config())
    assert d.speech_gate_depth_db == 14.0 and not d.speech_gate_narrow_gap
    assert (d.speech_gate_dynamic_range, d.speech_gate_quiet_speech_estimate, d.speech_gate_speech_separation,
            d.speech_gate_speech_headroom, d.speech_gate_threshold_unclamped) == (0, 0, 0, 0, 0) and d.speech_gate_clamp_reason == b""


def test_calculate_speech_gate_threshold():                # adaptive_test.go:962-1029
    for p10, sep in ((-34.0, 26.0), (-40.0, 18.0), (-42.0, 8.0)):
        t, _ = A.calculate_speech_gate_threshold(p10, sep)
        assert abs(lin_to_db(t) - (p10 - 6.0)) < 0.01
    for sep, narrow in ((8.0, True), (11.9, True), (12.0, False), (26.0, False)):
        assert A.calculate_speech_gate_threshold(-34.0, sep)[1] is narrow
    t, narrow = A.calculate_speech_gate_threshold(-42.0, 4.0)
    assert narrow and abs(lin_to_db(t) - (-48.0)) < 0.01 and lin_to_db(t) < -40.0
    # clamps of the global gate limits (adaptive_speech_gate.go:267)
    assert abs(lin_to_db(A.calculate_speech_gate_threshold(-10.0, 30.0)[0]) - (-25.0)) < 1e-9
    assert abs(lin_to_db(A.calculate_speech_gate_threshold(-90.0, 30.0)[0]) - (-80.0)) < 1e-9


def test_tune_speech_gate_new_basis():                     # adaptive_test.go:1039-1130
    cfg, d = A.adapt_config(*go_measurements(floor=-60.0, input_i=-20.0, input_lra=12.0, speech_profile=dict(rms=-24.0), voiced_low=-34.0,
                                             noise_high=-60.0, separation=26.0), go_test_base_config())
    assert abs(lin_to_db(cfg.speech_gate.threshold) - (-40.0)) < 0.01 and abs(-lin_to_db(cfg.speech_gate.range) - 14.0) < 0.5
    assert not d.speech_gate_narrow_gap and _fixed_gate(cfg) and d.speech_gate_depth_db == 14.0
    cfg, d = A.adapt_config(*go_measurements(floor=-48.0, input_i=-30.0, input_lra=9.0, speech_profile=dict(rms=-28.0), voiced_low=-42.0,
                                             noise_high=-48.0, separation=6.0), go_test_base_config())
    assert d.speech_gate_narrow_gap and abs(lin_to_db(cfg.speech_gate.threshold) - (-48.0)) < 0.01
    assert abs(-lin_to_db(cfg.speech_gate.range) - 8.0) < 0.5 and cfg.speech_gate.range > 0 and _fixed_gate(cfg)
    assert d.speech_gate_clamp_reason == b"narrow_gap"
    cfg, d = A.adapt_config(*go_measurements(floor=-55.0, input_i=-22.0, input_lra=14.0), go_test_base_config())
    assert -80.0 <= lin_to_db(cfg.speech_gate.threshold) <= -25.0 and _fixed_gate(cfg) and d.speech_gate_depth_db == 14.0


NAN, INF = math.nan, math.inf


def test_sanitize_config():                                # adaptive_test.go:1262-1416
    c = A.FilterConfig()
    c.rumble_highpass.frequency, c.rumble_highpass.width, c.rumble_highpass.mix = 100.0, 0.5, 0.8
    c.bandlimit_lowpass.frequency, c.bandlimit_lowpass.width, c.bandlimit_lowpass.mix = 14000.0, 0.7, 0.9
    n = c.noise_reduction
    n.strength, n.patch_s, n.research_s, n.smooth, n.afftdn_noise_reduction = 0.00001, 0.006, 0.0058, 11.0, 12.0
    g = c.speech_gate
    g.threshold, g.ratio, g.attack, g.release, g.range, g.knee, g.makeup = 0.02, 2.0, 12, 250, 0.0625, 3.0, 1.0
    k = c.levelling_compressor
    k.threshold, k.ratio, k.attack, k.release, k.makeup, k.knee, k.mix = -24.0, 3.0, 10, 200, 0, 4.0, 1.0
    c.deesser.intensity, c.deesser.amount, c.deesser.frequency = 0.3, 0.5, 0.5
    before = bytes(c)
    assert bytes(A.sanitize_config(c)) == before
    # non-finite values get the defaults
    c = A.FilterConfig()
    c.rumble_highpass.frequency, c.rumble_highpass.width, c.rumble_highpass.mix = NAN, INF, -INF
    c.bandlimit_lowpass.frequency, c.bandlimit_lowpass.width, c.bandlimit_lowpass.mix = INF, NAN, -INF
    n = c.noise_reduction
    n.strength, n.patch_s, n.research_s, n.smooth, n.afftdn_noise_reduction = NAN, INF, -INF, NAN, INF
    g = c.speech_gate
    g.threshold, g.ratio, g.attack, g.release, g.range, g.knee, g.makeup = NAN, INF, -INF, NAN, INF, -INF, NAN
    k = c.levelling_compressor
    k.threshold, k.ratio, k.attack, k.release, k.makeup, k.knee, k.mix = NAN, INF, -INF, NAN, INF, -INF, NAN
    c.deesser.intensity, c.deesser.amount, c.deesser.frequency = NAN, INF, -INF
    A.sanitize_config(c)
    d = A.default_filter_config()
    assert (c.rumble_highpass.frequency, c.rumble_highpass.width, c.rumble_highpass.mix) == (80.0, 0.707, 1.0)
    assert (c.bandlimit_lowpass.frequency, c.bandlimit_lowpass.width, c.bandlimit_lowpass.mix) == (20500.0, 0.707, 1.0)
    dn = d.noise_reduction
    assert (n.strength, n.patch_s, n.research_s, n.smooth, n.afftdn_noise_reduction) == \
           (dn.strength, dn.patch_s, dn.research_s, dn.smooth, dn.afftdn_noise_reduction)
    assert not n.enabled and not n.afftdn_enabled and n.afftdn_noise_type == b"" and not n.afftdn_track_noise
    dg = d.speech_gate
    assert (g.threshold, g.ratio, g.attack, g.release, g.range, g.knee, g.makeup) == \
           (0.01, dg.ratio, dg.attack, dg.release, dg.range, dg.knee, dg.makeup) and g.detection == b""
    dk = d.levelling_compressor
    assert (k.threshold, k.ratio, k.attack, k.release, k.makeup, k.knee, k.mix) == \
           (dk.threshold, dk.ratio, dk.attack, dk.release, dk.makeup, dk.knee, dk.mix)
    assert (c.deesser.intensity, c.deesser.amount, c.deesser.frequency) == (0.0, 0.50, 0.80)
    for thr in (NAN, INF, -INF, 0.0, -0.5):
        c = A.FilterConfig()
        c.speech_gate.threshold = thr
        assert A.sanitize_config(c).speech_gate.threshold == 0.01
    c = A.FilterConfig()
    c.speech_gate.threshold = 1e-10
    A.sanitize_config(c)
    assert c.speech_gate.threshold == 1e-10 and c.rumble_highpass.frequency == 0.0 and c.levelling_compressor.threshold == 0.0
    c = A.FilterConfig()
    c.levelling_compressor.threshold, c.speech_gate.threshold = -40.0, 0.02
    assert A.sanitize_config(c).levelling_compressor.threshold == -40.0
    # custom noise type without a shape reverts to white (filters_test.go:868-879)
    c = A.default_filter_config()
    c.noise_reduction.afftdn_noise_type = b"custom"
    assert A.sanitize_config(c).noise_reduction.afftdn_noise_type == b"w"


@pytest.mark.parametrize("kw,want", [
    (dict(peak=-6.0, rms=-32.0, speech_profile=dict(rms=-24.0)), -15.0),          # adaptive_test.go:1418-1431
    (dict(rms=-20.0, speech_profile=dict(rms=-10.0)), -6.0),                      # :1433-1446
    (dict(rms=NAN, speech_profile=dict(rms=-60.0)), -45.0),                       # :1448-1462
    (dict(peak=-6.0), -26.0),                                                     # :1464-1476
    (dict(peak=0.0), -20.0),                                                      # :1478-1489
    (dict(peak=NAN), -18.0),                                                      # :1491-1502
    (dict(rms=-40.0, speech_profile=dict(rms=-24.0)), -15.0),                     # :1504-1578
    (dict(rms=-40.0, speech_profile=dict(rms=-50.0)), -31.0),
    (dict(rms=NAN, speech_profile=dict(rms=-24.0)), -15.0),
    (dict(rms=INF, speech_profile=dict(rms=-24.0)), -15.0),
    (dict(rms=-8.0, speech_profile=dict(rms=-50.0)), -6.0),
    (dict(rms=0.0, speech_profile=dict(rms=-24.0)), -15.0),
    (dict(rms=-INF, speech_profile=dict(rms=-24.0)), -15.0)])
def test_levelling_compressor_threshold(kw, want):
    cfg, _ = A.adapt_config(*go_measurements(**kw), go_test_base_config())
    k = cfg.levelling_compressor
    assert abs(k.threshold - want) < 0.001
    assert (k.ratio, k.attack, k.release, k.knee, k.mix, k.makeup) == (3.0, 10.0, 200.0, 4.0, 1.0, 0.0)


def _nr(**kw):
    base = A.FilterConfig()
    d = A.default_filter_config()
    base.noise_reduction = d.noise_reduction
    return A.adapt_config(*go_measurements(**kw), base)


QUAL = dict(floor=-58.0, separation=15.0, noise_profile=dict(flatness=0.6, bands_measured=True, bands=[-61.0, -60.0, -59.0]))


def test_tune_noise_reduction():                           # adaptive_test.go:1748-1964
    cfg, d = _nr(floor=-58.0, voice_activated=True)
    assert not cfg.noise_reduction.afftdn_enabled and not d.afftdn_enabled and d.afftdn_disable_reason == b"voice_activated"
    assert cfg.noise_reduction.afftdn_noise_floor == 0
    cfg, d = _nr(floor=-58.0)
    n = cfg.noise_reduction
    assert n.afftdn_enabled and n.afftdn_noise_floor == -58.0 and not n.afftdn_track_noise and d.afftdn_noise_floor_db == -58.0 and d.afftdn_enabled
    assert _nr(floor=-120.0)[0].noise_reduction.afftdn_noise_floor == -80.0
    assert _nr(floor=-5.0)[0].noise_reduction.afftdn_noise_floor == -20.0
    n = _nr(floor=0.0)[0].noise_reduction
    assert n.afftdn_enabled and n.afftdn_track_noise and n.afftdn_noise_floor == 0
    cfg, d = _nr(**QUAL)
    n = cfg.noise_reduction
    assert n.afftdn_noise_type == b"custom" and n.afftdn_band_noise == b"-1.0|0.0|1.0" and n.afftdn_noise_floor == -58.0
    assert not n.afftdn_track_noise and d.afftdn_noise_type == b"custom"
    q = dict(QUAL, noise_profile=dict(QUAL["noise_profile"], bands=[-61.0, -60.0, -59.0, NAN]))
    n = _nr(**q)[0].noise_reduction
    assert n.afftdn_noise_type == b"custom" and n.afftdn_band_noise == b"-1.0|0.0|1.0|0.0"
    q = dict(QUAL, noise_profile=dict(QUAL["noise_profile"], bands=[NAN, -INF, INF]))
    n = _nr(**q)[0].noise_reduction
    assert n.afftdn_noise_type == b"w" and n.afftdn_band_noise == b""
    for mutate in (dict(noise_profile=dict(QUAL["noise_profile"], bands_measured=False)), dict(separation=11.0),
                   dict(noise_profile=dict(QUAL["noise_profile"], flatness=0.40)), dict(noise_profile=None)):
        n = _nr(**dict(QUAL, **mutate))[0].noise_reduction
        assert n.afftdn_noise_type == b"w" and n.afftdn_band_noise == b""


@pytest.mark.parametrize("bands,want", [([], ""), ([-50.0, -40.0, -30.0], "-10.0|0.0|10.0"), ([-100.0, 0.0], "-24.0|24.0"),
                                        ([-50.0, -40.0, -30.0, NAN], "-10.0|0.0|10.0|0.0"), ([-50.0, -INF, -30.0], "-10.0|0.0|10.0"),
                                        ([-120.0, -40.0, -40.0], "-24.0|24.0|24.0"), ([NAN, INF, -INF], "")])
def test_build_afftdn_band_noise(bands, want):             # adaptive_test.go:1967-2031
    assert A.build_afftdn_band_noise(bands) == want


# ---- filters_test.go --------------------------------------------------------------------------------------------
def test_default_filter_config():                          # filters_test.go:80-116, filters.go:421-532
    d = A.default_filter_config()
    assert d.downmix_enabled and d.analysis_enabled and d.resample_enabled and (d.resample_rate, d.resample_frame_size) == (44100, 4096)
    assert d.speech_gate.range == 10 ** (-14 / 20) and d.deesser.frequency == 0.80 and d.adeclick.threshold == 1.7
    assert (d.loudnorm.target_i, d.loudnorm.target_tp, d.loudnorm.target_lra) == (-16.0, -1.0, 20.0)


def test_build_filter_spec_basics():                       # filters_test.go:118-289
    c = go_test_base_config()
    assert A.build_filter_spec(c) == ""
    c.resample_enabled = 1
    spec = A.build_filter_spec(c)
    assert "aformat=sample_rates=44100" in spec and "asetnsamples=n=4096" in spec
    assert not any(p in spec for p in ("highpass=", "anlmdn=", "agate=", "acompressor=", "alimiter="))
    assert "adeclick=" not in A.build_filter_spec(A.default_filter_config())
    c = go_test_base_config()
    c.rumble_highpass.enabled = c.speech_gate.enabled = c.levelling_compressor.enabled = c.deesser.enabled = c.resample_enabled = 1
    spec = A.build_filter_spec(c)
    for p in ("highpass=f=", "agate=threshold=", "acompressor=threshold=", "deesser=i=", "aformat=sample_rates=44100"):
        assert p in spec
    assert "NaN" not in spec and "Inf" not in spec and "inf" not in spec
    c = go_test_base_config()
    c.deesser.enabled, c.deesser.intensity = 1, 0.0
    assert "deesser=" not in A.build_filter_spec(c)
    c = go_test_base_config()
    c.analysis_enabled = c.resample_enabled = 1
    spec = A.build_filter_spec(c)
    assert spec.index("ebur128=") < spec.index("aformat=sample_rates=44100") < spec.index("asetnsamples=")
    assert A.build_filter_spec(None) == ""


def _only(filter_id, **setup):
    c = go_test_base_config()
    c.set_order([filter_id])
    return c


def test_build_filter_spec_golden():                       # filters_test.go:291-429 (the reference's golden strings)
    assert A.build_filter_spec(A.default_filter_config()) == (
        "aformat=channel_layouts=mono,"
        "highpass=f=80:poles=2:width_type=q:width=0.707:normalize=1:a=tdii,"
        "lowpass=f=20500:poles=2:width_type=q:width=0.707:normalize=1:a=tdii,"
        "anlmdn=s=0.00001:p=0.0060:r=0.0020:m=3,"
        "afftdn=nr=12:nt=w:tn=1,"
        "agate=threshold=0.010000:ratio=2.0:attack=5.00:release=200:range=0.1995:knee=3.0:detection=rms:makeup=1.0,"
        "acompressor=threshold=0.125893:ratio=3.0:attack=10:release=200:makeup=1.00:knee=4.0:detection=rms:mix=1.00,"
        "astats=metadata=1:measure_perchannel=all,"
        "aspectralstats=win_size=2048:win_func=hann:measure=all,"
        "ebur128=metadata=1:peak=sample+true:dualmono=true:target=-16,"
        "aformat=sample_rates=44100:channel_layouts=mono:sample_fmts=s16,asetnsamples=n=4096")
    assert A.build_filter_spec(A.default_filter_config()) == gpudsp.default_pass2_spec()
    c = _only(A.FILTER_BANDLIMIT_LOWPASS)
    assert A.build_filter_spec(c) == ""
    lp = c.bandlimit_lowpass
    lp.enabled, lp.frequency, lp.poles, lp.width, lp.mix, lp.transform = 1, 14500.0, 1, 0.5, 0.75, b"zdf"
    assert A.build_filter_spec(c) == "lowpass=f=14500:poles=1:width_type=q:width=0.500:normalize=1:a=zdf:m=0.75"
    c = _only(A.FILTER_SPEECH_GATE)
    g = c.speech_gate
    g.enabled, g.threshold, g.ratio, g.attack, g.release, g.range, g.knee, g.detection, g.makeup = 1, 0.003162, 3.5, 10.5, 425, 0.0316, 4.5, b"peak", 1.2
    assert A.build_filter_spec(c) == "agate=threshold=0.003162:ratio=3.5:attack=10.50:release=425:range=0.0316:knee=4.5:detection=peak:makeup=1.2"
    c = _only(A.FILTER_LEVELLING_COMPRESSOR)
    k = c.levelling_compressor
    k.enabled, k.threshold, k.ratio, k.attack, k.release, k.makeup, k.knee, k.mix = 1, -30.0, 4.0, 10, 60, 0, 6.0, 0.85
    assert A.build_filter_spec(c) == "acompressor=threshold=0.031623:ratio=4.0:attack=10:release=60:makeup=1.00:knee=6.0:detection=rms:mix=0.85"
    c = _only(A.FILTER_NOISE_REDUCTION)
    c.noise_reduction.enabled, c.noise_reduction.afftdn_enabled = 1, 0
    assert A.build_filter_spec(c) == "anlmdn=s=0.00001:p=0.0060:r=0.0058:m=11"
    c.noise_reduction.afftdn_enabled = 1
    assert A.build_filter_spec(c) == "anlmdn=s=0.00001:p=0.0060:r=0.0058:m=11,afftdn=nr=12:nt=w:tn=1"
    c = _only(A.FILTER_DEESSER)
    c.deesser.enabled, c.deesser.intensity = 1, 0
    assert A.build_filter_spec(c) == ""
    c.deesser.intensity, c.deesser.amount, c.deesser.frequency = 0.6, 0.4, 0.7
    assert A.build_filter_spec(c) == "deesser=i=0.60:m=0.40:f=0.70"


def test_single_filter_builders():                         # filters_test.go:468-923
    c = go_test_base_config()
    assert A.build_filter(c, A.FILTER_RUMBLE_HIGHPASS) == ""
    c.rumble_highpass.enabled = 1
    assert "highpass=f=80:" in A.build_filter(c, A.FILTER_RUMBLE_HIGHPASS)
    for thr, want in ((0.01, "agate=threshold=0.010"), (0.001, "agate=threshold=0.001"), (0.05, "agate=threshold=0.050")):
        c = go_test_base_config()
        c.speech_gate.enabled, c.speech_gate.threshold = 1, thr
        s = A.build_filter(c, A.FILTER_SPEECH_GATE)
        assert want in s and "detection=rms" in s
    assert A.build_filter(go_test_base_config(), A.FILTER_SPEECH_GATE) == ""
    for f in (16000.0, 12000.0, 14500.0):
        c = go_test_base_config()
        c.bandlimit_lowpass.enabled, c.bandlimit_lowpass.frequency = 1, f
        assert f"lowpass=f={f:.0f}:" in A.build_filter(c, A.FILTER_BANDLIMIT_LOWPASS)
    c = go_test_base_config()
    c.levelling_compressor.enabled, c.levelling_compressor.threshold, c.levelling_compressor.ratio = 1, -20.0, 2.5
    s = A.build_filter(c, A.FILTER_LEVELLING_COMPRESSOR)
    assert "acompressor=threshold=" in s and "ratio=2.5" in s and "detection=rms" in s and "threshold=-" not in s
    for en, i, want in ((1, 0.5, "deesser=i=0.50"), (1, 0.8, "deesser=i=0.80"), (0, 0.5, None), (1, 0.0, None), (1, -0.1, None)):
        c = go_test_base_config()
        c.deesser.enabled, c.deesser.intensity = en, i
        s = A.build_filter(c, A.FILTER_DEESSER)
        assert (s == "") if want is None else (want in s)
    # afftdn clause (filters_test.go:797-879)
    c = A.default_filter_config()
    c.set_order([A.FILTER_NOISE_REDUCTION])
    n = c.noise_reduction
    assert A.build_filter_spec(c).endswith("afftdn=nr=12:nt=w:tn=1")
    n.afftdn_noise_floor, n.afftdn_track_noise = -58.0, 0
    assert A.build_filter_spec(c).endswith("afftdn=nr=12:nt=w:tn=0:nf=-58")
    n.afftdn_noise_type, n.afftdn_band_noise = b"custom", b"0.0|3.5|-2.0"
    assert A.build_filter_spec(c).endswith("afftdn=nr=12:nt=custom:bn=0.0|3.5|-2.0:tn=0:nf=-58")
    n.afftdn_band_noise = b""
    assert "bn=" not in A.build_filter_spec(c)
    n.afftdn_enabled = 0
    assert "afftdn" not in A.build_filter_spec(c)
    # the measured floor keeps every digit the reference's %g prints (shortest round-trip, not C's six digits)
    c = A.default_filter_config()
    c.set_order([A.FILTER_NOISE_REDUCTION])
    c.noise_reduction.afftdn_noise_floor = -52.37421875
    assert A.build_filter_spec(c).endswith(":nf=-52.37421875")
    c.noise_reduction.afftdn_noise_floor = -61.3
    assert A.build_filter_spec(c).endswith(":nf=-61.3")


def test_build_adeclick_filter():                          # filters_test.go:925-989
    assert A.build_adeclick_filter(A.default_filter_config()) == "adeclick=t=1.7:w=55:o=50:m=s"
    c = go_test_base_config()
    c.adeclick.window = 100.0
    s = A.build_adeclick_filter(c)
    assert all(p in s for p in ("adeclick=", "t=2.0", "w=100", "o=50", "m=s"))
    c = go_test_base_config()
    c.adeclick.method = b""
    assert A.build_adeclick_filter(c) == "adeclick=t=2.0:w=55:o=50"
    c.adeclick.enabled = 0
    assert A.build_adeclick_filter(c) == ""


def test_filter_order_respected():                         # filters_test.go:991-1019, :1588-1650
    c = go_test_base_config()
    c.rumble_highpass.enabled = c.speech_gate.enabled = c.deesser.enabled = c.resample_enabled = 1
    c.deesser.intensity = 0.5
    spec = A.build_filter_spec(c)
    assert spec.index("highpass=") < spec.index("agate=") < spec.index("deesser=") < spec.index("aformat=sample_rates=")
    d = A.default_filter_config()
    d.set_order(A.PASS1_ORDER)
    assert A.build_filter_spec(d) == gpudsp.pass1_spec()


@pytest.mark.parametrize("v,want", [(12.0, "12"), (-58.0, "-58"), (-52.37421875, "-52.37421875"), (0.5, "0.5"), (1e6, "1e+06"),
                                    (123456789.0, "1.23456789e+08"), (1e-5, "1e-05"), (100000.0, "100000"), (0.0001, "0.0001"),
                                    (-79.99999999999999, "-79.99999999999999"), (0.0, "0")])
def test_go_format_g(v, want):                             # fmt %g = strconv 'g', shortest (filters.go:806-826)
    assert A.go_format_g(v) == want


def test_band_plan():                                      # analyser_bands.go:19-24,98-103; analyser_noise_bands.go:15-52
    lo, hi = A.band_plan()
    assert (lo[0], hi[0], lo[1], hi[1]) == (1000.0, 3000.0, 6000.0, 9000.0)
    c = [80, 125, 195, 290, 440, 660, 1000, 1500, 2250, 3350, 5000, 7500, 11200, 16000, 24000]
    assert lo[2] == c[0] / math.sqrt(c[1] / c[0]) and hi[16] == c[14] * math.sqrt(c[14] / c[13])
    for i in range(1, 15):
        assert lo[2 + i] == math.sqrt(c[i - 1] * c[i]) == hi[1 + i]


def test_apply_band_rms():                                 # analyser_bands.go:150-162, analyser_noise_bands.go:94-118
    va = A.VoiceActivity()
    va.has_speech_profile, va.has_noise_profile = 1, 1
    va.speech_profile.region = A.Region.of(0, 30 * A.NS_S)
    va.noise_profile.duration_ns = 10 * A.NS_S
    noise = [-70.0 - i for i in range(14)] + [math.nan]
    A.apply_band_rms(va, ([-30.0, -36.0], [1, 1]), (noise, [1] * 15))
    assert va.speech_profile.bands_measured and (va.speech_profile.body_band_rms, va.speech_profile.sib_band_rms) == (-30.0, -36.0)
    assert va.noise_profile.bands_measured and va.noise_profile.n_band_noise == 15
    A.apply_band_rms(va, ([-30.0, 0.0], [1, 0]), ([math.nan] * 6 + noise[:9], [1] * 15))
    assert not va.speech_profile.bands_measured and not va.noise_profile.bands_measured     # 9 finite bands < afftdnMinFiniteBands


def test_abi_argument_validation():
    """whole-call failure with a negative code, never a partial result (SURVEY 8b error convention)"""
    import ctypes as C
    L = A._L()
    m = A.new_measurements(input_i=-20.0)
    va = A.VoiceActivity()
    assert L.jt_detect_voice_activity(None, None, 0, C.byref(va), None, 0, None, 0) == -1
    assert L.jt_detect_voice_activity(C.byref(m), None, 5, C.byref(va), None, 0, None, 0) == -1
    assert L.jt_adapt_config(None, None, C.byref(va), C.byref(A.FilterConfig()), None) == -1
    buf = C.create_string_buffer(16)
    assert L.jt_build_filter_spec(C.byref(A.default_filter_config()), buf, 16) == -7          # JT_ERR_BUFFER
    # a short regions buffer: JT_ERR_BUFFER, and nothing written to *out
    # two speech runs separated by a loud interval the spectral veto rejects (the loud-gap guard, analyser_vad.go:531-536), then room tone
    iv = [A.interval(i * A.HOP_NS, rms=-15.0, momentary=-15.0, centroid=9000.0 if i == 60 else 2000.0, entropy=0.4) for i in range(121)]
    iv += [A.interval((121 + i) * A.HOP_NS, rms=-60.0, momentary=-60.0, centroid=2000.0, entropy=0.4) for i in range(60)]
    arr = (A.Interval * len(iv))(*iv)
    ok_va, runs, _ = A.detect_voice_activity(m, iv)
    assert len(runs) >= 2
    out = A.VoiceActivity()
    one = (A.Region * 1)()
    assert L.jt_detect_voice_activity(C.byref(m), arr, len(iv), C.byref(out), one, 1, None, 0) == -7
    assert out.n_speech_regions == 0 and out.split == 0.0
    # empty stream: the detector degrades to "unmeasured" (floor 0 -> AdaptConfig keeps the safe defaults, adaptive.go:151-155)
    va0, runs0, cands0 = A.detect_voice_activity(m, [])
    assert (va0.floor, va0.n_speech_regions, va0.has_noise_profile, va0.has_speech_profile) == (0.0, 0, 0, 0) and not runs0 and not cands0
    cfg, _ = A.adapt_config(m, va0)
    assert "afftdn=nr=12:nt=w:tn=1," in A.build_filter_spec(cfg)
