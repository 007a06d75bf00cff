"""Seeded differential test of the Pass 1 -> Pass 2 host logic: libjtdsp's C++ (csrc/jt_adapt.cu, through the C ABI) against
the oracle's plain-Python restatement (oracle/adapt_oracle.py) on random conversational interval streams -- speech runs
of random level, room-tone gaps, loud non-speech interruptions, digital-silence stretches, NaN / -inf momentary values.
Complements the table tests (tests/test_adapt_*.py replay the reference's own cases); here every intermediate of the
detector and the final Pass-2 spec string must agree on inputs nobody hand-picked.  Host-only."""
import math
import random

import pytest

import adapt_oracle as AO
from jivetalking_b200 import adapt as A

HOP = A.HOP_NS


def random_stream(seed):
    rnd = random.Random(seed)
    ivs, t = [], 0
    n_target = rnd.randint(5, 1500)
    frame = rnd.choice([4096 / 48000, 4096 / 44100, 0.25, 0.1])            # interval timestamps follow decoder frames
    step_ns = int(math.ceil(0.25 / frame)) * int(frame * 1e9) if frame < 0.25 else int(frame * 1e9)
    noise = rnd.uniform(-75, -45)
    gated = rnd.random() < 0.2
    veto_fail_p = rnd.choice([0.0, 0.0, 0.005, 0.03])
    while len(ivs) < n_target:
        kind = rnd.choices(["speech", "gap", "loud", "silence"], [5, 3, 0.5, 1.0 if gated else 0.1])[0]
        length = {"speech": rnd.randint(4, 400), "gap": rnd.randint(1, 80), "loud": rnd.randint(1, 6), "silence": rnd.randint(1, 60)}[kind]
        base = rnd.uniform(-40, -12)
        for _ in range(length):
            if kind == "speech":
                m = base + rnd.gauss(0, 2.0)
                spec = dict(centroid=rnd.uniform(300, 5500), entropy=rnd.uniform(0.2, 0.68), kurtosis=rnd.uniform(1, 12),
                            flatness=rnd.uniform(0.05, 0.6), rolloff=rnd.uniform(2000, 11000), flux=rnd.uniform(0.0005, 0.04))
                if rnd.random() < veto_fail_p:                              # an interval the spectral veto rejects inside a phrase
                    spec.update(centroid=rnd.choice([120.0, 7000.0]), entropy=rnd.uniform(0.71, 0.9))
                if rnd.random() < 0.05:
                    m = base - rnd.uniform(3, 25)                           # a breath: below or around the split
            elif kind == "loud":
                m = rnd.uniform(-25, -10)
                spec = dict(centroid=rnd.uniform(6500, 12000), entropy=rnd.uniform(0.6, 0.95), kurtosis=rnd.uniform(1, 4),
                            flatness=rnd.uniform(0.4, 0.9), rolloff=rnd.uniform(9000, 16000), flux=rnd.uniform(0.01, 0.08))
            elif kind == "gap":
                m = noise + rnd.gauss(0, 1.0)
                spec = dict(centroid=rnd.uniform(3000, 12000), entropy=rnd.uniform(0.7, 0.99), kurtosis=rnd.uniform(1, 3.5),
                            flatness=rnd.uniform(0.3, 0.95), rolloff=rnd.uniform(8000, 20000), flux=rnd.uniform(0.0001, 0.01))
            else:
                m = rnd.choice([-math.inf, math.nan, -120.0, -117.7, -130.0])
                spec = dict(centroid=0.0, entropy=0.0, kurtosis=0.0, flatness=0.0, rolloff=0.0, flux=0.0)
            rms = (m if math.isfinite(m) else -120.0) - rnd.uniform(0, 4)
            spec.update(mean=rnd.uniform(0, 1e-3), variance=rnd.uniform(0, 1e-6), spread=rnd.uniform(500, 6000), skewness=rnd.uniform(-1, 5),
                        crest=rnd.uniform(5, 60), slope=rnd.uniform(-1e-4, 0), decrease=rnd.uniform(-0.2, 0.2))
            ivs.append(dict(ts=t, rms=rms, peak=rms + rnd.uniform(3, 20), M=m, S=(m if math.isfinite(m) else -120.0) - rnd.uniform(0, 1),
                            tp=rms + rnd.uniform(3, 21), sp=rms + rnd.uniform(3, 20), spectral=spec))
            t += step_ns
    return ivs[:n_target]


def to_c(ivs):
    out = []
    for d in ivs:
        iv = A.interval(d["ts"], rms=d["rms"], peak=d["peak"], momentary=d["M"], short_term=d["S"], true_peak=d["tp"], sample_peak=d["sp"])
        for k, v in d["spectral"].items():
            iv.spectral[A.SP_NAMES.index(k)] = v
        out.append(iv)
    return out


def same(a, b):
    return (a != a and b != b) or a == b or abs(a - b) <= 1e-12 * max(1.0, abs(a), abs(b))


def check_sample(c, o):
    assert same(c.rms_level, o["rms"]) and same(c.peak_level, o["peak"]) and same(c.crest_factor, o["crest"])
    assert same(c.momentary_lufs, o["M"]) and same(c.short_term_lufs, o["S"]) and same(c.true_peak, o["tp"]) and same(c.sample_peak, o["sp"])
    for k, name in enumerate(A.SP_NAMES):
        assert same(c.spectral[k], o["spectral"][name]), name


@pytest.mark.parametrize("seed", range(120))
def test_detector_and_spec_agree_with_the_oracle(seed):
    ivs = random_stream(seed)
    civ = to_c(ivs)
    rnd = random.Random(1000 + seed)
    meas = dict(input_i=rnd.uniform(-45, -12), input_lra=rnd.uniform(2, 22), rms_level=rnd.choice([0.0, rnd.uniform(-50, -15)]),
                peak_level=rnd.uniform(-20, 0))
    m = A.new_measurements(input_i=meas["input_i"], input_lra=meas["input_lra"], Dynamic_range=50.0, RMS_level=meas["rms_level"],
                           Peak_level=meas["peak_level"], Noise_floor=-70.0)
    va, runs, cands = A.detect_voice_activity(m, civ)
    o = AO.detect_full(ivs)
    # seed and split machinery
    f, t, ok = A.estimate_noise_floor_and_threshold(civ)
    of, ot, ook = AO.estimate_noise_floor(ivs)
    assert ok == ook and same(f, of) and same(t, ot)
    assert same(va.floor_prescan, o["prescan"]) and same(va.room_tone_detect_level, o["detect_level"])
    for k, ok_ in (("split", "split"), ("floor", "floor"), ("margin", "margin")):
        assert same(getattr(va, k), o[ok_]), (k, getattr(va, k), o[ok_])
    assert va.gap_tolerance == o["tol"]
    assert [(r.start_ns, r.end_ns) for r in runs] == o["runs"]
    assert all(r.duration_ns == r.end_ns - r.start_ns for r in runs)
    # election
    assert len(cands) == len(o["cands"])
    for c, oc in zip(cands, o["cands"]):
        assert (c.region.start_ns, c.region.end_ns) == oc["region"] and same(c.score, oc["score"]) and same(c.voicing_density, oc["voicing"])
        assert bool(c.was_refined) == oc["refined"]
        if oc["refined"]:
            assert (c.original_start_ns, c.original_duration_ns) == (oc["orig"][0], oc["orig"][1] - oc["orig"][0])
        check_sample(c.sample, oc)
    assert bool(va.has_speech_profile) == (o["speech"] is not None)
    if o["speech"] is not None:
        assert (va.speech_profile.region.start_ns, va.speech_profile.region.end_ns) == o["speech"]["region"]
    assert bool(va.has_noise_profile) == (o["noise_profile"] is not None)
    if o["noise_profile"] is not None:
        p, op = va.noise_profile, o["noise_profile"]
        assert (p.start_ns, p.duration_ns) == (op["start"], op["duration"]) and same(p.measured_noise_floor, op["floor"])
        assert same(p.peak_level, op["peak"]) and same(p.crest_factor, op["crest"]) and same(p.entropy, op["entropy"])
        assert (va.noise_region.start_ns, va.noise_region.end_ns) == o["noise_region"]
        check_sample(va.room_tone_sample, o["room_tone"])
    assert same(va.voiced_low_percentile, o["voiced_low"]) and same(va.noise_high_percentile, o["noise_high"]) and same(va.gate_separation_db, o["separation"])
    assert same(va.floored_fraction, o["floored_fraction"]) and bool(va.voice_activated) == o["voice_activated"]
    # bands + AdaptConfig + BuildFilterSpec
    body = rnd.uniform(-45, -20)
    speech_bands = (body, body + rnd.uniform(-9, 3)) if rnd.random() < 0.8 else None
    noise_bands = [rnd.uniform(-95, -60) for _ in range(14)] + [rnd.choice([math.nan, -110.0, math.inf])] if rnd.random() < 0.8 else None
    if rnd.random() < 0.1 and noise_bands:
        noise_bands = [math.nan] * 7 + noise_bands[7:]
    A.apply_band_rms(va, (list(speech_bands), [1, 1]) if speech_bands else None, (noise_bands, [1] * 15) if noise_bands else None)
    cfg, _ = A.adapt_config(m, va)
    spec = A.build_filter_spec(cfg)
    ospec = AO.adapt_spec(meas, o, speech_bands, noise_bands)
    assert spec == ospec


@pytest.mark.parametrize("v", [12.0, -58.0, -52.37421875, -54.60033333333333, -79.99999999999999, -20.0, -61.3, 1e6, 123456.7, 1e-5, 0.000123])
def test_go_g_agrees(v):
    assert A.go_format_g(v) == AO.go_g(v)


@pytest.mark.parametrize("seed", range(0, 120, 7))
def test_detector_before_astats_then_assign_equals_one_call(seed):
    """The adaptive driver runs the detector while Pass 1's astats is still on the GPU (its whole-file values are the one product
    of Pass 1 the host needs last) and fills the two values that read it afterwards (jt_vad_assign_astats): byte for byte the
    VoiceActivity of a single jt_detect_voice_activity call on complete measurements (analyser.go:374-395, 515-531)."""
    import ctypes as C
    civ = to_c(random_stream(seed))
    rnd = random.Random(2000 + seed)
    full = dict(input_i=rnd.uniform(-45, -12), input_lra=rnd.uniform(2, 22), Dynamic_range=50.0, RMS_level=rnd.choice([0.0, rnd.uniform(-50, -15)]),
                Peak_level=rnd.uniform(-20, 0), Noise_floor=rnd.choice([-70.0, math.nan, -55.5]))
    m_full = A.new_measurements(**full)
    va_full, runs_full, cands_full = A.detect_voice_activity(m_full, civ)
    m_part = A.new_measurements(input_i=full["input_i"], input_lra=full["input_lra"])        # astats not collected yet
    va, runs, cands = A.detect_voice_activity(m_part, civ)
    for k in A.AS_NAMES:
        if k in full:
            m_part.astats[A.AS_NAMES.index(k)] = full[k]
    L = A._L()
    L.jt_vad_assign_astats.restype = None
    L.jt_vad_assign_astats.argtypes = [C.c_void_p, C.c_void_p]
    L.jt_vad_assign_astats(C.addressof(m_part), C.addressof(va))

    def raw(s):                   # NaNs compare equal as bytes
        return bytes(C.string_at(C.addressof(s), C.sizeof(s)))
    assert raw(va) == raw(va_full)
    assert [(r.start_ns, r.end_ns) for r in runs] == [(r.start_ns, r.end_ns) for r in runs_full] and len(cands) == len(cands_full)
