"""CPU: the oracle's own pieces are consistent with each other and with published answers
(ITU-R BS.1770 / EBU Tech 3341 known answers, scipy, torchaudio), and with the loose sanity
ranges the reference's tests hold (analyser_test.go:183-205)."""
import math
import numpy as np
import pytest
import jt_oracle as O
import oracle_graph as OG
from jivetalking_b200 import synth, gpudsp


def test_run_spec_matches_pass1_composition():
    x = synth.reference_test_audio(3.0, 48000, 440.0, -23.0, -60.0, 2.0, 0.5)
    a = OG.pass1_meta(x, 48000)
    b = OG.run_spec(gpudsp.pass1_spec(), x, 48000, want_pcm=False)["meta"]
    assert len(a) == len(b)
    for ra, rb in zip(a, b):
        for k in ("first_sample", "nb_samples", "ready"):
            assert ra[k] == rb[k]
        for k in ("M", "S", "I", "LRA", "true_peak", "sample_peak"):
            assert (math.isnan(ra[k]) and math.isnan(rb[k])) or ra[k] == rb[k]
        assert ra["spectral"] == rb["spectral"]
        assert (ra["astats"] is None) == (rb["astats"] is None)
        if ra["astats"]:
            assert ra["astats"] == rb["astats"]


def test_reference_sanity_ranges():
    # TestAnalyseAudio: -23 dBFS 440 Hz + -60 dBFS noise, 5 s, 0.5 s gap at 2 s @44.1k
    x = synth.reference_test_audio(5.0, 44100, 440.0, -23.0, -60.0, 2.0, 0.5)
    r = O.ebur128(x.astype(np.float64) / 32768.0, 44100)
    assert -30 < r["I"] < -20
    assert -30 < 20 * math.log10(r["true_peak"]) < 0
    assert 0 <= r["LRA"] <= 15
    # TestMeasureOutputRoomToneRegion: room tone RMS < -40, peak < -30 (analyser_test.go:1071-1078)
    gap = x[int(2.05 * 44100):int(2.45 * 44100)]
    a = O.astats(gap, 44100)
    assert a["RMS_level"] < -40 and a["Peak_level"] < -30


def test_bs1770_known_answers():
    fs = 48000
    t = np.arange(fs * 20) / fs
    # EBU Tech 3341 case 1 analogue: 1 kHz sine at -23 dBFS, dual-mono == stereo -> -23.0 LUFS
    x = 10 ** (-23 / 20) * np.sin(2 * np.pi * 1000 * t)
    r = O.ebur128(x, fs, dualmono=True)
    assert abs(r["I"] - (-23.0)) < 0.1 and r["LRA"] < 0.1
    assert abs(r["M"][-1] - (-23.0)) < 0.1 and abs(r["S"][-1] - (-23.0)) < 0.1
    m = O.loudnorm_meter(x, fs, dual_mono=True)
    # loudnorm's meter (libavfilter/ebur128.c) keeps its gating blocks in 0.1 LU histogram bins: I is a bin centre
    assert abs(m["I"] - (-23.0)) < 0.1 and abs(m["I"] - r["I"]) < 0.06
    # without dual-mono a mono signal reads 3.01 LU lower
    assert abs(O.ebur128(x, fs, dualmono=False)["I"] - (r["I"] - 3.0103)) < 0.011
    # true peak of a sine sampled off-peak: fs/4 with 45 degree phase -> sample peak -3.01 dB, true peak ~0 dB
    # (faded in: swr mirrors the stream start, an abrupt onset would ring above the steady-state peak)
    y = np.sin(2 * np.pi * (fs / 4) * t + np.pi / 4) * np.minimum(1.0, t / 0.5)
    r = O.ebur128(0.5 * y, fs)
    assert abs(20 * math.log10(r["sample_peak"]) - (-6.02 - 3.01)) < 0.02
    assert abs(20 * math.log10(r["true_peak"]) - (-6.02)) < 0.1


def test_loudness_vs_torchaudio():
    torch = pytest.importorskip("torch")
    ta = pytest.importorskip("torchaudio")
    x = synth.speech_like(30.0, 48000, seed=2)
    ref = float(ta.functional.loudness(torch.tensor(x)[None], 48000)) + 3.0103
    r = O.ebur128(x.astype(np.float64), 48000)
    assert abs(r["I"] - ref) < 0.15
    assert abs(O.loudnorm_meter(x.astype(np.float64), 48000)["I"] - ref) < 0.15


def test_biquads_vs_scipy():
    ss = pytest.importorskip("scipy.signal")
    rng = np.random.default_rng(0)
    x = rng.standard_normal(20000) * 0.1
    for kind, f in (("highpass", 80.0), ("lowpass", 20500.0), ("highpass", 1000.0), ("lowpass", 3000.0)):
        c = np.zeros(5)
        O._proto("orc_biquad_design", None, [O.C.c_int, O._D, O._D, O.C.c_int, O.C.c_int, O._P])(int(kind == "highpass"), f, 0.707, 48000, 1, O._ptr(c))
        ref = ss.lfilter(c[:3], [1.0, c[3], c[4]], x)
        for tdii in (False, True):
            y = O.biquad(x, 48000, kind, f, normalize=True, tdii=tdii)
            assert np.max(np.abs(y - ref)) < 1e-12
        # RBJ q=0.707 high/low-pass is (to 3 digits of Q) the 2nd-order Butterworth
        b, a = ss.butter(2, f / 24000, "high" if kind == "highpass" else "low")
        assert np.max(np.abs(ss.lfilter(b, a, x) - ref)) < 2e-3


def test_spectral_stats_formulas():
    fs = 48000
    t = np.arange(fs) / fs
    x = (0.25 * np.sin(2 * np.pi * 3000 * t)).astype(np.float32)
    rows = O.aspectralstats(x, fs)
    mid = rows[10]
    names = O.SPEC_NAMES
    assert abs(mid[names.index("centroid")] - 3000) < 30          # a pure tone's centroid is its frequency
    assert abs(mid[names.index("rolloff")] - 3000) < 30
    assert mid[names.index("flatness")] < 0.01 and mid[names.index("crest")] > 100
    rng = np.random.default_rng(1)
    w = (rng.standard_normal(fs) * 0.1).astype(np.float32)
    rw = O.aspectralstats(w, fs)[10]
    assert rw[names.index("flatness")] > 0.5 and 9000 < rw[names.index("centroid")] < 15000


def test_dynamics_static_curves():
    fs = 48000
    x = np.full(fs, 0.5)
    # compressor: threshold -18 dB (0.125893), ratio 3, steady 0.5 input -> gain = (thr/x)^(1-1/3)
    y = O.acompressor(x, fs, 0.125893, 3.0, 10, 200, 1.0, 4.0)
    assert abs(y[-1] / 0.5 - (0.125893 / 0.5) ** (2 / 3)) < 1e-6
    # gate: below threshold by a lot -> clamps at range
    q = np.full(fs, 1e-4)
    g = O.agate(q, fs, 0.01, 2.0, 5, 200, 0.1995, 3.0)
    assert abs(g[-1] / 1e-4 - 0.1995) < 1e-9
    # above the knee the gate is transparent
    assert abs(O.agate(x, fs, 0.01, 2.0, 5, 200, 0.1995, 3.0)[-1] - 0.5) < 1e-12
    # limiter holds the ceiling and passes quiet material untouched
    s = synth.speech_like(3.0, 44100).astype(np.float64) * 4
    l = O.alimiter(s, 44100, 0.3, 5, 100, level=False, asc=True, asc_level=0.8)
    assert np.max(np.abs(l)) <= 0.3 + 1e-12
    quiet = s * 0.01
    assert np.array_equal(O.alimiter(quiet, 44100, 0.3, 5, 100, level=False, asc=True, asc_level=0.8), quiet)


def test_delays_and_lengths():
    fs = 48000
    x = synth.speech_like(2.0, fs)
    y = O.anlmdn(x, fs, 0.00001, 0.006, 0.002, 3)
    assert len(y) == len(x) and np.all(y[:384] == 0)               # K+S = 288+96 samples of latency, no flush
    z = O.afftdn(x, fs, 12, -50)
    assert len(z) == len(x) and np.max(np.abs(z[:600])) < 1e-6     # window - hop = 1200 samples of latency
    d, _ = O.adeclick(x.astype(np.float64), fs, 55, 50, 2, 1.7, 2, True)
    assert len(d) == len(x)
