"""GPU parity: Pass-1 analysis graph (astats + aspectralstats + ebur128 incl. true peak) through
the C ABI vs the CPU oracle, on the reference's own synthetic recipes
(analyser_test.go:135-148, benchmark_test.go:96-109) and on edge cases."""
import math
import numpy as np
import pytest
import oracle_graph as OG
from jivetalking_b200 import gpudsp, synth

pytestmark = pytest.mark.gpu


def run_pass1(ctx, x, rate, channels=1):
    got = ctx.run_graph(gpudsp.pass1_spec(), x, rate, channels=channels, want_pcm=False)
    exp = OG.pass1_meta(x, rate, channels)
    OG.assert_meta_close(got["meta"], exp)
    return got, exp


def test_reference_recipe_s16_48k(ctx):
    # C1 / TestAnalyseAudio: 440 Hz -23 dBFS + -60 dBFS noise, 0.5 s gap at 2 s
    x = synth.reference_test_audio(10.0, 48000, 440.0, -23.0, -60.0, 2.0, 0.5)
    got, _ = run_pass1(ctx, x, 48000)
    last = [m for m in got["meta"] if not math.isnan(m.r128_I)][-1]
    # the reference's own sanity ranges (analyser_test.go:183-205)
    assert -30 < last.r128_I < -20
    assert 0 <= last.r128_LRA <= 15
    assert -30 < 20 * math.log10(last.r128_true_peak) < 0


def test_benchmark_recipe_s16_44k(ctx):
    x = synth.reference_test_audio(40.0, 44100, 180.0, -18.0, -58.0, 30.0, 1.5)
    run_pass1(ctx, x, 44100)


def test_speech_like_f32_48k(ctx):
    x = synth.speech_like(65.0, 48000, seed=12345)
    got, _ = run_pass1(ctx, x, 48000)
    last = [m for m in got["meta"] if not math.isnan(m.r128_I)][-1]
    assert last.r128_LRA > 1.0


def test_stereo_f32_96k(ctx):
    x = synth.stereo_from_mono(synth.speech_like(12.0, 96000, seed=5))
    run_pass1(ctx, x, 96000, channels=2)


@pytest.mark.parametrize("n", [1, 100, 4095, 4800, 4801, 19199, 19200, 48000 * 3 + 17])
def test_short_and_ragged_lengths(ctx, n):
    x = synth.speech_like(4.0, 48000, seed=3)[:n]
    run_pass1(ctx, x, 48000)


def test_digital_silence(ctx):
    x = np.zeros(48000 * 2, dtype=np.float32)
    got = ctx.run_graph(gpudsp.pass1_spec(), x, 48000, want_pcm=False)
    ticks = [m for m in got["meta"] if not math.isnan(m.r128_M)]
    assert all(m.r128_M < -115 for m in ticks)        # isFlooredLevel relies on this (analyser_vad.go:72-74)
    assert ticks[-1].r128_sample_peak == 0.0


def test_analyse_intervals(ctx):
    """jt_analyse: interval accumulation of collectAnalysisFrames (analyser.go:571-638) against the oracle's restatement
    of the Go accumulation over the oracle's sink-frame records (oracle_graph.pass1_analyse)."""
    x = synth.reference_test_audio(5.0, 48000, 440.0, -23.0, -60.0, 2.0, 0.5)
    m, iv = ctx.analyse(x, 48000)
    meas, intervals = OG.pass1_analyse(x, 48000)
    assert len(iv) == len(intervals)
    for a, b in zip(iv, intervals):
        assert abs(a.timestamp_s - b["ts_ns"] * 1e-9) < 1e-9 and a.frame_count == b["fc"]
        assert abs(a.rms_level - b["rms"]) < 1e-6 and abs(a.peak_level - b["pk"]) < 1e-6
        assert abs(a.momentary_lufs - b["M"]) < 2e-3 and abs(a.short_term_lufs - b["S"]) < 2e-3
        assert abs(a.true_peak - b["tp"]) < 0.2 and abs(a.sample_peak - b["sp"]) < 0.2
        assert bool(a.spectral_found) == b["found"]
        for k in range(gpudsp.SP_COUNT):
            assert abs(a.spectral[k] - b["spectral"][k]) <= 2e-3 * abs(b["spectral"][k]) + 1e-9, (k, a.spectral[k], b["spectral"][k])
    assert abs(m.duration_s - 5.0) < 1e-9 and m.sink_frames == meas["sink_frames"]
    assert abs(m.input_i - meas["input_i"]) < 2e-3 and abs(m.input_lra - meas["input_lra"]) < 2e-3
