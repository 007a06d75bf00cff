"""GPU parity for 32-bit integer input (JT_FMT_S32 -- what 24-bit FLAC / WAV masters decode to; the reference handles it in
its raw-frame path, analyser_metrics.go:285-340, and FFmpeg's filters in theirs: astats and af_biquads have native s32 paths,
swr converts s32 to flt for rematrix / resampling) against the oracle."""
import math
import numpy as np
import pytest
import jt_oracle as O
import oracle_graph as OG
from jivetalking_b200 import gpudsp, synth
from jivetalking_b200 import adapt as A

pytestmark = pytest.mark.gpu


def to_s24_in_s32(x):
    """24-bit samples as a FLAC / pcm_s24le decoder delivers them: left-justified in int32"""
    q = np.clip(np.round(np.asarray(x, dtype=np.float64) * (1 << 23)), -(1 << 23), (1 << 23) - 1).astype(np.int32)
    return (q << 8).astype(np.int32)


@pytest.fixture(scope="module")
def speech24():
    return to_s24_in_s32(synth.speech_like(30.0, 48000, seed=31))


def test_pass1_s32_mono(ctx, speech24):
    got = ctx.run_graph(gpudsp.pass1_spec(), speech24, 48000, want_pcm=False)
    exp = OG.pass1_meta(speech24, 48000)
    OG.assert_meta_close(got["meta"], exp)
    last = [m for m in got["meta"] if not math.isnan(m.astats[0])][-1]
    assert 16.0 <= last.astats[gpudsp.AS_NAMES.index("Bit_depth")] <= 24.0     # astats counts the bits in use: never the low 8 of a 24-bit master


def test_analyse_s32_stereo_intervals(ctx):
    x = synth.stereo_from_mono(synth.speech_like(12.0, 48000, seed=32))
    s = to_s24_in_s32(x)
    m, iv = ctx.analyse(s, 48000, channels=2)
    em, eiv = OG.pass1_analyse(s, 48000, channels=2)
    assert len(iv) == len(eiv)
    assert abs(m.input_i - em["input_i"]) < 0.0011 and abs(m.input_tp - em["input_tp"]) < 0.01
    for g, e in zip(iv, eiv):
        assert abs(g.rms_level - e["rms"]) < 1e-6 and abs(g.peak_level - e["pk"]) < 1e-6           # a2: /2^31, all channels pooled
        assert abs(g.momentary_lufs - e["M"]) < 0.0011


def test_biquads_and_band_rms_s32(ctx, speech24):
    for spec in ("highpass=f=80:poles=2:width_type=q:width=0.707:normalize=1:a=tdii", "lowpass=f=3000:p=2"):
        got = ctx.run_graph(spec, speech24, 48000, want_meta=False)
        exp = OG.run_spec(spec, speech24, 48000)
        assert got["pcm"].dtype == np.int32
        d = np.abs(got["pcm"].astype(np.int64) - exp["pcm"].astype(np.int64))
        assert d.max() <= 1, d.max()          # f64 state; a lane started from zero state re-merges below one 32-bit LSB
    lo, hi = A.band_plan()
    got, found = ctx.band_rms(speech24, 48000, 3.25, 9.5, lo, hi)
    s0, n = round(3.25 * 48000), round(9.5 * 48000)
    reg = speech24[s0:s0 + n]
    for b in range(17):
        y = O.biquad(O.biquad(reg, 48000, "highpass", lo[b]), 48000, "lowpass", hi[b])
        exp = O.astats(y, 48000)["RMS_level"]
        assert found[b] == 1 and abs(got[b] - exp) < 1e-5, (b, got[b], exp)


def test_pass2_s32(ctx, speech24):
    spec = gpudsp.default_pass2_spec()
    got = ctx.run_graph(spec, speech24, 48000)
    exp = OG.run_spec(spec, speech24, 48000)
    d = (got["pcm"].astype(np.int32) - exp["pcm"].astype(np.int32)) / 32768.0
    assert len(got["pcm"]) == len(exp["pcm"]) and float(np.sqrt(np.mean(d * d))) < 1e-4
    OG.assert_meta_close(got["meta"], exp["meta"], spectral_rtol=5e-3, astats_atol=2e-3, roundoff_only_below_lufs=-100.0)


def test_process_audio_24_bit_matches_float_input_closely(ctx):
    """a 24-bit master and the same material as f32 give the same result to within the quantisation of the input"""
    x = synth.speech_like(30.0, 48000, seed=33)
    pcm24, res24 = ctx.process_audio(to_s24_in_s32(x), 48000)
    pcmf, resf = ctx.process_audio(x, 48000)
    assert len(pcm24) == len(pcmf)
    d = (pcm24.astype(np.int32) - pcmf.astype(np.int32)) / 32768.0
    assert float(np.sqrt(np.mean(d * d))) < 1e-4
    assert abs(res24.final.input_i - resf.final.input_i) < 0.05 and abs(res24.input.input_i - resf.input.input_i) < 0.01
