"""GPU parity: polyphase resampler / format conversion kernels vs the oracle and vs the real
libswresample golden vectors."""
import os
import numpy as np
import pytest
import jt_oracle as O
from jivetalking_b200 import gpudsp

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(os.path.dirname(__file__), "golden", "swr_golden.npz"))


@pytest.mark.parametrize("rates", [(48000, 192000), (48000, 44100), (44100, 192000), (96000, 44100)])
def test_aresample_f64_vs_real_swr_golden(ctx, rates):
    x = G["in_noise"]
    ref = G[f"dbl_noise_{rates[0]}_{rates[1]}_1"]
    got = ctx.run_graph(f"aresample={rates[1]}", x, rates[0], want_meta=False)
    assert got["rate"] == rates[1] and got["fmt"] == gpudsp.FMT_DBL
    assert len(got["pcm"]) == len(ref)
    assert np.max(np.abs(got["pcm"] - ref)) < 1e-14


@pytest.mark.parametrize("rates", [(22050, 192000), (11025, 192000), (47999, 44100)])
def test_aresample_inexact_ratio_vs_real_swr_golden(ctx, rates):
    """swr's linear-interpolated path (k_swr_linear) against outputs of the real library"""
    x = G["in_noise"][:2000]
    ref = G[f"dbl_noise_{rates[0]}_{rates[1]}_1"]
    got = ctx.run_graph(f"aresample={rates[1]}", x, rates[0], want_meta=False)
    assert got["rate"] == rates[1] and len(got["pcm"]) == len(ref)
    assert np.max(np.abs(got["pcm"] - ref)) < 1e-14
    if rates[0] == 22050:
        gf = ctx.run_graph("aresample=192000", x.astype(np.float32), 22050, want_meta=False)
        assert gf["fmt"] == gpudsp.FMT_FLT and len(gf["pcm"]) == len(G["flt_22050_192000"])
        assert np.max(np.abs(gf["pcm"] - G["flt_22050_192000"])) < 5e-7


@pytest.mark.parametrize("rate", [22050, 44110])
def test_true_peak_of_a_22k_source(ctx, rate):
    """ebur128 peak=true on a source whose ratio to 192 kHz is inexact (reduced phase count 1280 / 19200 > 1024): per-frame and
    whole-stream true peaks vs the oracle.  (11 025 Hz is not a multiple of 10: the 100 ms tick of the meter is not whole there.)"""
    import oracle_graph as OG
    from jivetalking_b200 import synth
    x = synth.speech_like(7.3, rate, seed=5)
    spec = "ebur128=metadata=1:peak=sample+true:dualmono=true:target=-16"
    got = ctx.run_graph(spec, x, rate, want_pcm=False)
    exp = OG.run_spec(spec, x, rate)
    OG.assert_meta_close(got["meta"], exp["meta"])
    assert got["meta"][-1].r128_true_peak > 0.01


def test_output_stage_48k_to_s16_44k(ctx):
    rng = np.random.default_rng(11)
    x = (rng.standard_normal(100003) * 0.25)
    got = ctx.run_graph("aformat=sample_rates=44100:channel_layouts=mono:sample_fmts=s16,asetnsamples=n=4096", x, 48000)
    y = O.swr_resample(x, 48000, 44100, flush=True)
    exp = np.zeros(len(y), dtype=np.int16)
    O.lib().orc_conv_f64_to_s16(O._ptr(y), len(y), O._ptr(exp))
    pad = (-len(exp)) % 4096
    exp = np.concatenate([exp, np.zeros(pad, dtype=np.int16)])
    assert got["fmt"] == gpudsp.FMT_S16 and got["rate"] == 44100
    assert len(got["pcm"]) == len(exp)
    d = np.abs(got["pcm"].astype(int) - exp.astype(int))
    assert d.max() <= 1 and np.count_nonzero(d) < 5          # rounding ties only
    assert all(m.nb_samples == 4096 for m in got["meta"])          # asetnsamples pads the last frame


def test_s16_to_192k_uses_f32_internal(ctx):
    # loudnorm dynamic mode on s16 input: swr resamples in FLTP (swresample.c int_sample_fmt rule)
    s16 = G["in_s16"]
    got = ctx.run_graph("loudnorm=I=-16.0:TP=-1.0:LRA=20.0:dual_mono=true:print_format=json", s16, 44100, want_pcm=False, want_meta=False)
    # (a stream shorter than 3 s takes af_loudnorm's single-gain fall-back, which uninit() prints as "linear")
    assert got["loudnorm"].valid == 1 and got["loudnorm"].normalization_type == 0
    assert got["rate"] == 192000


def test_stereo_downmix_matches_real_swr(ctx):
    st = G["in_stereo_f32"]
    got = ctx.run_graph("aformat=channel_layouts=mono", st, 48000, channels=2, want_meta=False)
    assert np.max(np.abs(got["pcm"] - G["stereo_f32_to_mono"])) < 1.2e-7
