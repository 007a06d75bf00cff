"""GPU parity for loudnorm's DYNAMIC mode (libavfilter/af_loudnorm.c at its forced 192 kHz links) -- the path the reference's
Pass 4 lands on when the linear-mode preconditions fail and survives through its aresample=<rate> barrier
(internal/processor/normalise.go:683-693, 1294-1304) -- against the sequential oracle restatement (oracle/orc_loudnorm.c).
Covers: both start-up branches (above / below the measured threshold, the latter watching its own output), the peak
limiter (including its first-frame and final-frame special cases), streams that end inside a 100 ms frame, exactly 3 s,
shorter than 3 s (single-gain fall-back, printed as "linear"), the re-metered flush frame behind input_i, and the whole
Pass-4 graph of a stream whose Pass-3 LRA prints as 0.00 -- the reference's own TestProcessAudio fixture
(processor_test.go:360-376)."""
import math
import numpy as np
import pytest
import jt_oracle as O
import oracle_graph as OG
from jivetalking_b200 import gpudsp, synth

pytestmark = pytest.mark.gpu
FS = 192000


def _speechy(seconds, seed, level=0.25, quiet_start=0.0):
    t = np.arange(int(seconds * FS)) / FS
    rng = np.random.default_rng(seed)
    env = 0.05 + level * (np.sin(2 * np.pi * 0.3 * t) > 0) * (0.5 + 0.5 * np.sin(2 * np.pi * 4 * t))
    if quiet_start:
        env = np.where(t < quiet_start, 0.002, env)
    return env * np.sin(2 * np.pi * 220 * t) + 1e-3 * rng.standard_normal(len(t))


def _check(ctx, x, spec_opts, atol=2e-9, **kw):
    spec = "loudnorm=" + spec_opts
    got = ctx.run_graph(spec, x, FS, want_meta=False)
    y, st = O.loudnorm(x, FS, **kw)
    assert got["rate"] == 192000 and got["fmt"] == gpudsp.FMT_DBL and len(got["pcm"]) == len(x)
    ln = got["loudnorm"]
    assert ln.normalization_type == st["normalization_type"]
    d = np.abs(got["pcm"] - y)
    assert d.max() < atol, (d.max(), int(d.argmax()))
    for k in ("input_i", "input_tp", "input_lra", "input_thresh", "output_i", "output_tp", "output_lra", "output_thresh", "target_offset"):
        assert abs(getattr(ln, k) - st[k]) < 1e-6, (k, getattr(ln, k), st[k])
    return got, y, st


def test_dynamic_above_threshold_from_the_start(ctx):
    x = _speechy(20.37, 1)                       # ends inside a 100 ms frame
    _check(ctx, x, "I=-16:TP=-1.5:LRA=11:dual_mono=true", I=-16, TP=-1.5, LRA=11, dual_mono=True)


def test_dynamic_quiet_start_watches_its_own_output(ctx):
    x = _speechy(16.0, 2, quiet_start=5.0)
    got, y, st = _check(ctx, x, "I=-14:TP=-2:LRA=7:measured_I=-24.5:measured_TP=-3:measured_LRA=12:measured_thresh=-34.8:offset=1.5:dual_mono=true",
                        I=-14, TP=-2, LRA=7, mI=-24.5, mTP=-3, mLRA=12, mTh=-34.8, offset=1.5, dual_mono=True)
    assert st["normalization_type"] == 1


def test_dynamic_limiter_engaged(ctx):
    t = np.arange(12 * FS + 777) / FS
    rng = np.random.default_rng(3)
    env = np.where(t < 4, 0.003, 0.05 + 0.5 * (np.sin(2 * np.pi * 0.7 * t) > 0.3))
    x = env * np.sin(2 * np.pi * 220 * t) * (1 + 0.8 * np.sin(2 * np.pi * 3 * t)) + 1e-4 * rng.standard_normal(len(t))
    x[::50000] += 0.9                            # isolated over-ceiling spikes, including one in the first 10 ms and in the flush frame
    x[500] = 0.95
    got, y, st = _check(ctx, x, "I=-10:TP=-3:LRA=5:measured_I=-25:measured_TP=-1:measured_LRA=12:measured_thresh=-35:offset=3:dual_mono=true",
                        I=-10, TP=-3, LRA=5, mI=-25.0, mTP=-1, mLRA=12, mTh=-35, offset=3, dual_mono=True)
    ceiling = 10 ** (-3 / 20)
    assert np.abs(y).max() <= ceiling and np.sum(np.abs(y) >= ceiling * 0.999999) > 100      # the limiter and the clamp did work
    assert abs(st["output_tp"] - (-3.0)) < 1e-9


def test_dynamic_sustained_limiting(ctx):
    # loud throughout: the limiter never returns to idle, one region spans the stream
    t = np.arange(int(7.3 * FS)) / FS
    x = 0.7 * np.sin(2 * np.pi * 330 * t) * (0.6 + 0.4 * np.sin(2 * np.pi * 2.3 * t))
    _check(ctx, x, "I=-8:TP=-6:LRA=7:dual_mono=true", I=-8, TP=-6, LRA=7, dual_mono=True)


@pytest.mark.parametrize("seconds", [3.0, 3.05, 3.1])
def test_dynamic_three_seconds(ctx, seconds):
    x = _speechy(seconds, 4)
    got, y, st = _check(ctx, x, "I=-16:TP=-1.5:LRA=11:dual_mono=true", I=-16, TP=-1.5, LRA=11, dual_mono=True)
    assert st["normalization_type"] == 1


def test_dynamic_short_input_falls_back_to_one_gain(ctx):
    x = _speechy(2.0, 5)
    got, y, st = _check(ctx, x, "I=-16:TP=-1.5:LRA=11:dual_mono=true", atol=1e-12, I=-16, TP=-1.5, LRA=11, dual_mono=True)
    assert st["normalization_type"] == 0         # uninit() prints "linear" for this fall-back


def test_flush_frame_is_metered_twice(ctx):
    # Pass 3's measure-only call: input_* are those of the stream followed by its last 2.9 s again
    x = _speechy(20.0, 1)
    x[-2 * FS:] *= 3.0                           # a loud ending makes the double count visible
    got = ctx.run_graph("loudnorm=I=-16:TP=-1.5:LRA=11:dual_mono=true:print_format=json", x, FS, want_pcm=False, want_meta=False)
    twice = O.loudnorm_meter(np.concatenate([x, x[len(x) - 556800:]]), FS, True)
    once = O.loudnorm_meter(x, FS, True)
    assert abs(twice["I"] - once["I"]) > 0.05
    ln = got["loudnorm"]
    assert abs(ln.input_i - twice["I"]) < 1e-6 and abs(ln.input_thresh - twice["thresh"]) < 1e-6 and abs(ln.input_lra - twice["LRA"]) < 1e-6
    y, st = O.loudnorm(x, FS, I=-16, TP=-1.5, LRA=11, dual_mono=True)
    assert abs(st["input_i"] - twice["I"]) < 1e-9          # the full filter restatement agrees with the shortcut


def test_pass4_graph_with_dynamic_fallback_and_aresample_barrier(ctx):
    # a steady tone: Pass 3 prints measured_LRA=0.00, the sentinel that keeps loudnorm out of linear mode
    x = synth.reference_test_audio(6.0, 44100, 440.0, -18.0, -55.0)
    spec3, plan = gpudsp.build_pass3_spec(-21.0, -18.0)
    p3 = ctx.run_graph(spec3, x, 44100, want_pcm=False, want_meta=False)["loudnorm"]
    assert float("%.2f" % p3.input_lra) == 0.0
    spec4, eff, off = gpudsp.build_pass4_spec(plan, p3)
    assert "measured_LRA=0.00" in spec4 and ",aresample=44100," in spec4
    got = ctx.run_graph(spec4, x, 44100)
    exp = OG.run_spec(spec4, x, 44100)
    assert got["loudnorm"].normalization_type == exp["loudnorm"]["normalization_type"] == 1
    assert got["rate"] == 44100 and got["pcm"].dtype == np.int16 and len(got["pcm"]) == len(exp["pcm"]) and len(got["pcm"]) % 4096 == 0
    d = (got["pcm"].astype(np.int32) - exp["pcm"].astype(np.int32)) / 32768.0
    assert float(np.sqrt(np.mean(d * d))) < 1e-4 and np.max(np.abs(d)) < 2e-2
    for k in ("input_i", "output_i", "output_tp", "target_offset"):
        assert abs(getattr(got["loudnorm"], k) - exp["loudnorm"][k]) < 5e-3, k
    OG.assert_meta_close(got["meta"], exp["meta"], spectral_rtol=5e-3, roundoff_only_below_lufs=-100.0)


@pytest.mark.parametrize("gap", [(1.0, 0.3, 3.0), (0.0, 0.0, 3.0), (0.0, 0.0, 33 * 4096 / 44100.0)])
def test_process_audio_on_the_reference_fixture(ctx, gap):
    """TestProcessAudio (processor_test.go:360-376): 3 s, 44.1 kHz s16, 440 Hz at -18 dBFS + noise at -55 dBFS, 0.3 s gap at
    1 s, through the test's own minimal chain.  Whether Pass 4's loudnorm stays linear hangs on how Pass 3's LRA prints
    (0.1 LU histogram bins over three short-term blocks, the flush frame counted twice, 65 ms of asetnsamples padding
    inside the second one): 0.00 selects dynamic mode.  The reference ships the file either way, so must we."""
    from jivetalking_b200 import adapt as A
    x = synth.reference_test_audio(gap[2], 44100, 440.0, -18.0, -55.0, gap[0], gap[1])
    base = A.default_filter_config()             # newTestBaseConfig + downmix, analysis, resample, 95 Hz high-pass
    base.bandlimit_lowpass.enabled = base.noise_reduction.enabled = base.speech_gate.enabled = 0
    base.levelling_compressor.enabled = base.deesser.enabled = 0
    base.rumble_highpass.frequency = 95.0
    pcm, res, an = A.process_audio_adaptive(ctx, x, 44100, base=base)
    assert "highpass=f=95" in an.pass2_spec.decode()
    assert res.pass4.valid
    # af_loudnorm's init(): measured_LRA == 0 is one of the sentinels that keep it out of linear mode
    assert res.pass4.normalization_type == (1 if float("%.2f" % res.pass3.input_lra) == 0.0 else 0)
    assert len(pcm) % 4096 == 0 and res.n_out == len(pcm) and len(pcm) >= 3 * 44100
    # same orchestration over the oracle
    p2 = OG.run_spec(an.pass2_spec.decode(), x, 44100)
    last = [m for m in p2["meta"] if not math.isnan(m["I"])][-1]
    out_i, out_tp = last["I"], 20 * math.log10(last["true_peak"])
    spec3, plan = gpudsp.build_pass3_spec(out_i, out_tp)
    p3 = OG.run_spec(spec3, p2["pcm"], 44100, want_pcm=False)
    st = gpudsp.LoudnormStats()
    for k in ("input_i", "input_tp", "input_lra", "input_thresh"):
        setattr(st, k, p3["loudnorm"][k])
        assert abs(getattr(res.pass3, k) - p3["loudnorm"][k]) < 5e-3, k
    spec4, eff, off = gpudsp.build_pass4_spec(plan, st)
    p4 = OG.run_spec(spec4, p2["pcm"], 44100)
    assert p4["loudnorm"]["normalization_type"] == res.pass4.normalization_type
    d = (pcm.astype(np.int32) - p4["pcm"].astype(np.int32)) / 32768.0
    assert len(pcm) == len(p4["pcm"]) and float(np.sqrt(np.mean(d * d))) < 1e-4
    fin = [m for m in p4["meta"] if not math.isnan(m["I"])][-1]
    assert abs(res.final.input_i - fin["I"]) < 0.1 and abs(res.final.input_tp - 20 * math.log10(fin["true_peak"])) < 0.1
