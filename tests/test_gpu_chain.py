"""GPU parity for the whole Pass 2 / Pass 3 / Pass 4 graphs and the four-pass driver
(ProcessAudio, processor.go:78-216) against the oracle chain, on seeded synthetic input."""
import math
import numpy as np
import pytest
import oracle_graph as OG
from jivetalking_b200 import gpudsp, synth

pytestmark = pytest.mark.gpu


def rms(a):
    return float(np.sqrt(np.mean(np.square(a.astype(np.float64))))) if len(a) else 0.0


def pcm_close_s16(g, e):
    assert g.dtype == np.int16 and e.dtype == np.int16 and len(g) == len(e)
    d = (g.astype(np.int32) - e.astype(np.int32)) / 32768.0
    assert rms(d) < 1e-4, rms(d)                 # north_star tolerance: 1e-4 RMS of full scale
    # adeclick's detector is a hard threshold on the AR residual: a last-bit difference upstream can flip
    # one sample's click flag and move that sample by a few hundred LSB; such samples must stay rare
    assert np.mean(np.abs(d) > 2.5 / 32768.0) < 2e-3, float(np.mean(np.abs(d) > 2.5 / 32768.0))
    assert np.max(np.abs(d)) < 2e-2


GOLDEN_ADAPTIVE = [  # adaptive_test.go:109-124 (Pass-2 golden spec strings), wrapped like BuildFilterSpec does
    "highpass=f=80:poles=2:width_type=q:width=0.707:normalize=1:a=tdii,lowpass=f=20500:poles=2:width_type=q:width=0.707:normalize=1,"
    "anlmdn=s=0.00001:p=0.0060:r=0.0058:m=11,afftdn=nr=12:nt=w:tn=0:nf=-58,"
    "agate=threshold=0.019953:ratio=2.0:attack=5.00:release=200:range=0.1995:knee=3.0:detection=rms:makeup=1.0,"
    "acompressor=threshold=0.031623:ratio=3.0:attack=10:release=200:makeup=1.00:knee=4.0:detection=rms:mix=1.00",
]


def wrap_pass2(core, deesser=""):
    return ("aformat=channel_layouts=mono," + core + deesser +
            ",astats=metadata=1:measure_perchannel=all,aspectralstats=win_size=2048:win_func=hann:measure=all,"
            "ebur128=metadata=1:peak=sample+true:dualmono=true:target=-16,"
            "aformat=sample_rates=44100:channel_layouts=mono:sample_fmts=s16,asetnsamples=n=4096")


@pytest.fixture(scope="module")
def speech():
    return synth.speech_like(40.0, 48000, seed=12345)


def test_pass2_default_spec(ctx, speech):
    spec = gpudsp.default_pass2_spec()
    got = ctx.run_graph(spec, speech, 48000)
    exp = OG.run_spec(spec, speech, 48000)
    pcm_close_s16(got["pcm"], exp["pcm"])
    assert len(got["pcm"]) % 4096 == 0
    OG.assert_meta_close(got["meta"], exp["meta"], spectral_rtol=5e-3, astats_atol=2e-3, roundoff_only_below_lufs=-100.0)   # behind the f32 stages: their round-off noise (~1e-5 of the signal) bounds agreement


def test_pass2_golden_adaptive_spec_with_deesser(ctx, speech):
    spec = wrap_pass2(GOLDEN_ADAPTIVE[0], ",deesser=i=0.35:m=0.50:f=0.80")
    got = ctx.run_graph(spec, speech, 48000)
    exp = OG.run_spec(spec, speech, 48000)
    pcm_close_s16(got["pcm"], exp["pcm"])
    OG.assert_meta_close(got["meta"], exp["meta"], spectral_rtol=5e-3, astats_atol=2e-3, roundoff_only_below_lufs=-100.0)   # behind the f32 stages: their round-off noise (~1e-5 of the signal) bounds agreement


def test_pass2_stereo_96k(ctx):
    x = synth.stereo_from_mono(synth.speech_like(8.0, 96000, seed=21))
    spec = gpudsp.default_pass2_spec()
    got = ctx.run_graph(spec, x, 96000, channels=2)
    exp = OG.run_spec(spec, x, 96000, channels=2)
    pcm_close_s16(got["pcm"], exp["pcm"])


def test_full_four_pass_chain(ctx, speech):
    pcm, res = ctx.process_audio(speech, 48000)
    # oracle: same orchestration in Python
    p2 = OG.run_spec(gpudsp.default_pass2_spec(), speech, 48000)
    last = [m for m in p2["meta"] if not math.isnan(m["I"])][-1]
    out_i = last["I"]
    out_tp = -120.0 if last["true_peak"] <= 0 else 20 * math.log10(last["true_peak"])
    spec3, plan = gpudsp.build_pass3_spec(out_i, out_tp)
    p3 = OG.run_spec(spec3, p2["pcm"], 44100, want_pcm=False)
    st = gpudsp.LoudnormStats()
    for k in ("input_i", "input_tp", "input_lra", "input_thresh"):
        setattr(st, k, p3["loudnorm"][k])
    spec4, eff, off = gpudsp.build_pass4_spec(plan, st)
    p4 = OG.run_spec(spec4, p2["pcm"], 44100)
    assert abs(res.filtered.input_i - out_i) < 0.011 and abs(res.filtered.input_tp - out_tp) < 0.2
    for k in ("input_i", "input_tp", "input_lra", "input_thresh"):
        assert abs(getattr(res.pass3, k) - p3["loudnorm"][k]) < 5e-3, k
    assert res.pass4.normalization_type == p4["loudnorm"]["normalization_type"] == 0
    # end to end only the contract tolerance applies (1e-4 RMS): Pass 4 re-reads the 16-bit Pass-2 output, and
    # adeclick's hard-threshold detector turns a single 1-LSB rounding tie there into a different click set
    d = (pcm.astype(np.int32) - p4["pcm"].astype(np.int32)) / 32768.0
    assert len(pcm) == len(p4["pcm"]) and rms(d) < 1e-4, rms(d)
    # ... while on IDENTICAL Pass-2 samples the Pass-3 / Pass-4 graphs agree tightly
    g3 = ctx.run_graph(spec3, p2["pcm"], 44100, want_pcm=False, want_meta=False)
    # (input_tp is the peak of the f32-internal 192 kHz resample: the 32 taps are summed in another order than the
    #  oracle's even/odd split -- as libswresample's own SIMD paths do -- so it carries one f32 ulp, ~1e-6 dB)
    for k in ("input_i", "input_tp", "input_lra", "input_thresh"):
        assert abs(getattr(g3["loudnorm"], k) - p3["loudnorm"][k]) < (1e-5 if k == "input_tp" else 1e-6), k
    g4 = ctx.run_graph(spec4, p2["pcm"], 44100)
    pcm_close_s16(g4["pcm"], p4["pcm"])
    OG.assert_meta_close(g4["meta"], p4["meta"], spectral_rtol=5e-3, roundoff_only_below_lufs=-100.0)
    fin = [m for m in p4["meta"] if not math.isnan(m["I"])][-1]
    assert abs(res.final.input_i - fin["I"]) < 0.011
    assert abs(res.final.input_lra - fin["LRA"]) < 0.05
    # the chain's purpose (filters.go:75-82): -16 LUFS +-0.5 LU, true peak under -1 dBTP
    assert abs(res.final.input_i - (-16.0)) <= 0.5
    assert res.final.input_tp <= -1.0 + 0.1
    assert res.n_out == len(pcm) and len(pcm) % 4096 == 0


def test_cancel_and_reuse(ctx):
    x = synth.speech_like(2.0, 48000, seed=1)
    a = ctx.run_graph(gpudsp.pass1_spec(), x, 48000, want_pcm=False)
    b = ctx.run_graph(gpudsp.pass1_spec(), x, 48000, want_pcm=False)
    assert [m.r128_M for m in a["meta"][5:10]] == [m.r128_M for m in b["meta"][5:10]]     # deterministic, ctx reusable
    assert ctx.launch_count() > 0


# a7: MeasureOutputRegions / measureOutputRegionFromReader (analyser_output.go:95-313) -- the region graph of
# analyser_output.go:18 on the s16 44.1 kHz Pass-2 / Pass-4 output, room-tone and speech regions
REGION_SPEC = ("atrim=start=%f:duration=%f,asetpts=PTS-STARTPTS,astats=metadata=1:measure_perchannel=0,"
               "aspectralstats=measure=all,ebur128=metadata=1:peak=sample+true")


@pytest.mark.parametrize("start,duration", [(0.0, 8.0), (12.25, 10.0), (31.7, 60.0), (39.99, 5.0)])
def test_output_region_measure(ctx, speech, start, duration):
    pcm = np.clip(np.round(synth.speech_like(40.0, 44100, seed=4242) * 32768.0), -32768, 32767).astype(np.int16)
    spec = REGION_SPEC % (start, duration)
    got = ctx.run_graph(spec, pcm, 44100, want_pcm=False)
    exp = OG.run_spec(spec, pcm, 44100, want_pcm=False)
    OG.assert_meta_close(got["meta"], exp["meta"])
    # what the reference keeps of it (analyser_output.go:134-169): last M / S / TP / SP, astats Overall, spectral means
    last = [m for m in got["meta"] if not math.isnan(m.r128_M)]
    elast = [m for m in exp["meta"] if not math.isnan(m["M"])]
    assert len(last) == len(elast)
    if last:
        assert abs(last[-1].r128_M - elast[-1]["M"]) < 0.0011 and abs(last[-1].r128_true_peak - elast[-1]["true_peak"]) < 0.0011
    ov = [m for m in got["meta"] if not math.isnan(m.astats_overall_RMS_level)]
    assert len(ov) == 1 and all(math.isnan(v) for v in ov[0].astats)         # measure_perchannel=0: Overall keys only


def test_concurrent_contexts(ctx):
    """SURVEY 8b threading contract: a jt_ctx is single-threaded, distinct contexts run concurrently from distinct threads
    (one per worker goroutine, cmd/jivetalking/pool.go:122-153 / CloneForWorker filters.go:368-373).  Four workers, own
    context and stream each, same file: every result equals the serial one bit for bit."""
    import threading
    xs = [synth.speech_like(20.0, 48000, seed=100 + i) for i in range(2)]
    serial = [ctx.process_audio(x, 48000) for x in xs]      # (fresh contexts below: their first call also uploads every table)
    results, errors = {}, []

    def worker(k):
        try:
            for rep in range(3):
                with gpudsp.Context(0) as c:             # a fresh context each time: first-call table uploads under load
                    pcm, res = c.process_audio(xs[(k + rep) % 2], 48000)
                    results[(k, rep)] = ((k + rep) % 2, pcm, res.final.input_i, res.final.input_tp, res.n_out)
        except Exception as e:          # noqa: BLE001
            errors.append(repr(e))
    threads = [threading.Thread(target=worker, args=(k,)) for k in range(6)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    assert len(results) == 18
    for (k, rep), (which, pcm, i, tp, n) in results.items():
        spcm, sres = serial[which]
        assert np.array_equal(pcm, spcm) and i == sres.final.input_i and tp == sres.final.input_tp and n == sres.n_out
