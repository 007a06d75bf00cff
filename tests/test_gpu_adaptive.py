"""GPU parity for the adaptive path on the library side: jt_analyse_adaptive (AnalyseAudio + AdaptConfig, analyser.go:325-372,
adaptive.go:13-40) and jt_process_audio_adaptive (ProcessAudio, processor.go:78-216) against the CPU oracle, on a seeded
conversational synthetic (speech runs / room-tone pauses) that drives every adaptive branch: elected speech profile and
room-tone region, measured custom afftdn profile (nt=custom:bn), static nf, voiced-anchored gate, speech-RMS-anchored
compressor, de-esser ramp."""
import math

import numpy as np
import pytest

import jt_oracle as O
import oracle_graph as OG
from jivetalking_b200 import adapt as A
from jivetalking_b200 import gpudsp, shard, synth

pytestmark = pytest.mark.gpu
RATE = 48000


def oracle_intervals(ivs):
    out = []
    for d in ivs:
        iv = A.interval(d["ts_ns"], rms=d["rms"], peak=d["pk"], momentary=d["M"], short_term=d["S"], true_peak=d["tp"],
                        sample_peak=d["sp"], found=d["found"])
        for k, v in enumerate(d["spectral"]):
            iv.spectral[k] = v
        iv.frame_count = d["fc"]
        out.append(iv)
    return out


def oracle_band_rms(x, rate, start_ns, dur_ns, lo, hi):
    """measureSpeechBandRMS (analyser_bands.go:33-104) with the oracle's biquads and astats"""
    st, du = float("%f" % (start_ns / 1e9)), float("%f" % (dur_ns / 1e9))
    s0 = (round(st * 1e6) * rate + 500000) // 1000000
    n = (round(du * 1e6) * rate + 500000) // 1000000
    reg = x[s0:s0 + n]
    return [O.astats(O.biquad(O.biquad(reg, rate, "highpass", l), rate, "lowpass", h), rate)["RMS_level"] for l, h in zip(lo, hi)]


def spec_numbers(spec):
    """every filter option of a spec as (filter, key) -> float or str, bn split into its 15 values"""
    out = {}
    for i, f in enumerate(spec.split(",")):
        name, _, opts = f.partition("=")
        for kv in opts.split(":"):
            k, _, v = kv.partition("=")
            if k == "bn":
                for j, b in enumerate(v.split("|")):
                    out[(i, name, f"bn{j}")] = float(b)
                continue
            try:
                out[(i, name, k)] = float(v)
            except ValueError:
                out[(i, name, k)] = v
    return out


@pytest.fixture(scope="module")
def podcast():
    return synth.podcast_like(80.0, RATE, sibilance_db=-14.0)


@pytest.fixture(scope="module")
def analysis(ctx, podcast):
    return A.analyse_adaptive(ctx, podcast, RATE)


def test_analyse_adaptive_vs_oracle(ctx, podcast, analysis):
    an, iv = analysis
    va = an.voice_activity
    # the oracle's Pass 1 (restated collectAnalysisFrames) through the same detector
    meas, oiv = OG.pass1_analyse(podcast, RATE)
    m = A.new_measurements(input_i=meas["input_i"], input_lra=meas["input_lra"], input_tp=meas["input_tp"], **meas["astats"])
    ova, oruns, ocands = A.detect_voice_activity(m, oracle_intervals(oiv))
    assert len(iv) == len(oiv)
    # discrete outcomes are identical, continuous ones agree to the wire's rounding (%.3f momentary values averaged per interval)
    assert (va.n_speech_regions, va.n_candidates, va.gap_tolerance, va.voice_activated) == (ova.n_speech_regions, ova.n_candidates, ova.gap_tolerance, ova.voice_activated)
    assert va.has_speech_profile and va.has_noise_profile and ova.has_speech_profile and ova.has_noise_profile
    for r, o in ((va.speech_profile.region, ova.speech_profile.region), (va.noise_region, ova.noise_region)):
        assert (r.start_ns, r.end_ns, r.duration_ns) == (o.start_ns, o.end_ns, o.duration_ns)
    for k in ("split", "floor", "floor_prescan", "margin", "voiced_low_percentile", "noise_high_percentile", "gate_separation_db", "floored_fraction"):
        assert abs(getattr(va, k) - getattr(ova, k)) < 3e-3, (k, getattr(va, k), getattr(ova, k))
    assert abs(va.speech_profile.sample.rms_level - ova.speech_profile.sample.rms_level) < 1e-5
    assert abs(va.speech_profile.score - ova.speech_profile.score) < 1e-4
    assert abs(va.noise_profile.spectral[A.SP_NAMES.index("flatness")] - ova.noise_profile.spectral[A.SP_NAMES.index("flatness")]) < 2e-3
    # the 17 band graphs over the elected regions
    lo, hi = A.band_plan()
    sp = oracle_band_rms(podcast, RATE, ova.speech_profile.region.start_ns, ova.speech_profile.region.duration_ns, lo[:2], hi[:2])
    nz = oracle_band_rms(podcast, RATE, ova.noise_profile.start_ns, ova.noise_profile.duration_ns, lo[2:], hi[2:])
    assert abs(va.speech_profile.body_band_rms - sp[0]) < 2e-3 and abs(va.speech_profile.sib_band_rms - sp[1]) < 2e-3
    assert va.speech_profile.bands_measured and va.noise_profile.bands_measured and va.noise_profile.n_band_noise == 15
    for b in range(15):
        assert abs(va.noise_profile.band_noise[b] - nz[b]) < 2e-3, (b, va.noise_profile.band_noise[b], nz[b])
    # ... and the adapted configuration / spec string
    A.apply_band_rms(ova, (sp, [1, 1]), (nz, [1] * 15))
    ocfg, _ = A.adapt_config(m, ova)
    ospec = A.build_filter_spec(ocfg)
    spec = an.pass2_spec.decode()
    g, o = spec_numbers(spec), spec_numbers(ospec)
    assert g.keys() == o.keys()
    for k in g:
        if isinstance(g[k], str):
            assert g[k] == o[k], k
        else:
            tol = 0.1001 if k[2].startswith("bn") else 3e-3 * max(1.0, abs(o[k]))      # bn is printed with one decimal
            assert abs(g[k] - o[k]) <= tol, (k, g[k], o[k])
    # every adaptive branch is live on this input
    assert "nt=custom:bn=" in spec and ":tn=0:nf=" in spec and "deesser=i=" in spec
    assert an.diagnostics.speech_gate_clamp_reason == b"none" and an.diagnostics.afftdn_noise_type == b"custom"


def test_analyse_adaptive_is_the_composition_of_its_parts(ctx, podcast, analysis):
    """jt_analyse_adaptive == jt_analyse -> jt_detect_voice_activity -> jt_band_rms x2 -> jt_apply_band_rms -> jt_adapt_config
    -> jt_build_filter_spec, value for value (what a Go caller keeping its own orchestration would get)."""
    an, iv = analysis
    m, iv2 = ctx.analyse(podcast, RATE)
    assert len(iv) == len(iv2) and all(bytes(a) == bytes(b) for a, b in zip(iv, iv2))
    va, runs, cands = A.detect_voice_activity(m, iv2)
    lo, hi = A.band_plan()
    r = va.speech_profile.region
    sp, spf = ctx.band_rms(podcast, RATE, float("%f" % (r.start_ns / 1e9)), float("%f" % (r.duration_ns / 1e9)), lo[:2], hi[:2])
    nz, nzf = ctx.band_rms(podcast, RATE, float("%f" % (va.noise_profile.start_ns / 1e9)), float("%f" % (va.noise_profile.duration_ns / 1e9)), lo[2:], hi[2:])
    A.apply_band_rms(va, (list(sp), list(spf)), (list(nz), list(nzf)))
    cfg, diag = A.adapt_config(m, va)
    assert A.build_filter_spec(cfg) == an.pass2_spec.decode()
    assert bytes(va) == bytes(an.voice_activity) and bytes(cfg) == bytes(an.config) and bytes(diag) == bytes(an.diagnostics)


def test_process_audio_adaptive(ctx, podcast, analysis):
    an, _ = analysis
    spec = an.pass2_spec.decode()
    pcm, res, an2 = A.process_audio_adaptive(ctx, podcast, RATE)
    assert an2.pass2_spec == an.pass2_spec and bytes(an2.voice_activity) == bytes(an.voice_activity)
    # same result as the caller-supplied-spec entry with that spec (the reference's Go side doing AdaptConfig itself)
    pcm_b, res_b = ctx.process_audio(podcast, RATE, pass2_spec=spec)
    assert np.array_equal(pcm, pcm_b) and res.n_out == res_b.n_out
    assert bytes(res.input) == bytes(res_b.input) and bytes(res.final) == bytes(res_b.final) and bytes(res.pass3) == bytes(res_b.pass3)
    # Pass 2 under the adapted spec against the oracle chain
    got = ctx.run_graph(spec, podcast, RATE)
    exp = OG.run_spec(spec, podcast, RATE)
    d = (got["pcm"].astype(np.int32) - exp["pcm"].astype(np.int32)) / 32768.0
    assert len(got["pcm"]) == len(exp["pcm"]) and float(np.sqrt(np.mean(d * d))) < 1e-4           # north_star: 1e-4 RMS of full scale
    last = [m for m in exp["meta"] if not math.isnan(m["I"])][-1]
    assert abs(res.filtered.input_i - last["I"]) < 0.011                                           # %.3f wire + f32 stage round-off
    assert abs(res.filtered.input_lra - last["LRA"]) < 0.05
    # the chain's purpose (filters.go:75-82): -16 LUFS +-0.5 LU, true peak at or under -1 dBTP
    assert abs(res.final.input_i - (-16.0)) <= 0.5 and res.final.input_tp <= -1.0 + 0.1
    assert len(pcm) % 4096 == 0 and res.n_out == len(pcm)
    # a7: the elected regions re-measured on the Pass-2 and Pass-4 outputs (MeasureOutputRegions, processor.go:150-160)
    va = an2.voice_activity
    for regions, audio in ((an2.filtered_regions, got["pcm"]), (an2.final_regions, pcm)):
        assert regions.has_room_tone and regions.has_speech
        for sample, (st, du) in ((regions.room_tone, (va.noise_profile.start_ns, va.noise_profile.duration_ns)),
                                 (regions.speech, (va.speech_profile.region.start_ns, va.speech_profile.region.duration_ns))):
            o, frames = OG.region_sample(audio, 44100, st, du)
            assert frames > 0
            for k in ("rms_level", "peak_level", "crest_factor"):
                assert abs(getattr(sample, k) - o[k]) < 2e-6, (k, getattr(sample, k), o[k])
            for k in ("momentary_lufs", "short_term_lufs"):
                assert abs(getattr(sample, k) - o[k]) < 0.0011, (k, getattr(sample, k), o[k])
            for k in ("true_peak", "sample_peak"):
                assert abs(getattr(sample, k) - o[k]) < 0.2, k           # linear peaks on a %.3f wire, in dB
            for k in range(gpudsp.SP_COUNT):
                assert abs(sample.spectral[k] - o["spectral"][k]) <= 2e-3 * abs(o["spectral"][k]) + 1e-9, (k, sample.spectral[k], o["spectral"][k])
            # the ABI entry on host audio gives the same sample
            s2, f2 = A.measure_output_region(ctx, audio, 44100, st, du)
            assert bytes(s2) == bytes(sample) and f2 == frames
    # noise reduction did its job in the room-tone region, speech came up to target
    assert an2.final_regions.speech.rms_level > an2.voice_activity.speech_profile.sample.rms_level


def test_voice_activated_capture_drops_afftdn(ctx):
    """A platform-gated capture (digital silence between phrases): floored fraction >= 0.20 => afftdn leaves the chain
    (adaptive.go:140-145) and the pre-scan seed falls back when no room tone is measurable."""
    x = synth.podcast_like(60.0, RATE, seed=7)
    gate = np.ones(len(x), dtype=np.float32)
    # hard-mute everything under the phrase level, as a conferencing platform's gate does
    env = np.convolve(np.abs(x), np.ones(2400) / 2400.0, mode="same")
    gate[env < 4e-3] = 0.0
    an, iv = A.analyse_adaptive(ctx, x * gate, RATE)
    assert an.voice_activity.voice_activated and an.voice_activity.floored_fraction >= 0.20
    assert "afftdn" not in an.pass2_spec.decode() and "anlmdn=" in an.pass2_spec.decode()
    assert an.diagnostics.afftdn_disable_reason == b"voice_activated"
    # the adapted spec runs through Pass 2
    got = ctx.run_graph(an.pass2_spec.decode(), x * gate, RATE, want_meta=False)
    assert got["rate"] == 44100 and len(got["pcm"]) % 4096 == 0 and len(got["pcm"]) >= 60 * 44100


def test_adaptive_chain_sharded_equals_single_gpu(ctx, podcast, analysis):
    """configs[3] with the adaptive spec: Pass 1 (chunk + merge) -> detector / AdaptConfig on every rank -> Passes 2-4 in
    chunks == jt_process_audio_adaptive on the whole stream."""
    an, _ = analysis
    pcm1, res1, _ = A.process_audio_adaptive(ctx, podcast, RATE)
    pcm, r = shard.process_stream_sharded_adaptive(ctx, shard.LocalComm(4), podcast, RATE)
    assert r["specs"][0] == an.pass2_spec.decode()
    assert bytes(r["voice_activity"]) == bytes(an.voice_activity)
    assert len(pcm) == len(pcm1) == res1.n_out
    d = (pcm.astype(np.int32) - pcm1.astype(np.int32)) / 32768.0
    assert float(np.sqrt(np.mean(d * d))) < 1e-4
    # the true peak reaches the host as a linear value printed with three decimals ("%.3f" of ~0.56: one step is 0.0154 dB), and
    # a chunk boundary moves the f32 biquads' rounding noise, so the last decimal may differ by one
    assert abs(r["final"].input_i - res1.final.input_i) <= 0.01 and abs(r["final"].input_tp - res1.final.input_tp) <= 0.02
    assert abs(r["final"].input_lra - res1.final.input_lra) <= 0.05
