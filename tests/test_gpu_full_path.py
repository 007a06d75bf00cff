"""The real path at BASELINE size against the oracle: jt_process_audio_adaptive (ProcessAudio, processor.go:78-216 -- Pass 1 ->
detector -> 17 band graphs -> AdaptConfig -> Pass 2 -> region re-measures -> Pass 3 -> Pass 4 -> region re-measures) on 10 minutes
of the C2 recipe (conversational synthetic, BASELINE.json configs[1]) versus the same orchestration over the oracle
(oracle/chain_oracle.py).  The kernels cut 10 minutes into thousands of lanes / hops / windows, the oracle walks it
sequentially.  JT_FULL_PATH_MINUTES=60 runs the whole hour (about a quarter of an hour of oracle CPU time)."""
import math
import os

import numpy as np
import pytest

import chain_oracle as CO
from jivetalking_b200 import adapt as A
from jivetalking_b200 import gpudsp, synth

pytestmark = pytest.mark.gpu
MINUTES = int(os.environ.get("JT_FULL_PATH_MINUTES", "10"))


def make_input(minutes):
    return np.concatenate([synth.podcast_like(600.0 if m + 10 <= minutes else (minutes - m) * 60.0, 48000, seed=12345 * 1000 + m)
                           for m in range(0, minutes, 10)])


def test_process_audio_adaptive_at_baseline_size_vs_oracle(ctx):
    x = make_input(MINUTES)
    pcm, res, an = A.process_audio_adaptive(ctx, x, 48000)
    o = CO.process_audio(x, 48000)
    oan = o["analysis"]
    # (i) measurement parity: the detector's elections and the adapted Pass-2 spec are the same decisions
    va, ova = an.voice_activity, oan["va"]
    assert bool(va.has_speech_profile) == (ova["speech"] is not None) and bool(va.has_noise_profile) == (ova["noise_profile"] is not None)
    if ova["speech"] is not None:
        assert (va.speech_profile.region.start_ns, va.speech_profile.region.end_ns) == ova["speech"]["region"]
    if ova["noise_profile"] is not None:
        assert (va.noise_profile.start_ns, va.noise_profile.duration_ns) == (ova["noise_profile"]["start"], ova["noise_profile"]["duration"])
    assert abs(va.floor - ova["floor"]) < 0.02 and abs(va.split - ova["split"]) < 0.02
    assert abs(res.input.input_i - oan["meas"]["input_i"]) < 0.0011 and abs(res.input.input_lra - oan["meas"]["input_lra"]) < 0.011
    assert abs(res.input.input_tp - oan["meas"]["input_tp"]) < 0.01
    gspec, ospec = an.pass2_spec.decode(), o["spec2"]
    assert [f.split("=")[0] for f in gspec.split(",")] == [f.split("=")[0] for f in ospec.split(",")]     # same chain
    if gspec != ospec:             # printf-rounded parameters may differ in their last digit: compare them as numbers
        for gf, of in zip(gspec.split(","), ospec.split(",")):
            for gkv, okv in zip(gf.partition("=")[2].split(":"), of.partition("=")[2].split(":")):
                gk, _, gv = gkv.partition("="); ok, _, ov = okv.partition("=")
                assert gk == ok
                if gv != ov:
                    gvals, ovals = [float(v) for v in gv.split("|")], [float(v) for v in ov.split("|")]
                    assert all(abs(a - b) <= 0.11 * max(1e-6, abs(b)) + 1e-6 for a, b in zip(gvals, ovals)), (gkv, okv)
    # (ii) kernel parity under the oracle's spec: Pass 2 PCM, then the whole chain end to end
    assert res.n_out == len(pcm) == len(o["pcm"])
    d = (pcm.astype(np.int32) - o["pcm"].astype(np.int32)) / 32768.0
    rms = float(np.sqrt(np.mean(d * d)))
    assert rms < 1e-4, rms                                                     # north_star: 1e-4 RMS of full scale
    assert abs(res.filtered.input_i - o["filtered"][0]) < 0.02 and abs(res.filtered.input_tp - o["filtered"][1]) < 0.1
    for k in ("input_i", "input_tp", "input_lra", "input_thresh"):
        assert abs(getattr(res.pass3, k) - o["p3"][k]) < 0.02, k
    assert res.pass4.normalization_type == o["p4"]["normalization_type"]
    assert abs(res.final.input_i - o["final"][0]) < 0.1 and abs(res.final.input_tp - o["final"][1]) < 0.1      # north_star: +-0.1 LU / dB
    assert abs(res.final.input_lra - o["final"][2]) < 0.1
    # region re-measures (a7) of both stages
    for got, exp in ((an.filtered_regions, o.get("filtered_regions")), (an.final_regions, o.get("final_regions"))):
        if exp is None:
            continue
        for has, g, e in ((got.has_room_tone, got.room_tone, exp["room_tone"]), (got.has_speech, got.speech, exp["speech"])):
            assert bool(has) == (e is not None)
            if e is not None:
                assert abs(g.rms_level - e["rms_level"]) < 0.05 and abs(g.peak_level - e["peak_level"]) < 0.2
                assert abs(g.momentary_lufs - e["momentary_lufs"]) < 0.05 and abs(g.true_peak - e["true_peak"]) < 0.2
