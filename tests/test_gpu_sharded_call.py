"""jt_process_audio_sharded (BASELINE.json configs[3]): the whole ProcessAudio of ONE stream, one call per rank, the
orchestration inside the library.  Here every rank is a thread with its own jt_ctx on the one GPU of the test box and the
all-gather is a barrier between the threads; the result must be what the unsharded jt_process_audio_adaptive gives."""
import math

import numpy as np
import pytest

from jivetalking_b200 import adapt as A
from jivetalking_b200 import gpudsp, shard, synth

pytestmark = pytest.mark.gpu


def _compare(pcm, infos, x, rate, channels, ctx):
    ref, res1, an1 = A.process_audio_adaptive(ctx, x, rate, channels)
    assert len(pcm) == len(ref)
    res1_ = res1
    d = (pcm.astype(np.int32) - ref.astype(np.int32)) / 32768.0
    assert float(np.sqrt(np.mean(d * d))) < 1e-4                     # north_star: 1e-4 RMS of full scale
    # Pass 4's gain is printed with two decimals of a dB from Pass 3's input_i: a chunked meter that differs in the
    # third decimal can tip that rounding and scale the whole stream by 0.01 dB; only with equal strings are the samples
    # expected to agree to the LSB almost everywhere
    # (at 96 kHz the 80 Hz f32 TDII high-pass carries ~3e-4 of its own round-off noise and two runs started from different
    #  states never re-merge -- scripts/debug/chunk_filter_diff.py -- so there only the contract's RMS bound applies)
    r0 = infos[0][0]
    if rate <= 48000 and all("%.2f" % getattr(r0.pass3, k) == "%.2f" % getattr(res1.pass3, k) for k in ("input_i", "input_tp", "input_lra", "input_thresh")):
        assert np.mean(np.abs(d) > 2.5 / 32768.0) < 2e-3
    for res, an, tm in infos:                                        # every rank holds the same merged measurements
        assert an.pass2_spec == an1.pass2_spec
        va, va1 = an.voice_activity, an1.voice_activity
        # (the band graphs sum their squares with atomics: the 17 band values are equal to the last bits, not always bit for bit)
        assert (va.floor, va.split, va.margin, va.voiced_low_percentile, va.gate_separation_db, va.voice_activated, va.n_speech_regions) == \
               (va1.floor, va1.split, va1.margin, va1.voiced_low_percentile, va1.gate_separation_db, va1.voice_activated, va1.n_speech_regions)
        assert bytes(va.noise_region) == bytes(va1.noise_region) and bytes(va.speech_profile.region) == bytes(va1.speech_profile.region)
        assert np.allclose(list(va.noise_profile.band_noise), list(va1.noise_profile.band_noise), rtol=0, atol=1e-6, equal_nan=True)
        assert abs(res.input.input_i - res1.input.input_i) < 1e-9 and abs(res.input.input_tp - res1.input.input_tp) < 1e-9
        assert abs(res.filtered.input_i - res1.filtered.input_i) < 0.0011
        for k in ("input_i", "input_tp", "input_lra", "input_thresh"):
            assert abs(getattr(res.pass3, k) - getattr(res1.pass3, k)) < 2e-3, k
        assert res.pass4.normalization_type == res1.pass4.normalization_type == 0
        assert abs(res.final.input_i - res1.final.input_i) < 0.0021 and abs(res.final.input_tp - res1.final.input_tp) < 0.05
        assert abs(res.final.input_lra - res1.final.input_lra) < 0.011
        assert res.n_out == res1.n_out
        for got, exp in ((an.filtered_regions, an1.filtered_regions), (an.final_regions, an1.final_regions)):
            assert (got.has_room_tone, got.has_speech) == (exp.has_room_tone, exp.has_speech)
            for g, e in ((got.room_tone, exp.room_tone), (got.speech, exp.speech)):
                assert abs(g.rms_level - e.rms_level) < 0.02 and abs(g.momentary_lufs - e.momentary_lufs) < 0.02
                assert abs(g.true_peak - e.true_peak) < 0.1
    return res1


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_call_equals_single_gpu_mono(ctx, world):
    x = synth.podcast_like(150.0, 48000, seed=71)
    ctxs = [gpudsp.Context(0) for _ in range(world)]
    try:
        pcm, infos = shard.process_stream_sharded_call(ctxs, x, 48000)
    finally:
        for c in ctxs:
            c.close()
    _compare(pcm, infos, x, 48000, 1, ctx)
    tm = infos[0][2]
    assert tm.exchange_calls > 0 and 0 < tm.halo_bytes < 2 * 60 * 44100 * 2          # halos, never the stream


def test_sharded_call_stereo_96k(ctx):
    x = synth.stereo_from_mono(synth.podcast_like(100.0, 96000, seed=72))
    ctxs = [gpudsp.Context(0) for _ in range(2)]
    try:
        pcm, infos = shard.process_stream_sharded_call(ctxs, x, 96000, channels=2)
    finally:
        for c in ctxs:
            c.close()
    _compare(pcm, infos, x, 96000, 2, ctx)


def test_sharded_plan_tiles_the_stream():
    for total, rate, world in ((150 * 48000, 48000, 3), (3 * 3600 * 96000, 96000, 8), (61 * 60 * 48000 + 17, 48000, 5)):
        plans = [A.sharded_plan(total, rate, world, r) for r in range(world)]
        pos = 0
        for p in plans:
            assert p.own_first == pos and p.own_first % p.unit == 0 and p.local_first % p.unit == 0
            assert p.local_first <= p.own_first and p.local_first + p.n_local >= p.own_first + p.owned
            assert p.local_first + p.n_local <= total
            pos += p.owned
        assert pos == total


def test_one_rank_is_the_unsharded_chain(ctx):
    x = synth.podcast_like(60.0, 48000, seed=73)
    own, first, res, an, tm = A.process_audio_sharded(ctx, x, 48000, 1, len(x), 1, 0)
    ref, res1, an1 = A.process_audio_adaptive(ctx, x, 48000)
    assert first == 0 and np.array_equal(own, ref) and an.pass2_spec == an1.pass2_spec
