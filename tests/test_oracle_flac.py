"""The oracle's FLAC encoder / decoder (oracle/orc_flac.c) pinned on the REAL FFmpeg libavcodec FLAC decoder found in this
image (oracle/ref_flac.py builds a probe against the reference's vendored headers): every stream the oracle encoder writes
must come back bit for bit from libavcodec and from the oracle's own decoder.  CPU only."""
import os

import numpy as np
import pytest

import ref_flac
from jivetalking_b200 import synth


def s16(x):
    return np.clip(np.round(x * 32768.0), -32768, 32767).astype(np.int16)


def signals():
    sp = s16(synth.speech_like(3.0, 44100, seed=5))
    rng = np.random.default_rng(3)
    out = {"speech": sp[: 4096 * 20], "ragged": sp[:50000], "one_frame": sp[:4096], "tiny": sp[:100],
           "silence": np.zeros(4096 * 3, dtype=np.int16), "dc": np.full(9000, -1234, dtype=np.int16),
           "white_full_scale": rng.integers(-32768, 32768, 4096 * 2).astype(np.int16),
           "clipped": s16(np.clip(3.0 * synth.speech_like(1.0, 44100, seed=9), -1, 1)),
           "extremes": np.tile(np.array([-32768, 32767], dtype=np.int16), 4096)}
    mixed = sp[: 4096 * 6].copy()
    mixed[4096:8192] = 0
    mixed[3 * 4096: 3 * 4096 + 2000] = 321
    out["mixed"] = mixed
    return out


SIG = signals()


@pytest.mark.parametrize("name", sorted(SIG))
@pytest.mark.parametrize("block_size", [4096, 1024, 100])
def test_round_trip_oracle_and_real_libavcodec(name, block_size):
    x = SIG[name]
    stream = ref_flac.encode(x, 44100, block_size)
    assert stream[:4] == b"fLaC" and stream[4] == 0x80 and stream[7] == 34
    y, rate = ref_flac.decode(stream, len(x) + 16)
    assert rate == 44100 and np.array_equal(x, y)
    ref = ref_flac.ref_decode(stream)
    if ref is None:
        pytest.skip("FFmpeg libavcodec / reference headers not present: oracle decoder only")
    pcm, rrate, ch = ref
    assert (rrate, ch) == (44100, 1) and np.array_equal(pcm, x)


def test_streaminfo_and_compression():
    x = SIG["speech"]
    stream = ref_flac.encode(x, 44100, 4096)
    si = stream[8:42]
    assert int.from_bytes(si[0:2], "big") == 4096 and int.from_bytes(si[2:4], "big") == 4096
    min_fs, max_fs = int.from_bytes(si[4:7], "big"), int.from_bytes(si[7:10], "big")
    assert 0 < min_fs <= max_fs < 8212
    v = int.from_bytes(si[10:18], "big")
    assert v >> 44 == 44100 and (v >> 41) & 7 == 0 and (v >> 36) & 31 == 15 and v & ((1 << 36) - 1) == len(x)
    assert si[18:34] == bytes(16)                           # MD5 "not known"
    assert len(stream) < 0.7 * 2 * len(x)                   # a speech-like signal compresses
    assert len(ref_flac.encode(SIG["silence"], 44100, 4096)) == 42 + 3 * (6 + 3 + 2)     # CONSTANT subframes
    assert len(ref_flac.encode(SIG["white_full_scale"], 44100, 4096)) == 42 + 2 * (6 + 1 + 8192 + 2)   # VERBATIM


def test_decoder_rejects_corruption():
    x = SIG["speech"][: 4096 * 3]
    stream = bytearray(ref_flac.encode(x, 44100, 4096))
    bad = bytearray(stream); bad[42 + 2] ^= 0x10            # frame header -> CRC-8
    with pytest.raises(ValueError):
        ref_flac.decode(bytes(bad), len(x) + 16)
    bad = bytearray(stream); bad[42 + 500] ^= 0x01          # residual bits -> CRC-16 (or a syntax error on the way)
    with pytest.raises(ValueError):
        ref_flac.decode(bytes(bad), len(x) + 16)


@pytest.mark.parametrize("name", ["speech", "ragged", "clipped", "mixed", "extremes", "silence", "white_full_scale"])
@pytest.mark.parametrize("level", [0, 5, 8])
def test_oracle_decoder_on_real_libavcodec_streams(name, level):
    """the other direction: streams written by the REAL libavcodec encoder (LPC subframes at the reference's level 5) decode
    bit for bit with the oracle's decoder -- the checker a GPU FLAC decoder of the input side will be held against"""
    x = SIG[name]
    stream = ref_flac.ref_encode(x, 44100, level)
    if stream is None:
        pytest.skip("FFmpeg libavcodec / reference headers not present")
    y, rate = ref_flac.decode(stream, len(x) + 16)
    assert rate == 44100 and np.array_equal(x, y)


# ---- the INPUT side: orc_flac_decode_pcm (checker of csrc/k_flac_dec.cu) ---------------------------------------------------
import flac_synth  # noqa: E402

GOLDEN_DEC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "flac_dec_golden.npz")


def test_input_decoder_on_the_golden_streams():
    """streams of the REAL libavcodec encoder (levels 0 / 5 / 8, mono / stereo, 16 / 24 bit) and of the stream generator, with the
    samples the REAL libavformat + libavcodec reader decoded from them (scripts/make_golden_flac.py)"""
    g = np.load(GOLDEN_DEC)
    names = sorted(k[:-7] for k in g.files if k.endswith("_stream"))
    assert len(names) >= 13
    for name in names:
        stream, want = g[name + "_stream"].tobytes(), g[name + "_pcm"]
        pcm, rate, ch, bps = ref_flac.decode_pcm(stream, len(want))
        assert pcm.dtype == want.dtype and np.array_equal(pcm, want), name


@pytest.mark.parametrize("seed", range(6))
def test_input_decoder_equals_the_real_reader_on_generated_streams(seed):
    """every subframe type, LPC orders up to 32, wasted bits, escape partitions, variable block sizes, all channel assignments,
    8..24 bits: oracle == the reference's reader (libavformat flac demuxer + libavcodec flacdec), where the probe exists"""
    ch, bps = ((1, 16), (2, 16), (2, 24), (2, 8), (3, 12), (8, 20))[seed]
    s = flac_synth.make_stream(seed, n_frames=5, channels=ch, bps=bps, rate=(44100, 48000, 37800)[seed % 3], variable=bool(seed & 1),
                               metadata_pad=(seed % 3) * 64)
    pcm, rate, c, b = ref_flac.decode_pcm(s, 40000)
    assert (c, b) == (ch, bps)
    ref = ref_flac.ref_wav_read(s)
    if ref is None:
        pytest.skip("no libavformat / libavcodec probe on this box")
    assert ref[1:] == (rate, ch) and np.array_equal(ref[0], pcm)


def test_input_decoder_rejects_corruption():
    s = bytearray(flac_synth.make_stream(3, n_frames=3, channels=2, bps=16))
    s[len(s) // 2] ^= 0x10
    with pytest.raises(ValueError):
        ref_flac.decode_pcm(bytes(s), 40000)


# ---- LPC subframes and the MD5 field ------------------------------------------------------------------------------------------
def test_lpc_streams_are_no_larger_than_libavcodec_level_5():
    """the encoder's LPC subframes (orders 1..8 scored by their measured residual) against the REAL libavcodec encoder at the
    reference's compression_level 5 (encoder.go:92-101), on the synthetic recipes: ours must not be larger"""
    for name, x in (("speech", synth.speech_like(20.0, 44100, seed=3)), ("podcast", synth.podcast_like(20.0, 44100, seed=4))):
        pcm = s16(x)
        ours = ref_flac.encode(pcm, 44100, 4096)
        y, _ = ref_flac.decode(ours, len(pcm) + 16)
        assert np.array_equal(y, pcm)
        lav = ref_flac.ref_encode(pcm, 44100, 5)
        if lav is None:
            pytest.skip("no libavcodec encoder probe on this box")
        assert len(ours) <= len(lav), (name, len(ours), len(lav))
        assert len(ours) < 0.96 * len(ref_flac.ref_encode(pcm, 44100, 0))          # level 0 = fixed predictors only
        ref = ref_flac.ref_decode(ours)
        assert ref is not None and np.array_equal(ref[0], pcm)


def test_md5_field():
    import hashlib
    from jivetalking_b200 import gpudsp
    rng = np.random.default_rng(2)
    for n in (0, 1, 27, 28, 55, 56, 63, 64, 65, 4096, 44100):
        pcm = rng.integers(-32768, 32767, size=n).astype(np.int16)
        st = ref_flac.encode(pcm, 44100, 4096)
        out = gpudsp.flac_set_md5(st, pcm)
        assert out[26:42] == hashlib.md5(pcm.tobytes()).digest() and out[:26] == st[:26] and out[42:] == st[42:]
    y, _ = ref_flac.decode(out, 44100 + 16)
    assert np.array_equal(y, pcm)
