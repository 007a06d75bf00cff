"""GPU parity of the input-side decoders (csrc/k_flac_dec.cu, jt_wav_decode; SURVEY 8f-3, the reference's audio.Reader:
internal/audio/reader.go:29-188).  The CUDA FLAC decoder (parallel header scan -> CRC-16 linking -> thread per frame) must give
the samples the oracle decoder gives -- which is pinned on the REAL libavformat + libavcodec reader (tests/test_oracle_flac.py,
tests/golden/flac_dec_golden.npz) -- bit for bit, on real-encoder streams, on generated streams that exercise every syntax
element, on its own encoder's output at full size, and must refuse damaged streams."""
import os
import struct

import numpy as np
import pytest

import flac_synth
import ref_flac
from jivetalking_b200 import gpudsp, synth

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "flac_dec_golden.npz")


def test_golden_streams_of_the_real_encoder(ctx):
    g = np.load(GOLDEN)
    names = sorted(k[:-7] for k in g.files if k.endswith("_stream"))
    for name in names:
        stream, want = g[name + "_stream"].tobytes(), g[name + "_pcm"]
        pcm, fmt, rate, ch = ctx.flac_decode(stream, cap_frames=len(want))
        assert pcm.dtype == want.dtype and np.array_equal(pcm, want), name
        assert fmt == (gpudsp.FMT_S16 if want.dtype == np.int16 else gpudsp.FMT_S32)


@pytest.mark.parametrize("seed", range(16))
def test_generated_streams_equal_the_oracle(ctx, seed):
    ch, bps = ((1, 16), (2, 16), (2, 24), (2, 8), (3, 12), (8, 20), (2, 20), (1, 24))[seed % 8]
    s = flac_synth.make_stream(1000 + seed, n_frames=12, channels=ch, bps=bps, rate=(44100, 48000, 37800, 96000)[seed % 4],
                               variable=bool(seed & 1), metadata_pad=(seed % 3) * 1000)
    want, rate, c, b = ref_flac.decode_pcm(s, 60000)
    pcm, fmt, grate, gch = ctx.flac_decode(s)
    assert (grate, gch) == (rate, ch) and pcm.dtype == want.dtype
    assert np.array_equal(pcm, want)
    info = gpudsp.flac_stream_info(s)
    assert (info["rate"], info["channels"], info["bits"], info["n_frames"]) == (rate, ch, bps, len(want) // ch)


def test_real_encoder_stereo_24_bit_minutes(ctx):
    """two minutes of stereo 24-bit speech-like audio through the REAL libavcodec encoder (level 5, the reference's setting for
    its own output and the usual setting of its input files) -> CUDA decoder == the original samples"""
    x = synth.podcast_like(120.0, 48000, seed=5)
    l = np.clip(np.round(x * (1 << 23)), -(1 << 23), (1 << 23) - 1).astype(np.int32)
    r = np.roll(l, 7) // 2
    st = (np.stack([l, r], 1).reshape(-1) << 8).astype(np.int32)
    s = ref_flac.ref_encode(st, 48000, 5, channels=2, bits=24)
    if s is None:
        pytest.skip("no libavcodec encoder probe on this box")
    pcm, fmt, rate, ch = ctx.flac_decode(s, cap_frames=len(l))
    assert (fmt, rate, ch) == (gpudsp.FMT_S32, 48000, 2) and np.array_equal(pcm, st)


def test_own_encoder_round_trip_full_size(ctx):
    """BASELINE configs[1] size: the chain's 60 min / 44.1 kHz / s16 output -> jt_flac_encode -> jt_flac_decode == the samples
    (38 760 frames; the stretch of digital silence gives 14-byte CONSTANT frames, the densest candidate list)"""
    seg = np.clip(np.round(synth.speech_like(60.0, 44100, seed=11) * 32767), -32768, 32767).astype(np.int16)
    seg = seg[: len(seg) // 4096 * 4096]
    x = np.tile(seg, 60 * 60 * 44100 // len(seg) + 1)[: 38760 * 4096].copy()
    x[10_000_000:10_500_000] = 0
    x = x[: len(x) - 1234]                                   # ragged last frame
    stream = ctx.flac_encode(x, 44100, 4096)
    pcm, fmt, rate, ch = ctx.flac_decode(stream)
    assert (fmt, rate, ch) == (gpudsp.FMT_S16, 44100, 1) and np.array_equal(pcm, x)


def test_decoded_input_feeds_the_chain(ctx):
    """file image -> jt_flac_decode -> jt_analyse == jt_analyse on the samples the file was made from"""
    x = synth.podcast_like(20.0, 44100, seed=9)
    s16 = np.clip(np.round(x * 32767), -32768, 32767).astype(np.int16)
    stream = ctx.flac_encode(s16, 44100, 4096)
    pcm, fmt, rate, ch = ctx.flac_decode(stream)
    a = ctx.run_graph(gpudsp.pass1_spec(), pcm, rate, want_pcm=False)
    b = ctx.run_graph(gpudsp.pass1_spec(), s16, 44100, want_pcm=False)
    assert [m.r128_I for m in a["meta"]][-1] == [m.r128_I for m in b["meta"]][-1]


def test_damaged_streams_are_refused(ctx):
    s = bytearray(flac_synth.make_stream(7, n_frames=6, channels=2, bps=16))
    bad = bytearray(s); bad[len(s) // 2] ^= 0x04             # payload bit flip: the frame's CRC-16 no longer links
    with pytest.raises(gpudsp.JtError) as e:
        ctx.flac_decode(bytes(bad))
    assert e.value.code == -1
    with pytest.raises(gpudsp.JtError):
        ctx.flac_decode(bytes(s[: len(s) - 37]))             # truncated
    with pytest.raises(gpudsp.JtError):
        ctx.flac_decode(b"RIFF" + bytes(100))
    with pytest.raises(gpudsp.JtError) as e:                 # output too small
        ctx.flac_decode(bytes(s), cap_frames=100)
    assert e.value.code == -7


def _wav(fmt_tag, bits, ch, rate, payload):
    hdr = struct.pack("<4sI4s4sIHHIIHH", b"RIFF", 36 + len(payload), b"WAVE", b"fmt ", 16, fmt_tag, ch, rate, rate * ch * bits // 8, ch * bits // 8, bits)
    return hdr + struct.pack("<4sI", b"data", len(payload)) + payload


def test_wav_decode_24_bit_and_others(ctx):
    rng = np.random.default_rng(3)
    v = rng.integers(-(1 << 23), 1 << 23, size=2 * 4801).astype(np.int32)
    b = np.zeros((len(v), 3), dtype=np.uint8)
    b[:, 0], b[:, 1], b[:, 2] = v & 0xFF, (v >> 8) & 0xFF, (v >> 16) & 0xFF
    img = _wav(1, 24, 2, 48000, b.tobytes())
    pcm, fmt, rate, ch = ctx.wav_decode(img)
    assert (fmt, rate, ch) == (gpudsp.FMT_S32, 48000, 2) and np.array_equal(pcm, (v << 8).astype(np.int32))
    ref = ref_flac.ref_wav_read(img)
    if ref is not None:                                      # the reference's reader (libavformat wav demuxer + pcm_s24le)
        assert np.array_equal(ref[0], pcm) and ref[1:] == (48000, 2)
    f = rng.normal(0, 0.1, 5000).astype(np.float32)
    pcm, fmt, rate, ch = ctx.wav_decode(_wav(3, 32, 1, 44100, f.tobytes()))
    assert fmt == gpudsp.FMT_FLT and np.array_equal(pcm, f)
    s = rng.integers(-32768, 32767, size=3001).astype(np.int16)
    pcm, fmt, rate, ch = ctx.wav_decode(_wav(1, 16, 1, 44100, s.tobytes()))
    assert fmt == gpudsp.FMT_S16 and np.array_equal(pcm, s)
    with pytest.raises(gpudsp.JtError):
        ctx.wav_decode(_wav(1, 8, 1, 44100, bytes(100)))
