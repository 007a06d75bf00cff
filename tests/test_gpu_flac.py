"""GPU parity of the FLAC encoder (csrc/k_flac.cu) -- the container the reference writes its result in
(mono / s16 / 4096-sample frames, encoder.go:92-101): the CUDA stream must equal the sequential oracle's stream BYTE for
byte (same integer decisions), decode bit-exactly with the REAL FFmpeg libavcodec decoder and the oracle's decoder, and
survive the edge cases (digital silence -> CONSTANT, full-scale noise -> VERBATIM, ragged last frame, tiny inputs)."""
import numpy as np
import pytest

import ref_flac
from jivetalking_b200 import gpudsp, synth
from test_oracle_flac import SIG, s16

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", sorted(SIG))
@pytest.mark.parametrize("block_size", [4096, 1024, 100])
def test_stream_equals_the_oracle_byte_for_byte(ctx, name, block_size):
    x = SIG[name]
    got = ctx.flac_encode(x, 44100, block_size)
    exp = ref_flac.encode(x, 44100, block_size)
    if got != exp:
        n = min(len(got), len(exp))
        first = next((i for i in range(n) if got[i] != exp[i]), n)
        raise AssertionError(f"streams differ: len {len(got)} vs {len(exp)}, first difference at byte {first}")
    y, rate = ref_flac.decode(got, len(x) + 16)
    assert rate == 44100 and np.array_equal(x, y)


def test_chain_output_round_trips_through_real_libavcodec(ctx):
    """ProcessAudio's result in its container: four passes -> FLAC -> the reference's own decoder -> the same samples"""
    pcm, res = ctx.process_audio(synth.speech_like(30.0, 48000, seed=77), 48000)
    stream = ctx.flac_encode(pcm, 44100, 4096)
    assert len(stream) < 0.75 * 2 * len(pcm)
    ref = ref_flac.ref_decode(stream)
    if ref is None:
        y, _ = ref_flac.decode(stream, len(pcm) + 16)
    else:
        y, rate, ch = ref
        assert (rate, ch) == (44100, 1)
    assert np.array_equal(y, pcm)
    # frame count / sample count in STREAMINFO
    v = int.from_bytes(stream[18:26], "big")
    assert v & ((1 << 36) - 1) == len(pcm) and v >> 44 == 44100


def test_full_size_round_trip(ctx):
    """BASELINE configs[1] size: 60 min of 44.1 kHz s16 (38 760 frames, three-byte frame numbers) -> stream -> real decoder"""
    seg = s16(synth.speech_like(60.0, 44100, seed=11))
    seg = seg[: len(seg) // 4096 * 4096]
    x = np.tile(seg, 60 * 60 * 44100 // len(seg) + 1)[: 38760 * 4096]
    x = x.copy()
    x[10_000_000:10_500_000] = 0                             # a stretch of digital silence
    stream = ctx.flac_encode(x, 44100, 4096)
    ref = ref_flac.ref_decode(stream)
    if ref is None:
        pytest.skip("no libavcodec on this box: the byte-for-byte and oracle-decoder tests cover the encoder")
    assert np.array_equal(ref[0], x)
    # size-independent property: the stream is the concatenation of independently coded frames -- a 100-frame slice from the
    # middle, re-encoded on its own, differs only in the frame numbers (and therefore the CRCs)
    a = 20000 * 4096
    part = ctx.flac_encode(x[a: a + 100 * 4096], 44100, 4096)
    y, _ = ref_flac.decode(part, 100 * 4096 + 16)
    assert np.array_equal(y, x[a: a + 100 * 4096])


def test_errors(ctx):
    with pytest.raises(gpudsp.JtError) as e:
        ctx.flac_encode(np.zeros(100000, dtype=np.int16), 44100, 8192)
    assert e.value.code == -5
    assert ctx.flac_encode(np.zeros(0, dtype=np.int16), 44100, 4096)[:4] == b"fLaC"
