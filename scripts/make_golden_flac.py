"""Golden vectors for the input-side FLAC decoder (tests/golden/flac_dec_golden.npz), made in the container that has the real
FFmpeg libraries and the reference's headers: streams written by the REAL libavcodec FLAC encoder (the codec of the reference's
own input files; levels 0 / 5 / 8, mono and stereo, 16 and 24 bit) and the samples the REAL libavformat + libavcodec reader
(oracle/ref_wav_probe.c -- the reference's audio.Reader path, internal/audio/reader.go:29-188) decodes from them and from a few
streams of oracle/flac_synth.py.  Small on purpose (~0.4 s each)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import flac_synth as FS  # noqa: E402
import ref_flac as RF  # noqa: E402
from jivetalking_b200 import synth  # noqa: E402


def main():
    out = {}
    x = synth.podcast_like(0.4, 44100, seed=21)
    l16 = np.clip(np.round(x * 32767), -32768, 32767).astype(np.int16)
    r16 = (np.roll(l16, 5) // 2).astype(np.int16)
    cases = {
        "mono16": (l16, 1, 16), "stereo16": (np.stack([l16, r16], 1).reshape(-1), 2, 16),
        "stereo24": ((np.stack([l16.astype(np.int32) * 256 + 91, r16.astype(np.int32) * 256 - 17], 1).reshape(-1) << 8).astype(np.int32), 2, 24),
    }
    for name, (pcm, ch, bits) in cases.items():
        for level in (0, 5, 8):
            s = RF.ref_encode(pcm, 44100 if bits == 16 else 48000, level, channels=ch, bits=bits)
            ref = RF.ref_wav_read(s)
            assert ref is not None and np.array_equal(ref[0], pcm), name
            out[f"{name}_l{level}_stream"] = np.frombuffer(s, dtype=np.uint8)
            out[f"{name}_l{level}_pcm"] = ref[0]
    for seed, (ch, bps, var) in enumerate(((2, 16, False), (2, 24, True), (1, 12, True), (6, 20, False))):
        s = FS.make_stream(100 + seed, n_frames=4, channels=ch, bps=bps, rate=48000, variable=var, block_sizes=(1024, 576, 256))
        ref = RF.ref_wav_read(s)
        assert ref is not None
        out[f"synth{seed}_stream"] = np.frombuffer(s, dtype=np.uint8)
        out[f"synth{seed}_pcm"] = ref[0]
    p = os.path.join(ROOT, "tests", "golden", "flac_dec_golden.npz")
    np.savez_compressed(p, **out)
    print(p, os.path.getsize(p), "bytes,", len(out) // 2, "streams")


if __name__ == "__main__":
    main()
