#!/usr/bin/env python
"""Small driver for ncu captures of the FLAC kernels: encodes 10 min of speech-like s16 audio twice."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from jivetalking_b200 import gpudsp, synth

seg = np.clip(np.round(synth.speech_like(60.0, 44100, seed=11) * 32768.0), -32768, 32767).astype(np.int16)
x = np.tile(seg[: len(seg) // 4096 * 4096], 10)
with gpudsp.Context(0) as ctx:
    s = ctx.flac_encode(x, 44100, 4096)
    ctx.enable_timing(True)
    ctx.reset_counters()
    for _ in range(3):
        s = ctx.flac_encode(x, 44100, 4096)
    print({t[0]: round(t[1] / 3, 4) for t in ctx.kernel_timings() if t[0].startswith("flac")})
print(len(x), len(s), len(s) / (2.0 * len(x)))
