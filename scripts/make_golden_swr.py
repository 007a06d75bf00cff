"""Generate tests/golden/swr_golden.npz from the REAL FFmpeg libswresample (8.0.1, 6.1.100)
bundled with opencv-python-headless in this image (oracle/ref_swr.py).  Run here, commit the
result: the GPU box and CI then pin oracle/orc_swr.c (and through it the CUDA resampler)
against real-FFmpeg outputs without needing the library."""
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, ROOT)
import ref_swr  # noqa: E402
from jivetalking_b200 import synth  # noqa: E402

assert ref_swr.available(), "libswresample not found"
out = {}
x16 = synth.reference_test_audio(0.25, 48000, 440.0, -23.0, -60.0, 0.1, 0.05)
xf = (x16.astype(np.float64) / 32768.0)
rng = np.random.default_rng(7)
noise = rng.standard_normal(6000) * 0.2
for name, sig in (("tone", xf), ("noise", noise)):
    out[f"in_{name}"] = sig
    for (ir, orr) in ((48000, 192000), (48000, 44100), (44100, 192000), (96000, 44100)):
        for flush in (0, 1):
            y = ref_swr.convert(sig, "dbl", ir, "dbl", orr, frame=4096, flush=bool(flush))
            out[f"dbl_{name}_{ir}_{orr}_{flush}"] = y
    # s16 -> dbl at 192 kHz runs swr's FLTP internal path (loudnorm dynamic mode on s16 input)
s16 = np.clip(np.round(noise * 32768), -32768, 32767).astype(np.int16)
out["in_s16"] = s16
out["s16_to_dbl_44100_192000"] = ref_swr.convert(s16, "s16", 44100, "dbl", 192000, frame=4096, flush=True)
out["flt_44100_192000"] = ref_swr.convert(noise.astype(np.float32), "flt", 44100, "flt", 192000, frame=4096, flush=True)
# format conversions and downmix
out["dbl_to_s16"] = ref_swr.convert(noise * 4.0, "dbl", 48000, "s16", 48000)
out["flt_to_s16"] = ref_swr.convert((noise * 4.0).astype(np.float32), "flt", 48000, "s16", 48000)
st = np.empty(2 * 3000, dtype=np.float32); st[0::2] = noise[:3000]; st[1::2] = noise[3000:]
out["in_stereo_f32"] = st
out["stereo_f32_to_mono"] = ref_swr.convert(st, "flt", 48000, "flt", 48000, in_ch=2, out_ch=1)
st16 = np.clip(np.round(st * 32768), -32768, 32767).astype(np.int16)
out["in_stereo_s16"] = st16
out["stereo_s16_to_mono"] = ref_swr.convert(st16, "s16", 48000, "s16", 48000, in_ch=2, out_ch=1)
# 32-bit integer samples (what 24-bit FLAC / WAV decode to): conversions, the flt-internal normalised downmix, and the
# flt-internal resample swr picks for a 32-bit integer input (swr_init's int_sample_fmt rule)
s32 = np.clip(np.round(noise * 0.9 * 2 ** 31), -2 ** 31, 2 ** 31 - 1).astype(np.int32)
s32[:6] = [2 ** 31 - 1, -2 ** 31, 2 ** 31 - 129, 12345678, -1, 1]
out["in_s32"] = s32
out["s32_to_flt"] = ref_swr.convert(s32, "s32", 48000, "flt", 48000)
out["s32_to_dbl"] = ref_swr.convert(s32, "s32", 48000, "dbl", 48000)
out["s32_to_s16"] = ref_swr.convert(s32, "s32", 48000, "s16", 48000)
out["flt_to_s32"] = ref_swr.convert((noise * 4.0).astype(np.float32), "flt", 48000, "s32", 48000)
out["dbl_to_s32"] = ref_swr.convert(noise * 4.0, "dbl", 48000, "s32", 48000)
st32 = np.empty(2 * 3000, dtype=np.int32); st32[0::2] = s32[:3000]; st32[1::2] = s32[3000:]
out["in_stereo_s32"] = st32
out["stereo_s32_to_mono"] = ref_swr.convert(st32, "s32", 48000, "s32", 48000, in_ch=2, out_ch=1)
out["s32_to_dbl_48000_192000"] = ref_swr.convert(s32, "s32", 48000, "dbl", 192000, frame=4096, flush=True)
# loudnorm's dynamic-mode output leaves at 192 kHz / dbl: the aresample barrier back to the source rate
for (ir, orr) in ((192000, 44100), (192000, 48000)):
    out[f"dbl_noise_{ir}_{orr}_1"] = ref_swr.convert(noise, "dbl", ir, "dbl", orr, frame=4096, flush=True)
# inexact ratios (reduced phase count > 1024): 1024 phases, fractional index advance, resample_linear
for (ir, orr) in ((22050, 192000), (11025, 192000), (47999, 44100)):
    for flush in (0, 1):
        out[f"dbl_noise_{ir}_{orr}_{flush}"] = ref_swr.convert(noise[:2000], "dbl", ir, "dbl", orr, frame=700, flush=bool(flush))
out["flt_22050_192000"] = ref_swr.convert(noise[:2000].astype(np.float32), "flt", 22050, "flt", 192000, frame=700, flush=True)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "swr_golden.npz"), **out)
print("wrote", len(out), "arrays")
