"""ctypes driver for the REAL FFmpeg libswresample bundled with opencv-python-headless.

TEST INFRASTRUCTURE ONLY.  Used (a) by scripts/make_golden_swr.py to generate the
committed golden vectors under tests/golden/ and (b) by tests that pin the oracle's
swr restatement (oracle/orc_swr.c) when the library happens to be present.

The library is FFmpeg 8.0.1 libswresample 6.1.100 (the reference pins FFmpeg 8.1.1,
libswresample 6.3.101, third_party/ffmpeg-statigo/lib/fetch.go:21-23); the native
resampler engine is believed unchanged between the two.
"""
import ctypes as C
import glob
import os
import numpy as np

_LIBDIR_GLOBS = [
    "/opt/prime-rl/.venv/lib/python3.12/site-packages/opencv_python_headless.libs",
]

AV_SAMPLE_FMT = dict(u8=0, s16=1, s32=2, flt=3, dbl=4, u8p=5, s16p=6, s32p=7, fltp=8, dblp=9)
_NP = {1: np.int16, 2: np.int32, 3: np.float32, 4: np.float64}


class AVChannelLayout(C.Structure):
    _fields_ = [("order", C.c_int), ("nb_channels", C.c_int), ("mask", C.c_uint64), ("opaque", C.c_void_p)]


def _layout(nch):
    # AV_CHANNEL_ORDER_NATIVE = 1; mono = FRONT_CENTER (0x4), stereo = FL|FR (0x3)
    return AVChannelLayout(1, nch, {1: 0x4, 2: 0x3}[nch], None)


_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    for d in _LIBDIR_GLOBS:
        if not os.path.isdir(d):
            continue
        for pat in ("libdrm-*", "libcrypto-*", "libssl-*", "libavutil-*"):
            for p in sorted(glob.glob(os.path.join(d, pat))):
                try:
                    C.CDLL(p, mode=C.RTLD_GLOBAL)
                except OSError:
                    pass
        c = sorted(glob.glob(os.path.join(d, "libswresample-*")))
        if c:
            _lib = C.CDLL(c[0], mode=C.RTLD_GLOBAL)
            _lib.swr_alloc_set_opts2.argtypes = [C.POINTER(C.c_void_p), C.POINTER(AVChannelLayout), C.c_int, C.c_int,
                                                 C.POINTER(AVChannelLayout), C.c_int, C.c_int, C.c_int, C.c_void_p]
            _lib.swr_init.argtypes = [C.c_void_p]
            _lib.swr_convert.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.c_int, C.POINTER(C.c_void_p), C.c_int]
            _lib.swr_free.argtypes = [C.POINTER(C.c_void_p)]
            return _lib
    return None


def available():
    return load() is not None


def convert(x, in_fmt, in_rate, out_fmt, out_rate, in_ch=1, out_ch=1, frame=4096, flush=True, per_call_counts=None):
    """Stream interleaved x (n_frames*in_ch) through swr in `frame`-sized calls.
    Returns interleaved output (np array in out_fmt's dtype)."""
    lib = load()
    assert lib is not None
    ctx = C.c_void_p(None)
    lo, li = _layout(out_ch), _layout(in_ch)
    r = lib.swr_alloc_set_opts2(C.byref(ctx), C.byref(lo), AV_SAMPLE_FMT[out_fmt], out_rate,
                                C.byref(li), AV_SAMPLE_FMT[in_fmt], in_rate, 0, None)
    assert r == 0
    assert lib.swr_init(ctx) == 0
    x = np.ascontiguousarray(x, dtype=_NP[AV_SAMPLE_FMT[in_fmt]])
    n = len(x) // in_ch
    odt = _NP[AV_SAMPLE_FMT[out_fmt]]
    outs = []
    cap = int(frame * out_rate / in_rate) + 4096
    obuf = np.zeros(cap * out_ch, dtype=odt)
    pos = 0
    while pos < n:
        m = min(frame, n - pos)
        ip = C.c_void_p(x[pos * in_ch:].ctypes.data)
        op = C.c_void_p(obuf.ctypes.data)
        got = lib.swr_convert(ctx, C.byref(op), cap, C.byref(ip), m)
        assert got >= 0
        if per_call_counts is not None:
            per_call_counts.append(got)
        outs.append(obuf[:got * out_ch].copy())
        pos += m
    if flush:
        while True:
            op = C.c_void_p(obuf.ctypes.data)
            got = lib.swr_convert(ctx, C.byref(op), cap, None, 0)
            if got <= 0:
                break
            if per_call_counts is not None:
                per_call_counts.append(got)
            outs.append(obuf[:got * out_ch].copy())
    lib.swr_free(C.byref(ctx))
    return np.concatenate(outs) if outs else np.zeros(0, dtype=odt)
