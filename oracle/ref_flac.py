"""TEST INFRASTRUCTURE ONLY: the oracle's FLAC encoder / decoder (oracle/orc_flac.c) and a driver for the REAL FFmpeg
libavcodec FLAC decoder found in this image (opencv-python-headless bundles FFmpeg 8.0.1: libavcodec 62.11.100; the reference
pins FFmpeg 8.1.1, same major).  build_ref() compiles oracle/ref_flac_probe.c against the reference's vendored headers
(/root/reference/third_party/ffmpeg-statigo/include) into oracle/_ref/ref_flac_decode -- only in the container that has
/root/reference; the GPU box uses the prebuilt binary that travels with the snapshot."""
import ctypes as C
import glob
import os
import subprocess
import tempfile

import numpy as np

import jt_oracle as O

_HERE = os.path.dirname(os.path.abspath(__file__))
_REF_BIN = os.path.join(_HERE, "_ref", "ref_flac_decode")
_LIBDIR = "/opt/prime-rl/.venv/lib/python3.12/site-packages/opencv_python_headless.libs"
_REF_INC = "/root/reference/third_party/ffmpeg-statigo/include"

_bound = False


def _lib():
    global _bound
    L = O.lib()
    if not _bound:
        L.orc_flac_max_bytes.restype = C.c_int64
        L.orc_flac_max_bytes.argtypes = [C.c_int64, C.c_int]
        L.orc_flac_encode_s16.restype = C.c_int64
        L.orc_flac_encode_s16.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_int64]
        L.orc_flac_decode_s16.restype = C.c_int64
        L.orc_flac_decode_s16.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.POINTER(C.c_int)]
        L.orc_flac_decode_pcm.restype = C.c_int64
        L.orc_flac_decode_pcm.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
        _bound = True
    return L


def encode(pcm, rate=44100, block_size=4096):
    """oracle encoder -> bytes"""
    pcm = np.ascontiguousarray(pcm, dtype=np.int16)
    cap = _lib().orc_flac_max_bytes(len(pcm), block_size)
    out = np.zeros(cap, dtype=np.uint8)
    n = _lib().orc_flac_encode_s16(pcm.ctypes.data_as(C.c_void_p), len(pcm), rate, block_size, out.ctypes.data_as(C.c_void_p), cap)
    if n < 0:
        raise RuntimeError("orc_flac_encode_s16 failed")
    return out[:n].tobytes()


def decode(stream, max_samples):
    """oracle decoder -> (int16 array, rate); raises on a malformed stream (CRC-8 -2, CRC-16 -3, syntax -4)"""
    buf = np.frombuffer(stream, dtype=np.uint8)
    out = np.zeros(max_samples, dtype=np.int16)
    rate = C.c_int(0)
    n = _lib().orc_flac_decode_s16(buf.ctypes.data_as(C.c_void_p), len(buf), out.ctypes.data_as(C.c_void_p), max_samples, C.byref(rate))
    if n < 0:
        raise ValueError(f"orc_flac_decode_s16: {n}")
    return out[:n], rate.value


def decode_pcm(stream, max_frames):
    """oracle input-side decoder -> (interleaved samples as libavcodec hands them out: int16 for <= 16 bit streams, int32 above,
    rate, channels, bits per sample); raises on a malformed stream"""
    buf = np.frombuffer(stream, dtype=np.uint8)
    rate, ch, bps = C.c_int(0), C.c_int(0), C.c_int(0)
    out = np.zeros(max(1, max_frames) * 8, dtype=np.int32)
    n = _lib().orc_flac_decode_pcm(buf.ctypes.data_as(C.c_void_p), len(buf), out.ctypes.data_as(C.c_void_p), max_frames,
                                   C.byref(rate), C.byref(ch), C.byref(bps))
    if n < 0:
        raise ValueError(f"orc_flac_decode_pcm: {n}")
    pcm = out[: n * ch.value]
    return (pcm.astype(np.int16) if bps.value <= 16 else pcm.copy()), rate.value, ch.value, bps.value


def build_ref():
    """Compile the probe when the reference headers and the bundled FFmpeg are both present; returns the binary or None."""
    if os.path.exists(_REF_BIN):
        return _REF_BIN
    libs = [sorted(glob.glob(os.path.join(_LIBDIR, p))) for p in ("libavcodec-*", "libavutil-*", "libswresample-*")]
    if not os.path.isdir(_REF_INC) or not all(libs):
        return None
    os.makedirs(os.path.dirname(_REF_BIN), exist_ok=True)
    cmd = ["gcc", "-O1", "-o", _REF_BIN, os.path.join(_HERE, "ref_flac_probe.c"), "-I" + _REF_INC] + [l[0] for l in libs] + \
          ["-Wl,-rpath," + _LIBDIR, "-Wl,--allow-shlib-undefined", "-lm"]
    try:
        subprocess.check_call(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    except (subprocess.CalledProcessError, OSError):
        return None
    return _REF_BIN


_REF_ENC = os.path.join(_HERE, "_ref", "ref_flac_encode")


def build_ref_encoder():
    if os.path.exists(_REF_ENC):
        return _REF_ENC
    libs = [sorted(glob.glob(os.path.join(_LIBDIR, p))) for p in ("libavcodec-*", "libavutil-*", "libswresample-*")]
    if not os.path.isdir(_REF_INC) or not all(libs):
        return None
    os.makedirs(os.path.dirname(_REF_ENC), exist_ok=True)
    cmd = ["gcc", "-O1", "-o", _REF_ENC, os.path.join(_HERE, "ref_flac_enc_probe.c"), "-I" + _REF_INC] + [l[0] for l in libs] + \
          ["-Wl,-rpath," + _LIBDIR, "-Wl,--allow-shlib-undefined", "-lm"]
    try:
        subprocess.check_call(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    except (subprocess.CalledProcessError, OSError):
        return None
    return _REF_ENC


def ref_encode(pcm, rate=44100, level=5, channels=1, bits=16):
    """REAL libavcodec FLAC encoder (LPC, 4096-sample frames) -> stream bytes, or None when the probe / libraries are absent.
    pcm: interleaved; bits 24 takes int32 samples with the value in the top 24 bits"""
    exe = build_ref_encoder()
    if exe is None or not os.path.isdir(_LIBDIR):
        return None
    with tempfile.TemporaryDirectory() as d:
        fi, fo = os.path.join(d, "a.raw"), os.path.join(d, "a.flac")
        np.ascontiguousarray(pcm, dtype=np.int16 if bits <= 16 else np.int32).tofile(fi)
        env = dict(os.environ, LD_LIBRARY_PATH=_LIBDIR + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
        r = subprocess.run([exe, fi, fo, str(rate), str(level), str(channels), str(bits)], capture_output=True, text=True, env=env)
        if r.returncode != 0:
            raise RuntimeError(f"ref_flac_encode rc={r.returncode}: {r.stdout} {r.stderr}")
        return open(fo, "rb").read()


_REF_WAV = os.path.join(_HERE, "_ref", "ref_wav_read")


def ref_wav_read(data):
    """REAL libavformat + libavcodec read of a file image (WAV, FLAC, ...: the demuxer is probed from the content, as the
    reference's audio.Reader does) -> (interleaved numpy samples, rate, channels) or None"""
    if not os.path.exists(_REF_WAV):
        libs = [sorted(glob.glob(os.path.join(_LIBDIR, p))) for p in ("libavformat-*", "libavcodec-*", "libavutil-*", "libswresample-*")]
        if not os.path.isdir(_REF_INC) or not all(libs):
            return None
        os.makedirs(os.path.dirname(_REF_WAV), exist_ok=True)
        cmd = ["gcc", "-O1", "-o", _REF_WAV, os.path.join(_HERE, "ref_wav_probe.c"), "-I" + _REF_INC] + [l[0] for l in libs] + \
              ["-Wl,-rpath," + _LIBDIR, "-Wl,--allow-shlib-undefined", "-lm"]
        try:
            subprocess.check_call(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        except (subprocess.CalledProcessError, OSError):
            return None
    with tempfile.TemporaryDirectory() as d:
        fi, fo = os.path.join(d, "a.wav"), os.path.join(d, "a.raw")
        with open(fi, "wb") as f:
            f.write(data)
        env = dict(os.environ, LD_LIBRARY_PATH=_LIBDIR + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
        r = subprocess.run([_REF_WAV, fi, fo], capture_output=True, text=True, env=env)
        if r.returncode != 0:
            raise RuntimeError(f"ref_wav_read rc={r.returncode}: {r.stdout} {r.stderr}")
        info = dict(kv.split("=") for kv in r.stdout.split())
        dt = {1: np.int16, 2: np.int32, 3: np.float32, 4: np.float64}[int(info["fmt"])]
        return np.fromfile(fo, dtype=dt), int(info["rate"]), int(info["channels"])


def ref_decode(stream):
    """REAL libavcodec -> (int16 array, rate, channels) or None when the probe / libraries are absent"""
    exe = build_ref()
    if exe is None or not os.path.isdir(_LIBDIR):
        return None
    with tempfile.TemporaryDirectory() as d:
        fi, fo = os.path.join(d, "a.flac"), os.path.join(d, "a.raw")
        with open(fi, "wb") as f:
            f.write(stream)
        env = dict(os.environ, LD_LIBRARY_PATH=_LIBDIR + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
        r = subprocess.run([exe, fi, fo], capture_output=True, text=True, env=env)
        if r.returncode != 0:
            raise RuntimeError(f"ref_flac_decode rc={r.returncode}: {r.stdout} {r.stderr}")
        info = dict(kv.split("=") for kv in r.stdout.split())
        return np.fromfile(fo, dtype=np.int16), int(info["rate"]), int(info["channels"])
