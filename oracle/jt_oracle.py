"""ctypes binding for the CPU oracle (oracle/_build/libjtoracle.so).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  The product package
(jivetalking_b200) must never import this module.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libjtoracle.so")

NSPEC = 13
SPEC_NAMES = ["mean", "variance", "centroid", "spread", "skewness", "kurtosis", "entropy",
              "flatness", "crest", "flux", "slope", "decrease", "rolloff"]
FMT_S16, FMT_S32, FMT_FLT, FMT_DBL = 1, 2, 3, 4


def build(force=False):
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".c", ".h"))]
    if force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


class R128Summary(C.Structure):
    _fields_ = [("n_ticks", C.c_int64), ("I", C.c_double), ("LRA", C.c_double), ("LRA_low", C.c_double),
                ("LRA_high", C.c_double), ("sample_peak", C.c_double), ("true_peak", C.c_double),
                ("rel_threshold_400", C.c_double)]


ASTATS_FIELDS = ["nb_samples", "DC_offset", "Min_level", "Max_level", "Min_difference", "Max_difference",
                 "Mean_difference", "RMS_difference", "Peak_level", "RMS_level", "RMS_peak", "RMS_trough",
                 "Crest_factor", "Flat_factor", "Peak_count", "Noise_floor", "Noise_floor_count", "Entropy",
                 "Bit_depth", "Dynamic_range", "Zero_crossings", "Zero_crossings_rate"]


class AstatsOut(C.Structure):
    _fields_ = [(f, C.c_double) for f in ASTATS_FIELDS]


_lib = None
_P = C.c_void_p
_I64 = C.c_int64


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.orc_swr_resample_f64.restype = _I64
        L.orc_swr_resample_f64.argtypes = [_P, _I64, C.c_int, C.c_int, C.c_int, _P, _I64]
        L.orc_swr_resample_f32.restype = _I64
        L.orc_swr_resample_f32.argtypes = [_P, _I64, C.c_int, C.c_int, C.c_int, _P, _I64]
        L.orc_swr_out_count.restype = _I64
        L.orc_swr_out_count.argtypes = [_I64, C.c_int, C.c_int]
        L.orc_swr_out_count_flush.restype = _I64
        L.orc_swr_out_count_flush.argtypes = [_I64, C.c_int, C.c_int]
        L.orc_ebur128.restype = C.c_int
        L.orc_ebur128.argtypes = [_P, _I64, C.c_int, C.c_int, C.c_int, _P, _P, _P, _P, _I64, C.POINTER(R128Summary)]
        L.orc_astats.restype = C.c_int
        L.orc_astats.argtypes = [_P, C.c_int, _I64, C.c_int, C.POINTER(AstatsOut)]
        L.orc_aspectralstats.restype = _I64
        L.orc_aspectralstats.argtypes = [_P, _I64, C.c_int, C.c_int, _P, _I64]
        for name in ("orc_conv_s16_to_f64", "orc_conv_s16_to_f32", "orc_conv_f64_to_s16", "orc_conv_f32_to_s16",
                     "orc_downmix_stereo_f32", "orc_downmix_stereo_s16"):
            getattr(L, name).restype = None
            getattr(L, name).argtypes = [_P, _I64, _P]
        _lib = L
    return _lib


def _ptr(a):
    return a.ctypes.data_as(_P)


def swr_resample(x, in_rate, out_rate, flush=True):
    x = np.ascontiguousarray(x)
    f = lib().orc_swr_resample_f64 if x.dtype == np.float64 else lib().orc_swr_resample_f32
    assert x.dtype in (np.float64, np.float32)
    cap = (lib().orc_swr_out_count_flush if flush else lib().orc_swr_out_count)(len(x), in_rate, out_rate)
    out = np.zeros(max(cap, 0), dtype=x.dtype)
    m = f(_ptr(x), len(x), in_rate, out_rate, int(flush), _ptr(out), len(out))
    assert m == len(out), (m, len(out))
    return out


def swr_out_count(n, in_rate, out_rate, flush=False):
    return (lib().orc_swr_out_count_flush if flush else lib().orc_swr_out_count)(n, in_rate, out_rate)


def ebur128(x, rate, dualmono=True, true_peak=True):
    x = np.ascontiguousarray(x, dtype=np.float64)
    cap = len(x) // max(rate // 10, 1) + 1
    M = np.zeros(cap); S = np.zeros(cap); sp = np.zeros(cap); tp = np.zeros(cap)
    s = R128Summary()
    r = lib().orc_ebur128(_ptr(x), len(x), rate, int(dualmono), int(true_peak), _ptr(M), _ptr(S), _ptr(sp), _ptr(tp),
                          cap, C.byref(s))
    assert r == 0
    k = s.n_ticks
    return dict(M=M[:k], S=S[:k], sample_peak_cum=sp[:k], true_peak_cum=tp[:k], I=s.I, LRA=s.LRA,
                LRA_low=s.LRA_low, LRA_high=s.LRA_high, sample_peak=s.sample_peak, true_peak=s.true_peak,
                n_ticks=k)


def astats(x, rate):
    x = np.ascontiguousarray(x)
    fmt = {np.dtype(np.int16): FMT_S16, np.dtype(np.int32): FMT_S32, np.dtype(np.float32): FMT_FLT, np.dtype(np.float64): FMT_DBL}[x.dtype]
    o = AstatsOut()
    assert lib().orc_astats(_ptr(x), fmt, len(x), rate, C.byref(o)) == 0
    return {f: getattr(o, f) for f in ASTATS_FIELDS}


def aspectralstats(x, rate, win_size=2048):
    x = np.ascontiguousarray(x, dtype=np.float32)
    hop = win_size // 2
    cap = (len(x) + hop - 1) // hop
    rows = np.zeros((max(cap, 1), NSPEC), dtype=np.float32)
    k = lib().orc_aspectralstats(_ptr(x), len(x), rate, win_size, _ptr(rows), cap)
    return rows[:k]


# ---- Pass-2 / Pass-4 filters ---------------------------------------------------------------
_D = C.c_double


def _proto(name, restype, argtypes):
    f = getattr(lib(), name)
    f.restype = restype
    f.argtypes = argtypes
    return f


def biquad(x, rate, kind, freq, q=0.707, normalize=False, tdii=False, mix=1.0):
    x = np.ascontiguousarray(x)
    c = np.zeros(5)
    _proto("orc_biquad_design", None, [C.c_int, _D, _D, C.c_int, C.c_int, _P])(1 if kind == "highpass" else 0, freq, q, rate, int(normalize), _ptr(c))
    name = {np.dtype(np.float32): "orc_biquad_f32", np.dtype(np.float64): "orc_biquad_f64", np.dtype(np.int16): "orc_biquad_s16",
            np.dtype(np.int32): "orc_biquad_s32"}[x.dtype]
    out = np.zeros_like(x)
    _proto(name, None, [_P, _P, _I64, _P, C.c_int, _D])(_ptr(x), _ptr(out), len(x), _ptr(c), int(tdii), mix)
    return out


def agate(x, rate, threshold, ratio, attack, release, range_, knee, makeup=1.0, detection_rms=True):
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.zeros_like(x)
    _proto("orc_agate", None, [_P, _P, _I64, C.c_int] + [_D] * 7 + [C.c_int])(_ptr(x), _ptr(out), len(x), rate, threshold, ratio, attack, release, range_, knee, makeup, int(detection_rms))
    return out


def acompressor(x, rate, threshold, ratio, attack, release, makeup, knee, mix=1.0, detection_rms=True):
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.zeros_like(x)
    _proto("orc_acompressor", None, [_P, _P, _I64, C.c_int] + [_D] * 7 + [C.c_int])(_ptr(x), _ptr(out), len(x), rate, threshold, ratio, attack, release, makeup, knee, mix, int(detection_rms))
    return out


def deesser(x, rate, i, m=0.5, f=0.5):
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.zeros_like(x)
    _proto("orc_deesser", None, [_P, _P, _I64, C.c_int, _D, _D, _D])(_ptr(x), _ptr(out), len(x), rate, i, m, f)
    return out


def volume_f32(x, volume):
    x = np.ascontiguousarray(x, dtype=np.float32)
    out = np.zeros_like(x)
    _proto("orc_volume_f32", None, [_P, _P, _I64, _D])(_ptr(x), _ptr(out), len(x), volume)
    return out


def anlmdn(x, rate, s=0.00001, p=0.002, r=0.006, m=11.0):
    x = np.ascontiguousarray(x, dtype=np.float32)
    out = np.zeros_like(x)
    _proto("orc_anlmdn", C.c_int, [_P, _P, _I64, C.c_int, _D, _D, _D, _D])(_ptr(x), _ptr(out), len(x), rate, s, p, r, m)
    return out


def alimiter(x, rate, limit, attack=5.0, release=50.0, level_in=1.0, level_out=1.0, level=True, asc=False, asc_level=0.5):
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.zeros_like(x)
    r = _proto("orc_alimiter", C.c_int, [_P, _P, _I64, C.c_int, _D, _D, _D, _D, _D, C.c_int, C.c_int, _D])(
        _ptr(x), _ptr(out), len(x), rate, limit, attack, release, level_in, level_out, int(level), int(asc), asc_level)
    assert r >= 0
    return out


def adeclick(x, rate, w=55.0, o=75.0, a=2.0, t=2.0, b=2.0, method_save=True):
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.zeros_like(x)
    det = _I64(0)
    r = _proto("orc_adeclick", _I64, [_P, _P, _I64, C.c_int, _D, _D, _D, _D, _D, C.c_int, C.POINTER(_I64)])(
        _ptr(x), _ptr(out), len(x), rate, w, o, a, t, b, int(method_save), C.byref(det))
    assert r == len(x), r
    return out, det.value


def loudnorm_meter(x, rate, dual_mono=True):
    x = np.ascontiguousarray(x, dtype=np.float64)
    o = np.zeros(4)
    _proto("orc_loudnorm_meter", C.c_int, [_P, _I64, C.c_int, C.c_int, _P])(_ptr(x), len(x), rate, int(dual_mono), _ptr(o))
    return dict(I=o[0], LRA=o[1], thresh=o[2], sample_peak=o[3])


def loudnorm_is_linear(I, TP, LRA, mI, mTP, mLRA, mTh, linear=True):
    opt = np.array([I, TP, LRA, mI, mTP, mLRA, mTh, 0.0])
    return bool(_proto("orc_loudnorm_is_linear", C.c_int, [_P, C.c_int])(_ptr(opt), int(linear)))


def loudnorm(x, rate, I=-24.0, TP=-2.0, LRA=7.0, mI=0.0, mTP=99.0, mLRA=0.0, mTh=-70.0, offset=0.0, linear=True, dual_mono=False):
    """af_loudnorm.c over its input link x (f64; 192 kHz when the mode is dynamic).  Returns (y, stats dict)."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    opt = np.array([I, TP, LRA, mI, mTP, mLRA, mTh, offset])
    y = np.zeros(len(x))
    st = np.zeros(10)
    n = _proto("orc_loudnorm", _I64, [_P, _I64, C.c_int, _P, C.c_int, C.c_int, _P, _P])(_ptr(x), len(x), rate, _ptr(opt), int(linear), int(dual_mono), _ptr(y), _ptr(st))
    assert n == len(x), (n, len(x))
    keys = ["input_i", "input_tp", "input_lra", "input_thresh", "output_i", "output_tp", "output_lra", "output_thresh"]
    d = {k: float(st[i]) for i, k in enumerate(keys)}
    d["normalization_type"] = int(st[8]); d["target_offset"] = float(st[9])
    return y, d


def afftdn(x, rate, nr=12.0, nf=-50.0, nt=0, bn=None, tn=False, ad=0.5, fo=1.0, bm=1.25):
    x = np.ascontiguousarray(x, dtype=np.float32)
    out = np.zeros_like(x)
    bnp = None
    if bn is not None:
        bna = np.zeros(15)
        bna[:len(bn)] = bn[:15]
        bnp = _ptr(bna)
    _proto("orc_afftdn", C.c_int, [_P, _P, _I64, C.c_int, _D, _D, C.c_int, _P, C.c_int, _D, _D, _D])(_ptr(x), _ptr(out), len(x), rate, nr, nf, nt, bnp, int(tn), ad, fo, bm)
    return out
