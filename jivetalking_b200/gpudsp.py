"""ctypes binding to libjtdsp.so (include/jtdsp.h) -- the B200 CUDA replacement for the
FFmpeg filter graphs jivetalking drives through setupFilterGraph / runFilterGraph
(reference: internal/processor/frame_processor.go:64-216).

This module is plumbing for tests and bench.py; the product is the shared library.  It
fails loudly when the library is missing or no CUDA device is present: there is no CPU
path (and it never imports oracle/).
"""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libjtdsp.so")

FMT_S16, FMT_S32, FMT_FLT, FMT_DBL = 1, 2, 3, 4
_NP_OF_FMT = {FMT_S16: np.int16, FMT_S32: np.int32, FMT_FLT: np.float32, FMT_DBL: np.float64}
_FMT_OF_NP = {np.dtype(np.int16): FMT_S16, np.dtype(np.int32): FMT_S32, np.dtype(np.float32): FMT_FLT, np.dtype(np.float64): FMT_DBL}

AS_NAMES = ["Dynamic_range", "RMS_level", "Peak_level", "RMS_trough", "RMS_peak", "DC_offset", "Flat_factor",
            "Crest_factor", "Zero_crossings_rate", "Zero_crossings", "Max_difference", "Min_difference",
            "Mean_difference", "RMS_difference", "Entropy", "Min_level", "Max_level", "Noise_floor",
            "Noise_floor_count", "Bit_depth", "Number_of_samples"]
SP_NAMES = ["mean", "variance", "centroid", "spread", "skewness", "kurtosis", "entropy", "flatness", "crest",
            "flux", "slope", "decrease", "rolloff"]
AS_COUNT, SP_COUNT = len(AS_NAMES), len(SP_NAMES)

ERRORS = {0: "JT_OK", -1: "JT_ERR_INVALID_ARG", -2: "JT_ERR_CUDA", -3: "JT_ERR_NOMEM", -4: "JT_ERR_SPEC",
          -5: "JT_ERR_UNSUPPORTED", -6: "JT_ERR_CANCELLED", -7: "JT_ERR_BUFFER"}


class JtError(RuntimeError):
    def __init__(self, code, detail=""):
        self.code = code
        super().__init__(f"{ERRORS.get(code, code)}: {detail}")


class FrameMeta(C.Structure):
    _fields_ = [("first_sample", C.c_int64), ("nb_samples", C.c_int32), ("reserved", C.c_int32),
                ("r128_M", C.c_double), ("r128_S", C.c_double), ("r128_I", C.c_double), ("r128_LRA", C.c_double),
                ("r128_LRA_low", C.c_double), ("r128_LRA_high", C.c_double),
                ("r128_true_peak", C.c_double), ("r128_sample_peak", C.c_double),
                ("astats", C.c_double * AS_COUNT),
                ("astats_overall_RMS_level", C.c_double), ("astats_overall_Peak_level", C.c_double),
                ("spectral", C.c_double * SP_COUNT)]


class LoudnormStats(C.Structure):
    _fields_ = [("input_i", C.c_double), ("input_tp", C.c_double), ("input_lra", C.c_double), ("input_thresh", C.c_double),
                ("output_i", C.c_double), ("output_tp", C.c_double), ("output_lra", C.c_double), ("output_thresh", C.c_double),
                ("target_offset", C.c_double), ("normalization_type", C.c_int32), ("valid", C.c_int32)]

    def as_dict(self):
        return {f: getattr(self, f) for f, _ in self._fields_}


class Interval(C.Structure):
    _fields_ = [("timestamp_s", C.c_double), ("rms_level", C.c_double), ("peak_level", C.c_double),
                ("spectral", C.c_double * SP_COUNT), ("spectral_found", C.c_int32), ("frame_count", C.c_int32),
                ("momentary_lufs", C.c_double), ("short_term_lufs", C.c_double), ("true_peak", C.c_double),
                ("sample_peak", C.c_double)]


class Measurements(C.Structure):
    _fields_ = [("input_i", C.c_double), ("input_tp", C.c_double), ("input_sp", C.c_double), ("input_lra", C.c_double),
                ("last_m", C.c_double), ("last_s", C.c_double), ("astats", C.c_double * AS_COUNT),
                ("spectral_mean", C.c_double * SP_COUNT), ("spectral_frames", C.c_int64), ("sink_frames", C.c_int64),
                ("duration_s", C.c_double)]

    def as_dict(self):
        d = {k: getattr(self, k) for k in ("input_i", "input_tp", "input_sp", "input_lra", "last_m", "last_s",
                                          "spectral_frames", "sink_frames", "duration_s")}
        d["astats"] = {n: self.astats[i] for i, n in enumerate(AS_NAMES)}
        d["spectral_mean"] = {n: self.spectral_mean[i] for i, n in enumerate(SP_NAMES)}
        return d


class ProcessResult(C.Structure):
    _fields_ = [("input", Measurements), ("filtered", Measurements), ("final", Measurements),
                ("pass3", LoudnormStats), ("pass4", LoudnormStats),
                ("limiter_ceiling_db", C.c_double), ("limiter_pregain_db", C.c_double), ("gain_db", C.c_double),
                ("effective_target_i", C.c_double),
                ("limiter_needed", C.c_int32), ("limiter_clamped", C.c_int32), ("linear_possible", C.c_int32),
                ("reserved", C.c_int32), ("n_out", C.c_int64)]


_lib = None
_P, _I64, _INT = C.c_void_p, C.c_int64, C.c_int


def lib():
    """Load libjtdsp.so; raises (never falls back) when it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    L.jt_version.restype = _INT
    L.jt_create.argtypes = [_INT, C.POINTER(_P)]
    L.jt_destroy.argtypes = [_P]
    L.jt_strerror.restype = C.c_char_p
    L.jt_strerror.argtypes = [_INT]
    L.jt_last_error.restype = C.c_char_p
    L.jt_last_error.argtypes = [_P]
    L.jt_cancel.argtypes = [_P]
    g = [_P, C.c_char_p, _P, _I64, _INT, _INT, _INT, _INT, _P, _I64, C.POINTER(_I64), C.POINTER(_INT), C.POINTER(_INT),
         C.POINTER(FrameMeta), _I64, C.POINTER(_I64), C.POINTER(LoudnormStats)]
    L.jt_run_graph.argtypes = g
    L.jt_run_graph_dev.argtypes = g
    L.jt_graph_max_out_frames.restype = _I64
    L.jt_graph_max_out_frames.argtypes = [C.c_char_p, _I64, _INT]
    L.jt_graph_max_meta.restype = _I64
    L.jt_graph_max_meta.argtypes = [C.c_char_p, _I64, _INT, _INT]
    L.jt_analyse.argtypes = [_P, _P, _I64, _INT, _INT, _INT, _INT, C.POINTER(Measurements), C.POINTER(Interval), _I64,
                             C.POINTER(_I64)]
    L.jt_analyse_chunk_unit.restype = _I64
    L.jt_analyse_chunk_unit.argtypes = [_INT]
    L.jt_analyse_chunk_bytes.restype = _I64
    L.jt_analyse_chunk_bytes.argtypes = [_I64, _INT]
    L.jt_analyse_chunk.argtypes = [_P, _P, _I64, _INT, _INT, _INT, _I64, _I64, _I64, _I64, _P, _I64, C.POINTER(_I64)]
    L.jt_analyse_merge.argtypes = [_INT, C.POINTER(_P), C.POINTER(Measurements), C.POINTER(Interval), _I64, C.POINTER(_I64)]
    L.jt_set_exchange.argtypes = [_P, _P, _P, _INT]
    L.jt_graph_exchanges.restype = _INT
    L.jt_graph_exchanges.argtypes = [C.c_char_p]
    L.jt_graph_chunk_unit.restype = _I64
    L.jt_graph_chunk_unit.argtypes = [C.c_char_p, _INT]
    L.jt_graph_chunk_context.argtypes = [C.c_char_p, _INT, C.POINTER(_I64), C.POINTER(_I64)]
    L.jt_graph_chunk_bytes.restype = _I64
    L.jt_graph_chunk_bytes.argtypes = [C.c_char_p, _I64, _INT]
    L.jt_graph_chunk.argtypes = [_P, C.c_char_p, _P, _I64, _INT, _INT, _INT, _I64, _I64, _I64, _I64, _INT,
                                 _P, _I64, C.POINTER(_I64), C.POINTER(_I64), C.POINTER(_INT), C.POINTER(_INT),
                                 _P, _I64, C.POINTER(_I64)]
    L.jt_graph_merge.argtypes = [C.c_char_p, _I64, _INT, _INT, _INT, _INT, _INT, C.POINTER(_P),
                                 C.POINTER(FrameMeta), _I64, C.POINTER(_I64), C.POINTER(LoudnormStats),
                                 C.POINTER(Measurements)]
    L.jt_band_rms.argtypes = [_P, _P, _I64, _INT, _INT, _INT, C.c_double, C.c_double, _P, _P, _INT, _P, _P]
    pa = [_P, _P, _I64, _INT, _INT, _INT, C.c_char_p, _P, _I64, C.POINTER(ProcessResult)]
    L.jt_process_audio.argtypes = pa
    L.jt_process_audio_dev.argtypes = pa
    L.jt_build_pass3_spec.argtypes = [C.c_double] * 5 + [C.c_char_p, C.c_size_t, C.POINTER(ProcessResult)]
    L.jt_build_pass4_spec.argtypes = [C.POINTER(ProcessResult), C.POINTER(LoudnormStats), C.c_double, C.c_double,
                                      C.c_double, _INT, C.c_char_p, C.c_size_t, C.POINTER(C.c_double),
                                      C.POINTER(C.c_double)]
    L.jt_default_pass2_spec.argtypes = [C.c_char_p, C.c_size_t]
    L.jt_pass1_spec.argtypes = [C.c_char_p, C.c_size_t]
    L.jt_loudnorm_stats_json.argtypes = [C.POINTER(LoudnormStats), C.c_char_p, C.c_size_t]
    L.jt_wav_parse.argtypes = [_P, _I64, C.POINTER(_INT), C.POINTER(_INT), C.POINTER(_INT), C.POINTER(_I64), C.POINTER(_I64)]
    L.jt_flac_max_bytes.restype = _I64
    L.jt_flac_max_bytes.argtypes = [_I64, _INT]
    fl = [_P, _P, _I64, _INT, _INT, _P, _I64, C.POINTER(_I64)]
    L.jt_flac_encode.argtypes = fl
    L.jt_flac_encode_dev.argtypes = fl
    L.jt_flac_stream_info.argtypes = [_P, _I64, C.POINTER(_INT), C.POINTER(_INT), C.POINTER(_INT), C.POINTER(_INT), C.POINTER(_I64), C.POINTER(_I64)]
    dec = [_P, _P, _I64, _P, _I64, C.POINTER(_I64), C.POINTER(_INT), C.POINTER(_INT), C.POINTER(_INT)]
    for name in ("jt_flac_decode", "jt_flac_decode_dev", "jt_wav_decode", "jt_wav_decode_dev"):
        getattr(L, name).argtypes = dec
    L.jt_prefetch_input.argtypes = [_P, _P, _I64, _INT, _INT]
    L.jt_md5.argtypes = [_P, _I64, _P]
    L.jt_flac_set_md5.argtypes = [_P, _I64, _P, _I64]
    L.jt_cuda_stream.restype = _P
    L.jt_cuda_stream.argtypes = [_P]
    L.jt_launch_count.restype = _I64
    L.jt_launch_count.argtypes = [_P]
    L.jt_reset_launch_count.argtypes = [_P]
    L.jt_enable_kernel_timing.argtypes = [_P, _INT]
    L.jt_kernel_timing.restype = C.c_char_p
    L.jt_kernel_timing.argtypes = [_P, _INT, C.POINTER(C.c_double), C.POINTER(_I64)]
    _lib = L
    return L


def wav_parse(data):
    """(numpy view of the interleaved samples, rate, channels) of a RIFF/WAVE file image (bytes); raises JtError."""
    buf = np.frombuffer(data, dtype=np.uint8)
    fmt, rate, ch, off, nfr = _INT(0), _INT(0), _INT(0), _I64(0), _I64(0)
    rc = lib().jt_wav_parse(buf.ctypes.data_as(_P), len(buf), C.byref(fmt), C.byref(rate), C.byref(ch), C.byref(off), C.byref(nfr))
    if rc != 0:
        raise JtError(rc, "jt_wav_parse")
    dt = _NP_OF_FMT[fmt.value]
    n = nfr.value * ch.value
    return np.frombuffer(data, dtype=dt, count=n, offset=off.value), rate.value, ch.value


def flac_stream_info(data):
    """STREAMINFO of a FLAC file image: dict(fmt, rate, channels, bits, n_frames, audio_offset); host only."""
    buf = np.frombuffer(data, dtype=np.uint8)
    fmt, rate, ch, bits, nfr, off = _INT(0), _INT(0), _INT(0), _INT(0), _I64(0), _I64(0)
    rc = lib().jt_flac_stream_info(buf.ctypes.data_as(_P), len(buf), C.byref(fmt), C.byref(rate), C.byref(ch), C.byref(bits), C.byref(nfr), C.byref(off))
    if rc != 0:
        raise JtError(rc, "jt_flac_stream_info")
    return dict(fmt=fmt.value, rate=rate.value, channels=ch.value, bits=bits.value, n_frames=nfr.value, audio_offset=off.value)


def flac_set_md5(stream, pcm_s16):
    """stream bytes with STREAMINFO's MD5 filled in from the mono s16 samples it encodes (host only)"""
    buf = bytearray(stream)
    pcm = np.ascontiguousarray(pcm_s16, dtype=np.int16)
    rc = lib().jt_flac_set_md5((C.c_char * len(buf)).from_buffer(buf), len(buf), pcm.ctypes.data_as(_P), len(pcm))
    if rc != 0:
        raise JtError(rc, "jt_flac_set_md5")
    return bytes(buf)


def pass1_spec():
    b = C.create_string_buffer(4096)
    lib().jt_pass1_spec(b, len(b))
    return b.value.decode()


def default_pass2_spec():
    b = C.create_string_buffer(4096)
    lib().jt_default_pass2_spec(b, len(b))
    return b.value.decode()


def meta_to_dict(m):
    d = {k: getattr(m, k) for k in ("first_sample", "nb_samples", "r128_M", "r128_S", "r128_I", "r128_LRA",
                                    "r128_LRA_low", "r128_LRA_high", "r128_true_peak", "r128_sample_peak",
                                    "astats_overall_RMS_level", "astats_overall_Peak_level")}
    d["astats"] = {n: m.astats[i] for i, n in enumerate(AS_NAMES)}
    d["spectral"] = {n: m.spectral[i] for i, n in enumerate(SP_NAMES)}
    return d


class Context:
    """One jt_ctx (one CUDA stream).  Mirrors a per-worker config clone (filters.go:368-373)."""

    def __init__(self, device=0):
        self._h = _P(None)
        rc = lib().jt_create(device, C.byref(self._h))
        if rc != 0:
            raise JtError(rc, lib().jt_strerror(rc).decode() + " (no CPU fallback)")

    def close(self):
        if self._h:
            lib().jt_destroy(self._h)
            self._h = _P(None)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc):
        if rc != 0:
            raise JtError(rc, lib().jt_last_error(self._h).decode())

    # -- S4 -------------------------------------------------------------------------------
    def run_graph(self, spec, pcm, rate, channels=1, frame_size=4096, want_pcm=True, want_meta=True):
        """pcm: numpy array (int16 / float32 / float64), interleaved if channels > 1.
        Returns dict(pcm, rate, fmt, meta=[FrameMeta...], loudnorm=LoudnormStats)."""
        pcm = np.ascontiguousarray(pcm)
        fmt = _FMT_OF_NP[pcm.dtype]
        n = pcm.size // channels
        bspec = spec.encode()
        cap = lib().jt_graph_max_out_frames(bspec, n, rate)
        mcap = lib().jt_graph_max_meta(bspec, n, rate, frame_size)
        out = np.zeros(cap, dtype=np.float64) if want_pcm else None     # 8 bytes/sample holds any format
        meta = (FrameMeta * mcap)() if want_meta else None
        n_out, n_meta, orate, ofmt = _I64(0), _I64(0), _INT(0), _INT(0)
        ln = LoudnormStats()
        rc = lib().jt_run_graph(self._h, bspec, pcm.ctypes.data_as(_P), n, rate, channels, fmt, frame_size,
                                out.ctypes.data_as(_P) if want_pcm else None, cap, C.byref(n_out), C.byref(orate),
                                C.byref(ofmt), meta, mcap, C.byref(n_meta), C.byref(ln))
        self._check(rc)
        res = dict(rate=orate.value, fmt=ofmt.value, n_out=n_out.value, loudnorm=ln, pcm=None, meta=[])
        if want_pcm:
            dt = _NP_OF_FMT[ofmt.value]
            res["pcm"] = out.view(np.uint8)[: n_out.value * np.dtype(dt).itemsize].view(dt).copy()
        if want_meta:
            res["meta"] = [meta[i] for i in range(n_meta.value)]
        return res

    def analyse(self, pcm, rate, channels=1, frame_size=4096):
        pcm = np.ascontiguousarray(pcm)
        fmt = _FMT_OF_NP[pcm.dtype]
        n = pcm.size // channels
        cap = int(n / rate / 0.25) + 8
        iv = (Interval * cap)()
        n_iv = _I64(0)
        m = Measurements()
        rc = lib().jt_analyse(self._h, pcm.ctypes.data_as(_P), n, rate, channels, fmt, frame_size, C.byref(m), iv, cap,
                              C.byref(n_iv))
        self._check(rc)
        return m, [iv[i] for i in range(n_iv.value)]

    def analyse_chunk(self, pcm_local, rate, channels, local_first, own_first, owned, total_frames):
        """One chunk of a long stream (include/jtdsp.h: jt_analyse_chunk).  Returns the mergeable blob (bytes)."""
        pcm_local = np.ascontiguousarray(pcm_local)
        fmt = _FMT_OF_NP[pcm_local.dtype]
        cap = lib().jt_analyse_chunk_bytes(owned, rate)
        buf = C.create_string_buffer(cap)
        n = _I64(0)
        rc = lib().jt_analyse_chunk(self._h, pcm_local.ctypes.data_as(_P), pcm_local.size // channels, rate, channels, fmt,
                                    local_first, own_first, owned, total_frames, buf, cap, C.byref(n))
        self._check(rc)
        return buf.raw[: n.value]

    def set_exchange(self, fn, n_ranks, raw=False):
        """fn(send: bytes) -> bytes of n_ranks records in rank order (an all-gather); None removes it.  The carries
        of a sharded stream (afftdn's tracked noise floor) cross chunk boundaries through it (jt_set_exchange).
        raw=True: fn(send_addr, nbytes, recv_addr) -> None works on the library's buffers in place (no Python copies)."""
        if fn is None:
            self._xfn = None
            lib().jt_set_exchange(self._h, None, None, 1)
            return
        if raw:
            def _cbr(user, send, nbytes, recv):
                try:
                    fn(send, nbytes, recv)
                    return 0
                except Exception:          # an exception must not unwind through the C frames
                    import traceback
                    traceback.print_exc()
                    return -1
            self._xfn = EXCHANGE_FN(_cbr)
            lib().jt_set_exchange(self._h, C.cast(self._xfn, _P), None, n_ranks)
            return

        def _cb(user, send, nbytes, recv):
            try:
                got = fn(C.string_at(send, nbytes))
                if len(got) != nbytes * n_ranks:
                    return -1
                C.memmove(recv, got, len(got))
                return 0
            except Exception:          # an exception must not unwind through the C frames
                import traceback
                traceback.print_exc()
                return -1

        self._xfn = EXCHANGE_FN(_cb)          # keep the trampoline alive as long as the context uses it
        lib().jt_set_exchange(self._h, C.cast(self._xfn, _P), None, n_ranks)

    def _pinned_bytes(self, nbytes):
        """Grow-only pinned host buffer owned by this context (torch provides the allocation): large results leave the
        GPU at PCIe rate and no call zero-fills gigabytes of pageable memory."""
        import torch
        buf = getattr(self, "_pin", None)
        if buf is None or buf.numel() < nbytes:
            self._pin = buf = torch.empty(int(nbytes * 1.05) + 4096, dtype=torch.uint8, pin_memory=True)
        return buf.numpy()

    def graph_chunk(self, spec, pcm_local, rate, channels, local_first, own_first, owned, total_frames,
                    frame_size=4096, want_pcm=True, want_blob=True, pinned=False):
        """Any graph on a window of a longer stream (include/jtdsp.h: jt_graph_chunk).  Returns
        dict(pcm=owned sink samples or None, out_first, n_out, rate, fmt, blob=bytes).  pinned=True: `pcm` is a view
        into the context's pinned result buffer, valid until the next call with pinned=True."""
        pcm_local = np.ascontiguousarray(pcm_local)
        fmt = _FMT_OF_NP[pcm_local.dtype]
        bspec = spec.encode()
        n_local = pcm_local.size // channels
        cap = lib().jt_graph_max_out_frames(bspec, owned, rate) + 4096
        out = None
        if want_pcm:
            out = self._pinned_bytes(cap * 8) if pinned else np.empty(cap * 8, dtype=np.uint8)     # 8 bytes/sample holds any format
        bcap = lib().jt_graph_chunk_bytes(bspec, owned, rate)
        buf = C.create_string_buffer(bcap) if want_blob else None
        o_first, n_out, orate, ofmt, nb = _I64(0), _I64(0), _INT(0), _INT(0), _I64(0)
        rc = lib().jt_graph_chunk(self._h, bspec, pcm_local.ctypes.data_as(_P), n_local, rate, channels, fmt,
                                  local_first, own_first, owned, total_frames, frame_size,
                                  out.ctypes.data_as(_P) if want_pcm else None, cap, C.byref(o_first), C.byref(n_out),
                                  C.byref(orate), C.byref(ofmt), buf, bcap, C.byref(nb))
        self._check(rc)
        res = dict(out_first=o_first.value, n_out=n_out.value, rate=orate.value, fmt=ofmt.value, pcm=None,
                   blob=buf.raw[: nb.value] if want_blob else b"")
        if want_pcm:
            dt = _NP_OF_FMT[ofmt.value]
            res["pcm"] = out[: n_out.value * np.dtype(dt).itemsize].view(dt)
        return res

    def band_rms(self, pcm, rate, start_s, duration_s, lo_hz, hi_hz, channels=1):
        pcm = np.ascontiguousarray(pcm)
        fmt = _FMT_OF_NP[pcm.dtype]
        lo = np.ascontiguousarray(lo_hz, dtype=np.float64)
        hi = np.ascontiguousarray(hi_hz, dtype=np.float64)
        out = np.zeros(len(lo))
        found = np.zeros(len(lo), dtype=np.int32)
        rc = lib().jt_band_rms(self._h, pcm.ctypes.data_as(_P), pcm.size // channels, rate, channels, fmt, start_s,
                               duration_s, lo.ctypes.data_as(_P), hi.ctypes.data_as(_P), len(lo),
                               out.ctypes.data_as(_P), found.ctypes.data_as(_P))
        self._check(rc)
        return out, found

    def process_audio(self, pcm, rate, channels=1, pass2_spec=None):
        """Full four-pass chain.  Returns (int16 mono 44.1 kHz array, ProcessResult)."""
        pcm = np.ascontiguousarray(pcm)
        fmt = _FMT_OF_NP[pcm.dtype]
        n = pcm.size // channels
        cap = int(n * 44100 / rate) + 3 * 4096
        out = np.zeros(cap, dtype=np.int16)
        res = ProcessResult()
        rc = lib().jt_process_audio(self._h, pcm.ctypes.data_as(_P), n, rate, channels, fmt,
                                    pass2_spec.encode() if pass2_spec else None, out.ctypes.data_as(_P), cap,
                                    C.byref(res))
        self._check(rc)
        return out[: res.n_out].copy(), res

    def flac_encode(self, pcm_s16, rate=44100, block_size=4096):
        """Complete FLAC stream (bytes) of mono s16 samples: the container of the chain's output (encoder.go:92-101)."""
        pcm = np.ascontiguousarray(pcm_s16, dtype=np.int16)
        cap = lib().jt_flac_max_bytes(len(pcm), block_size)
        out = np.zeros(cap, dtype=np.uint8)
        nb = _I64(0)
        self._check(lib().jt_flac_encode(self._h, pcm.ctypes.data_as(_P), len(pcm), rate, block_size, out.ctypes.data_as(_P), cap, C.byref(nb)))
        return out[: nb.value].tobytes()

    def _decode(self, fn, data, cap_frames, max_channels=8):
        buf = np.frombuffer(data, dtype=np.uint8)
        out = np.zeros(max(1, cap_frames * max_channels), dtype=np.int64)          # room for any format
        nfr, fmt, rate, ch = _I64(0), _INT(0), _INT(0), _INT(0)
        self._check(fn(self._h, buf.ctypes.data_as(_P), len(buf), out.ctypes.data_as(_P), cap_frames, C.byref(nfr), C.byref(fmt), C.byref(rate), C.byref(ch)))
        pcm = np.frombuffer(out.tobytes(), dtype=_NP_OF_FMT[fmt.value], count=nfr.value * ch.value).copy()
        return pcm, fmt.value, rate.value, ch.value

    def flac_decode(self, data, cap_frames=None):
        """FLAC file image (bytes) -> (interleaved samples as libavcodec gives them, JT_FMT_*, rate, channels); reader.go:29-188."""
        if cap_frames is None:
            info = flac_stream_info(data)
            cap_frames = info["n_frames"] if info["n_frames"] else len(data) * 8
            return self._decode(lib().jt_flac_decode, data, cap_frames, info["channels"])
        return self._decode(lib().jt_flac_decode, data, cap_frames)

    def wav_decode(self, data):
        """RIFF/WAVE file image -> (interleaved samples, JT_FMT_*, rate, channels), 24-bit PCM included (as s32 << 8)."""
        return self._decode(lib().jt_wav_decode, data, len(data), 1)

    def decode_ptr(self, kind, in_ptr, n_bytes, out_ptr, cap_frames, on_device):
        fn = getattr(lib(), "jt_%s_decode%s" % (kind, "_dev" if on_device else ""))
        nfr, fmt, rate, ch = _I64(0), _INT(0), _INT(0), _INT(0)
        self._check(fn(self._h, _P(in_ptr), n_bytes, _P(out_ptr), cap_frames, C.byref(nfr), C.byref(fmt), C.byref(rate), C.byref(ch)))
        return nfr.value, fmt.value, rate.value, ch.value

    def flac_encode_ptr(self, in_ptr, n, rate, block_size, out_ptr, out_cap, on_device):
        nb = _I64(0)
        fn = lib().jt_flac_encode_dev if on_device else lib().jt_flac_encode
        self._check(fn(self._h, _P(in_ptr), n, rate, block_size, _P(out_ptr), out_cap, C.byref(nb)))
        return nb.value

    # -- raw pointer variants for bench.py (torch owns the memory) ----------------------------
    def process_audio_ptr(self, in_ptr, n, rate, channels, fmt, out_ptr, out_cap, on_device, pass2_spec=None):
        res = ProcessResult()
        fn = lib().jt_process_audio_dev if on_device else lib().jt_process_audio
        rc = fn(self._h, _P(in_ptr), n, rate, channels, fmt, pass2_spec.encode() if pass2_spec else None, _P(out_ptr),
                out_cap, C.byref(res))
        self._check(rc)
        return res

    def cuda_stream(self):
        """cudaStream_t the context launches on (for event timing / torch.cuda.ExternalStream)."""
        return lib().jt_cuda_stream(self._h)

    def launch_count(self):
        return lib().jt_launch_count(self._h)

    def prefetch_input_ptr(self, in_ptr, n_frames, channels, fmt):
        """start the upload of a LATER call's (pinned) host input; the call that gets the same pointer finds it resident"""
        self._check(lib().jt_prefetch_input(self._h, _P(in_ptr), n_frames, channels, fmt))

    def reset_counters(self):
        lib().jt_reset_launch_count(self._h)

    def enable_timing(self, on=True):
        lib().jt_enable_kernel_timing(self._h, 1 if on else 0)

    def kernel_timings(self):
        out, i = [], 0
        while True:
            ms, ln = C.c_double(0), _I64(0)
            name = lib().jt_kernel_timing(self._h, i, C.byref(ms), C.byref(ln))
            if not name:
                break
            out.append((name.decode(), ms.value, ln.value))
            i += 1
        return out


def analyse_chunk_unit(rate):
    return lib().jt_analyse_chunk_unit(rate)


def analyse_merge(blobs, total_frames, rate):
    """Host-only merge of jt_analyse_chunk blobs -> (Measurements, [Interval]) of the whole stream."""
    keep = [C.create_string_buffer(b, len(b)) for b in blobs]
    arr = (_P * len(keep))(*[C.cast(k, _P) for k in keep])
    cap = int(total_frames / rate / 0.25) + 8
    iv = (Interval * cap)()
    n_iv = _I64(0)
    m = Measurements()
    rc = lib().jt_analyse_merge(len(keep), arr, C.byref(m), iv, cap, C.byref(n_iv))
    if rc:
        raise JtError(rc, lib().jt_strerror(rc).decode())
    return m, [iv[i] for i in range(n_iv.value)]


EXCHANGE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p)


def graph_chunk_unit(spec, rate):
    return lib().jt_graph_chunk_unit(spec.encode(), rate)


def graph_exchanges(spec):
    return lib().jt_graph_exchanges(spec.encode())


def graph_chunk_context(spec, rate):
    """Recommended (left, right) context of a mid-stream chunk, input frames."""
    le, ri = _I64(0), _I64(0)
    rc = lib().jt_graph_chunk_context(spec.encode(), rate, C.byref(le), C.byref(ri))
    if rc:
        raise JtError(rc, lib().jt_strerror(rc).decode())
    return le.value, ri.value


def graph_merge(spec, blobs, total_frames, rate, channels=1, fmt=FMT_FLT, frame_size=4096, want_meta=True):
    """Host-only merge of jt_graph_chunk blobs -> dict(meta=[FrameMeta], loudnorm, measurements) of the whole stream."""
    bspec = spec.encode()
    keep = [C.create_string_buffer(b, len(b)) for b in blobs]
    arr = (_P * len(keep))(*[C.cast(k, _P) for k in keep])
    mcap = lib().jt_graph_max_meta(bspec, total_frames, rate, frame_size) if want_meta else 0
    meta = (FrameMeta * mcap)() if want_meta else None
    n_meta = _I64(0)
    ln, m = LoudnormStats(), Measurements()
    rc = lib().jt_graph_merge(bspec, total_frames, rate, channels, fmt, frame_size, len(keep), arr, meta, mcap,
                              C.byref(n_meta), C.byref(ln), C.byref(m))
    if rc:
        raise JtError(rc, lib().jt_strerror(rc).decode())
    return dict(meta=[meta[i] for i in range(n_meta.value)] if want_meta else [], loudnorm=ln, measurements=m)


def build_pass3_spec(output_i, output_tp, target_i=-16.0, target_tp=-1.0, target_lra=20.0):
    b = C.create_string_buffer(4096)
    plan = ProcessResult()
    rc = lib().jt_build_pass3_spec(output_i, output_tp, target_i, target_tp, target_lra, b, len(b), C.byref(plan))
    if rc:
        raise JtError(rc)
    return b.value.decode(), plan


def build_pass4_spec(plan, pass3_stats, target_i=-16.0, target_tp=-1.0, target_lra=20.0, source_rate=44100):
    b = C.create_string_buffer(8192)
    eff, off = C.c_double(0), C.c_double(0)
    rc = lib().jt_build_pass4_spec(C.byref(plan), C.byref(pass3_stats), target_i, target_tp, target_lra, source_rate, b,
                                   len(b), C.byref(eff), C.byref(off))
    if rc:
        raise JtError(rc)
    return b.value.decode(), eff.value, off.value
