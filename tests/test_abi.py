"""The C-ABI library loads and exports every symbol include/jtdsp.h declares (no compute)."""
import ctypes as C
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "jtdsp.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(jt_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    from jivetalking_b200 import gpudsp
    lib = C.CDLL(gpudsp.LIB_PATH)
    syms = declared_symbols()
    assert len(syms) >= 20
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing


def test_error_strings_and_specs():
    from jivetalking_b200 import gpudsp
    L = gpudsp.lib()
    assert L.jt_version() >= 100
    for code in range(0, -8, -1):
        assert L.jt_strerror(code)
    p1 = gpudsp.pass1_spec()
    # Pass1FilterOrder + buildAnalysisFilter (reference filters.go:42-45, 684-689)
    assert p1 == ("aformat=channel_layouts=mono,astats=metadata=1:measure_perchannel=all,"
                  "aspectralstats=win_size=2048:win_func=hann:measure=all,"
                  "ebur128=metadata=1:peak=sample+true:dualmono=true:target=-16")
    p2 = gpudsp.default_pass2_spec()
    assert p2.startswith("aformat=channel_layouts=mono,highpass=f=80:poles=2:width_type=q:width=0.707:normalize=1:a=tdii,")
    assert p2.endswith("aformat=sample_rates=44100:channel_layouts=mono:sample_fmts=s16,asetnsamples=n=4096")
    assert "anlmdn=s=0.00001:p=0.0060:r=0.0020:m=3,afftdn=nr=12:nt=w:tn=1," in p2


def test_no_device_fails_loudly():
    """Without a CUDA device jt_create must fail (no CPU fallback); with one it must succeed."""
    import torch
    from jivetalking_b200 import gpudsp
    if torch.cuda.is_available():
        gpudsp.Context(0).close()
    else:
        try:
            gpudsp.Context(0)
        except gpudsp.JtError as e:
            assert e.code == -2
        else:
            raise AssertionError("jt_create succeeded without a GPU")


def test_pass34_spec_builders_match_reference_goldens():
    """Golden strings of the reference: internal/processor/normalise_test.go:2158-2188."""
    from jivetalking_b200 import gpudsp
    cases = [
        (-20.0, -10.0, dict(i=-20.0, tp=-10.0, lra=5.0, th=-30.0), "",
         "loudnorm=I=-16.00:TP=-5.70:LRA=20.0:measured_I=-20.00:measured_TP=-10.00:measured_LRA=5.00:measured_thresh=-30.00:offset=4.00:dual_mono=true:linear=true:print_format=json,aresample=48000,adeclick=t=1.7:w=55:o=50:m=s,alimiter=limit=0.803526:attack=1:release=50:level_in=1:level_out=1:level=0:latency=1:asc=1:asc_level=0.8,astats=metadata=1:measure_perchannel=all,aspectralstats=win_size=2048:win_func=hann:measure=all,ebur128=metadata=1:peak=sample+true:dualmono=true,aformat=sample_rates=44100:channel_layouts=mono:sample_fmts=s16,asetnsamples=n=4096"),
        (-24.9, -5.0, dict(i=-24.9, tp=-5.0, lra=6.0, th=-35.0),
         "alimiter=limit=0.319890:attack=5:release=100:level_in=1:level_out=1:level=0:latency=1:asc=1:asc_level=0.8", None),
        (-43.2, -18.6, dict(i=-36.5, tp=-24.0, lra=8.0, th=-46.5),
         "volume=4.2dB,alimiter=limit=0.063096:attack=5:release=100:level_in=1:level_out=1:level=0:latency=1:asc=1:asc_level=0.8", None),
    ]
    for out_i, out_tp, m, want_prefix, want_p4 in cases:
        spec3, plan = gpudsp.build_pass3_spec(out_i, out_tp)
        tail = "loudnorm=I=-16.0:TP=-1.0:LRA=20.0:dual_mono=true:print_format=json"
        assert spec3 == (want_prefix + "," + tail if want_prefix else tail)
        st = gpudsp.LoudnormStats()
        st.input_i, st.input_tp, st.input_lra, st.input_thresh = m["i"], m["tp"], m["lra"], m["th"]
        spec4, eff, off = gpudsp.build_pass4_spec(plan, st, source_rate=48000)
        assert spec4.startswith(want_prefix + ",loudnorm=" if want_prefix else "loudnorm=")
        assert abs(eff - (-16.0)) < 1e-9
        if want_p4:
            # the reference fixture passes offset=0.00 explicitly (its `offset` argument); ours derives
            # offset = effectiveTargetI - measured_I as ApplyNormalisation does (normalise.go:873)
            assert spec4.replace("offset=4.00", "offset=0.00") == want_p4.replace("offset=4.00", "offset=0.00")


def test_wav_parse(tmp_path):
    """jt_wav_parse on files written by Python's wave module (the reference's fixtures are s16 WAVs, testutil_test.go:140-190)"""
    import io
    import struct
    import wave
    import numpy as np
    import pytest
    from jivetalking_b200 import gpudsp
    x = (np.arange(2000) % 257 - 128).astype(np.int16)
    wavs = []
    for ch in (1, 2):
        bio = io.BytesIO()
        with wave.open(bio, "wb") as w:
            w.setnchannels(ch); w.setsampwidth(2); w.setframerate(48000); w.writeframes(x.tobytes())
        wavs.append(bio.getvalue())
        pcm, rate, channels = gpudsp.wav_parse(bio.getvalue())
        assert (rate, channels) == (48000, ch) and pcm.dtype == np.int16 and np.array_equal(pcm, x)
    # float32 WAV with a LIST chunk in front of the data and an odd-sized chunk (padding byte)
    f = np.linspace(-1, 1, 333, dtype=np.float32)
    fmt = struct.pack("<HHIIHH", 3, 1, 44100, 44100 * 4, 4, 32)
    body = b"WAVE" + b"fmt " + struct.pack("<I", 16) + fmt + b"LIST" + struct.pack("<I", 5) + b"abcde\0" + b"data" + struct.pack("<I", f.nbytes) + f.tobytes()
    wavs.append(b"RIFF" + struct.pack("<I", len(body)) + body)
    pcm, rate, channels = gpudsp.wav_parse(wavs[-1])
    assert rate == 44100 and channels == 1 and pcm.dtype == np.float32 and np.array_equal(pcm, f)
    # ... and the reference's actual reader (libavformat demuxer + libavcodec PCM decoder, internal/audio/reader.go) sees
    # the same samples in every case above, when the real FFmpeg libraries are in the image
    import ref_flac
    for data in wavs:
        ref = ref_flac.ref_wav_read(data)
        if ref is not None:
            g, grate, gch = gpudsp.wav_parse(data)
            assert (ref[1], ref[2]) == (grate, gch) and ref[0].dtype == g.dtype and np.array_equal(ref[0], g)
    # 24-bit PCM and garbage fail loudly
    bio = io.BytesIO()
    with wave.open(bio, "wb") as w:
        w.setnchannels(1); w.setsampwidth(3); w.setframerate(48000); w.writeframes(b"\0" * 300)
    with pytest.raises(gpudsp.JtError) as e:
        gpudsp.wav_parse(bio.getvalue())
    assert e.value.code == -5
    with pytest.raises(gpudsp.JtError):
        gpudsp.wav_parse(b"not a wav file at all")


def test_limiter_planners_match_reference_tables():
    """a9 planners through jt_build_pass3_spec / jt_build_pass4_spec against the reference's tables:
    TestCalculateLimiterCeiling (normalise_test.go:1204-1388), TestCalculatePreGain (:1987-2044),
    TestPreGainCeilingRederivation (:1764-1852), TestLoudnormInternalTargetTPCancellation (:1029-1062)."""
    import pytest
    from jivetalking_b200 import gpudsp
    MIN = -24.0
    table = [  # measuredI, measuredTP, targetI, targetTP, wantCeiling, wantNeeded, wantClamped
        (-24.9, -5.0, -16.0, -2.0, -10.9, True, False), (-20.0, -3.0, -16.0, -2.0, -6.0, True, False),
        (-20.0, -10.0, -16.0, -2.0, 0.0, False, False), (-12.0, -1.0, -16.0, -2.0, 0.0, False, False),
        (-20.0, -6.0, -16.0, -2.0, 0.0, False, False), (-43.0, -20.0, -16.0, -2.0, MIN, True, True),
        (-40.0, -15.0, -16.0, -2.0, MIN, True, True), (-33.5, -15.0, -16.0, -2.0, -19.5, True, False),
        (-43.2, -18.6, -16.0, -2.0, MIN, True, True), (-36.6, -15.0, -16.0, -2.0, -22.6, True, False)]
    for mi, mtp, ti, ttp, want, needed, clamped in table:
        spec3, plan = gpudsp.build_pass3_spec(mi, mtp, ti, ttp)
        assert bool(plan.limiter_needed) == needed and bool(plan.limiter_clamped) == clamped, (mi, mtp)
        if needed:
            # a clamped ceiling is re-derived after the pre-gain and lands on the minimum again (TestPreGainCeilingRederivation)
            assert abs(plan.limiter_ceiling_db - want) < 0.01, (mi, mtp, plan.limiter_ceiling_db)
            assert ("alimiter=limit=%.6f" % 10 ** (plan.limiter_ceiling_db / 20)) in spec3
        else:
            assert "alimiter" not in spec3 and "volume" not in spec3
        ideal = ttp - (ti - mi)
        assert abs(plan.limiter_pregain_db - (MIN - ideal if ideal < MIN else 0.0)) < 1e-9
        assert (("volume=%.1fdB," % plan.limiter_pregain_db) in spec3) == (clamped and plan.limiter_pregain_db > 0)
    # TestCalculatePreGain
    for mi, want_gain, want_ceil in ((-43.2, 5.2, -24.0), (-24.9, 0.0, None), (-38.0, 0.0, None)):
        _, plan = gpudsp.build_pass3_spec(mi, -10.0 if want_ceil is None else -18.6, -16.0, -2.0)
        assert abs(plan.limiter_pregain_db - want_gain) < 0.01
        if want_ceil is not None:
            assert abs(plan.limiter_ceiling_db - want_ceil) < 0.01
    # the internal TP target cancels out: linear mode always reaches the desired loudness, emitted TP clamped to [-9, 0]
    for mi, mtp in ((-24.9, -5.0), (-36.5, -24.0), (-12.0, -1.0), (-30.0, -3.0)):
        _, plan = gpudsp.build_pass3_spec(mi, mtp)
        st = gpudsp.LoudnormStats()
        st.input_i, st.input_tp, st.input_lra, st.input_thresh = mi, mtp, 6.0, mi - 10.0
        spec4, eff, off = gpudsp.build_pass4_spec(plan, st)
        assert eff == pytest.approx(-16.0) and off == pytest.approx(-16.0 - mi)
        tp = float(spec4.split("loudnorm=")[1].split(":TP=")[1].split(":")[0])
        assert -9.0 <= tp <= 0.0 and abs(tp - max(-9.0, min(0.0, mtp + (-16.0 - mi) + 0.3))) < 0.006


def test_loudnorm_stats_json_has_the_reference_wire_shape():
    """jt_loudnorm_stats_json renders what af_loudnorm writes to stats_file: the ten keys of LoudnormStats
    (normalise.go:64-75) as decimal STRINGS ("%.2f"), which the reference parses with strconv.ParseFloat
    (normalise.go:321-343; fixture shape: normalise_test.go:19)."""
    import ctypes as C
    import json
    from jivetalking_b200 import gpudsp
    L = gpudsp.lib()
    st = gpudsp.LoudnormStats()
    st.input_i, st.input_tp, st.input_lra, st.input_thresh = -23.004, -4.0, 5.0, -33.0
    st.output_i, st.output_tp, st.output_lra, st.output_thresh = -16.0, -2.0, 5.0, -26.0
    st.target_offset, st.normalization_type, st.valid = 0.0, 0, 1
    buf = C.create_string_buffer(1024)
    assert L.jt_loudnorm_stats_json(C.byref(st), buf, len(buf)) == 0
    doc = json.loads(buf.value.decode())
    assert sorted(doc) == sorted(["input_i", "input_tp", "input_lra", "input_thresh", "output_i", "output_tp", "output_lra",
                                  "output_thresh", "normalization_type", "target_offset"])
    assert all(isinstance(v, str) for v in doc.values())
    assert doc["input_i"] == "-23.00" and doc["normalization_type"] == "linear" and float(doc["output_tp"]) == -2.0
    st.normalization_type = 1
    L.jt_loudnorm_stats_json(C.byref(st), buf, len(buf))
    assert json.loads(buf.value.decode())["normalization_type"] == "dynamic"
    assert L.jt_loudnorm_stats_json(C.byref(st), buf, 10) == -7
