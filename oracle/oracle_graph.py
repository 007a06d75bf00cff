"""ORACLE (test infrastructure only): composition of the CPU oracle filters (oracle/*.c) into the sink-frame metadata wire the
reference reads (analyser_metrics.go:432-483), independent of the product's C++ executor.
Frame cadence follows libavfilter's re-framing rules as described in SURVEY.md 8a-bis:
4096-sample decoder frames -> astats stamps (cumulative at frame end) -> aspectralstats
re-frames to 1024-sample hops (props of the frame holding the hop's first sample) ->
ebur128 re-frames to 100 ms (props of the hop frame holding the tick's first sample)."""
import math
import numpy as np
import jt_oracle as O

NAN = float("nan")
AS_NAMES = ["Dynamic_range", "RMS_level", "Peak_level", "RMS_trough", "RMS_peak", "DC_offset", "Flat_factor",
            "Crest_factor", "Zero_crossings_rate", "Zero_crossings", "Max_difference", "Min_difference",
            "Mean_difference", "RMS_difference", "Entropy", "Min_level", "Max_level", "Noise_floor",
            "Noise_floor_count", "Bit_depth", "Number_of_samples"]


def wire(fmt, v):
    """what strconv.ParseFloat reads back from FFmpeg's snprintf"""
    return float(fmt % v)


def downmix(x, channels):
    x = np.ascontiguousarray(x)
    if channels == 1:
        return x
    assert channels == 2
    n = x.size // 2
    if x.dtype == np.int32:
        # swr's internal format for a 32-bit integer rematrix is FLTP; the matrix is the normalised 0.5 / 0.5 (pinned on the
        # real libswresample: tests/golden/swr_golden.npz)
        L, R = to_f32(x[0::2]), to_f32(x[1::2])
        mix = (L * np.float32(0.5) + R * np.float32(0.5)).astype(np.float32)
        return f32_to_s32(mix)
    out = np.zeros(n, dtype=x.dtype)
    fn = {np.dtype(np.float32): O.lib().orc_downmix_stereo_f32, np.dtype(np.int16): O.lib().orc_downmix_stereo_s16}[x.dtype]
    fn(x.ctypes.data_as(O._P), n, out.ctypes.data_as(O._P))
    return out


def f32_to_s32(v):
    """audioconvert.c: av_clipl_int32(llrintf(v * (1U << 31)))"""
    s = np.rint((v.astype(np.float32) * np.float32(2147483648.0)).astype(np.float64))
    return np.clip(s, -2147483648.0, 2147483647.0).astype(np.int32)


def f64_to_s32(v):
    return np.clip(np.rint(v.astype(np.float64) * 2147483648.0), -2147483648.0, 2147483647.0).astype(np.int32)


def to_f32(x):
    if x.dtype == np.float32:
        return x
    if x.dtype == np.int32:
        return (x.astype(np.float32) * np.float32(1.0 / 2147483648.0)).astype(np.float32)
    if x.dtype == np.int16:
        return (x.astype(np.float32) * np.float32(1.0 / 32768.0)).astype(np.float32)
    return x.astype(np.float32)


def to_f64(x):
    if x.dtype == np.int16:
        return x.astype(np.float64) * (1.0 / 32768.0)
    if x.dtype == np.int32:
        return x.astype(np.float64) * (1.0 / 2147483648.0)
    return x.astype(np.float64)


def analysis_meta(astats_in, spec_in_f32, r128_in_f64, rate, src_frame_bounds, dualmono=True, true_peak=True,
                  astats_overall_only=False):
    """Sink-frame records of `astats -> aspectralstats -> ebur128` given the signal each filter
    sees and the frame boundaries (ascending end positions) on the astats link."""
    n = len(r128_in_f64)
    T = rate // 10
    rows = O.aspectralstats(spec_in_f32, rate)
    r = O.ebur128(r128_in_f64, rate, dualmono=dualmono, true_peak=true_peak)
    ends = np.asarray(src_frame_bounds, dtype=np.int64)
    starts = np.concatenate(([0], ends[:-1]))

    def src_frame_of(sample):
        return int(np.searchsorted(starts, sample, side="right") - 1)

    recs = []
    nticks_total = (n + T - 1) // T
    last_tick = n // T - 1
    for k in range(nticks_total):
        s0 = k * T
        nb = min(T, n - s0)
        j = s0 // 1024                                   # hop frame holding the first sample
        f = src_frame_of(j * 1024)                       # src frame holding that hop's first sample
        last = s0 + nb - 1
        j2 = last // 1024
        hop_last = min(1024 * (j2 + 1), n) - 1
        ready = int(ends[src_frame_of(hop_last)])
        rec = dict(first_sample=s0, nb_samples=nb, ready=ready, astats_pos=int(ends[f]), hop=j,
                   M=NAN, S=NAN, I=NAN, LRA=NAN, true_peak=NAN, sample_peak=NAN,
                   spectral=[wire("%g", float(v)) for v in rows[j]], astats=None)
        if nb == T and k < r["n_ticks"]:
            rec.update(M=wire("%.3f", r["M"][k]), S=wire("%.3f", r["S"][k]),
                       sample_peak=wire("%.3f", r["sample_peak_cum"][k]))
            if true_peak:
                rec.update(true_peak=wire("%.3f", r["true_peak_cum"][k]))
            if k == last_tick:
                rec.update(I=wire("%.3f", r["I"]), LRA=wire("%.3f", r["LRA"]))
        recs.append(rec)
    if recs:
        a = O.astats(astats_in[: recs[-1]["astats_pos"]], rate)
        recs[-1]["astats"] = {k: wire("%f", a[k]) for k in AS_NAMES if k != "Number_of_samples"}
        recs[-1]["astats"]["Number_of_samples"] = a["nb_samples"]
        recs[-1]["overall_only"] = astats_overall_only
    return recs


def pass1_meta(x, rate, channels=1, frame_size=4096):
    mono = downmix(x, channels)
    n = len(mono)
    ends = [min((f + 1) * frame_size, n) for f in range((n + frame_size - 1) // frame_size)]
    # aspectralstats turns the link to flt; ebur128 widens THAT (exact for s16 / flt input, float-rounded for s32)
    return analysis_meta(mono, to_f32(mono), to_f64(to_f32(mono)) if mono.dtype == np.int32 else to_f64(mono), rate, ends)


def pass1_analyse(x, rate, channels=1, frame_size=4096):
    """ORACLE (test infrastructure) restatement of collectAnalysisFrames (analyser.go:538-650): the Go-side accumulation
    of the Pass-1 sink-frame records into whole-file values and 250 ms IntervalSamples (analyser_metrics.go:165-420).
    Returns (measurements dict, [interval dict]); interval["spectral"] is the 13-value mean in aspectralstats order."""
    exp = pass1_meta(x, rate, channels, frame_size)
    F = frame_size
    n = len(x) // channels
    if x.dtype == np.int16:
        xs = x.astype(np.float64) / 32768.0
    elif x.dtype == np.int32:
        xs = x.astype(np.float64) / 2147483648.0
    else:
        xs = x.astype(np.float64)
    xs = xs.reshape(n, channels) if channels > 1 else xs
    nsrc = (n + F - 1) // F

    def db(v):
        return -120.0 if v <= 0 else 20 * math.log10(v)

    def new_acc(first):
        return dict(fc=0, ss=0.0, n=0, pk=0.0, M=0.0, S=0.0, tp=0.0 if first else -120.0, sp=0.0 if first else -120.0,
                    spec=[0.0] * 13, found=False)

    def fin(a, ts):
        rms = math.sqrt(a["ss"] / a["n"]) if a["n"] else 0.0
        fc = a["fc"]
        return dict(ts_ns=ts, rms=-120.0 if (a["n"] == 0 or rms < 1e-5) else 20 * math.log10(rms),
                    pk=20 * math.log10(a["pk"]) if a["pk"] > 0 else -120.0,
                    M=a["M"] / fc if fc else 0.0, S=a["S"] / fc if fc else 0.0, tp=a["tp"], sp=a["sp"], fc=fc,
                    spectral=[v / fc for v in a["spec"]] if fc else [0.0] * 13, found=a["found"])

    whole = dict(I=NAN, M=NAN, S=NAN, LRA=NAN, tp=NAN, sp=NAN, astats=None, spec_sum=[0.0] * 13, spec_n=0)

    def add_sink(acc, e):
        tp = 0.0 if math.isnan(e["true_peak"]) else db(e["true_peak"])
        sp = 0.0 if math.isnan(e["sample_peak"]) else db(e["sample_peak"])
        if acc["fc"] == 0 or tp > acc["tp"]:
            acc["tp"] = tp
        if acc["fc"] == 0 or sp > acc["sp"]:
            acc["sp"] = sp
        for k, v in enumerate(e["spectral"]):
            if not math.isnan(v):
                acc["spec"][k] += v
                acc["found"] = True
        acc["M"] += 0.0 if math.isnan(e["M"]) else e["M"]
        acc["S"] += 0.0 if math.isnan(e["S"]) else e["S"]
        acc["fc"] += 1
        if any(not math.isnan(v) for v in e["spectral"]):
            whole["spec_n"] += 1
            for k, v in enumerate(e["spectral"]):
                whole["spec_sum"][k] += 0.0 if math.isnan(v) else v
        for key, src in (("I", "I"), ("M", "M"), ("S", "S"), ("LRA", "LRA")):
            if not math.isnan(e[src]):
                whole[key] = e[src]
        if not math.isnan(e["true_peak"]):
            whole["tp"] = db(e["true_peak"])
        if not math.isnan(e["sample_peak"]):
            whole["sp"] = db(e["sample_peak"])
        if e["astats"] is not None:
            whole["astats"] = e["astats"]

    intervals, acc, start_ns, pushed, sink = [], new_acc(True), 0, 0, 0
    for f in range(nsrc):
        t_ns = int(pushed / rate * 1e9)
        seg = xs[pushed:pushed + F]
        pushed += len(seg)
        acc["ss"] += float(np.sum(seg * seg)); acc["n"] += seg.size
        acc["pk"] = max(acc["pk"], float(np.max(np.abs(seg))) if seg.size else 0.0)
        if t_ns - start_ns >= 250_000_000:
            intervals.append(fin(acc, start_ns)); start_ns = t_ns; acc = new_acc(False)
        while sink < len(exp) and exp[sink]["ready"] <= pushed:
            add_sink(acc, exp[sink]); sink += 1
    while sink < len(exp):
        add_sink(acc, exp[sink]); sink += 1
    if acc["n"] > 0:
        intervals.append(fin(acc, start_ns))
    meas = dict(input_i=whole["I"], input_tp=whole["tp"], input_sp=whole["sp"], input_lra=whole["LRA"], last_m=whole["M"], last_s=whole["S"],
                astats=whole["astats"], spectral_mean=[v / whole["spec_n"] for v in whole["spec_sum"]] if whole["spec_n"] else [0.0] * 13,
                sink_frames=len(exp), duration_s=n / rate)
    return meas, intervals


REGION_SPEC = ("atrim=start=%f:duration=%f,asetpts=PTS-STARTPTS,astats=metadata=1:measure_perchannel=0,"
               "aspectralstats=measure=all,ebur128=metadata=1:peak=sample+true")          # analyser_output.go:18


def region_sample(pcm, rate, start_ns, dur_ns):
    """ORACLE (test infrastructure) restatement of measureOutputRegionFromReader (analyser_output.go:95-233): the region
    graph over the oracle filters and the Go-side reduction of its sink frames.  Returns (dict, frames processed)."""
    def secs(ns):
        return float(ns // 1_000_000_000) + float(ns % 1_000_000_000) / 1e9
    exp = run_spec(REGION_SPEC % (secs(start_ns), secs(dur_ns)), pcm, rate, want_pcm=False)["meta"]
    rms = peak = M = S = tp = sp = 0.0
    rms_found, spec_sum, spec_n = False, [0.0] * 13, 0
    for e in exp:
        if e["astats"] is not None:
            if not math.isnan(e["astats"]["RMS_level"]):
                rms, rms_found = e["astats"]["RMS_level"], True
            if not math.isnan(e["astats"]["Peak_level"]):
                peak = e["astats"]["Peak_level"]
        if any(not math.isnan(v) for v in e["spectral"]):
            spec_n += 1
            for k, v in enumerate(e["spectral"]):
                spec_sum[k] += 0.0 if math.isnan(v) else v
        M = e["M"] if not math.isnan(e["M"]) else M
        S = e["S"] if not math.isnan(e["S"]) else S
        tp = e["true_peak"] if not math.isnan(e["true_peak"]) else tp
        sp = e["sample_peak"] if not math.isnan(e["sample_peak"]) else sp

    def db(v):
        return -120.0 if v <= 0 else 20 * math.log10(v)
    out = dict(rms_level=rms if rms_found else -60.0, peak_level=peak, crest_factor=peak - rms if (rms_found and peak != 0) else 0.0,
               spectral=[v / spec_n for v in spec_sum] if spec_n else [0.0] * 13, momentary_lufs=M, short_term_lufs=S,
               true_peak=db(tp), sample_peak=db(sp))
    return out, len(exp)


# ---------------------------------------------------------------------------------------------
# whole-graph oracle: parses the reference's spec strings and chains the oracle filters with the
# sample-format conversions libavfilter's negotiation would insert (hard part 4)
# ---------------------------------------------------------------------------------------------
def parse_spec(spec):
    nodes = []
    for f in spec.split(","):
        name, _, rest = f.partition("=")
        opts, pos = {}, []
        if rest:
            for o in rest.split(":"):
                k, eq, v = o.partition("=")
                if eq:
                    opts[k] = v
                else:
                    pos.append(k)
        nodes.append((name, opts, pos))
    return nodes


def _get(opts, *names, default=None):
    for n in names:
        if n in opts:
            return opts[n]
    return default


def _reframe(frames, n, F):
    """frames: list of dict(start, nb, ready, astats_pos, hop, tick) -> uniform F-sample frames"""
    out, a, b = [], 0, 0
    for s in range(0, n, F):
        nb = min(F, n - s)
        while a + 1 < len(frames) and frames[a + 1]["start"] <= s:
            a += 1
        b = max(b, a)
        while b + 1 < len(frames) and frames[b + 1]["start"] <= s + nb - 1:
            b += 1
        f = dict(frames[a]) if frames else dict(astats_pos=-1, hop=-1, tick=-1, ready=0)
        f.update(start=s, nb=nb, ready=frames[b]["ready"] if frames else 0)
        out.append(f)
    return out


def run_spec(spec, x, rate, channels=1, frame_size=4096, want_pcm=True):
    """Returns dict(pcm, rate, meta=[records like analysis_meta], loudnorm=dict or None)."""
    cur = downmix(x, channels) if channels > 1 else np.ascontiguousarray(x)
    n = len(cur)
    frames = [dict(start=s, nb=min(frame_size, n - s), ready=min(s + frame_size, n), astats_pos=-1, hop=-1, tick=-1)
              for s in range(0, n, frame_size)]
    ln = None
    astats_sig = spec_rows = r128 = None
    nodes = parse_spec(spec)

    def as_fmt(sig, dt):
        if sig.dtype == dt:
            return sig
        if dt == np.float64:
            return to_f64(sig)
        if dt == np.float32:
            return to_f32(sig) if sig.dtype in (np.int16, np.int32) else sig.astype(np.float32)
        if dt == np.int32:
            return f32_to_s32(sig) if sig.dtype == np.float32 else f64_to_s32(sig) if sig.dtype == np.float64 else sig.astype(np.int32) * 65536
        if dt == np.int16 and sig.dtype == np.int32:
            return (sig >> 16).astype(np.int16)
        if dt == np.int16:
            out = np.zeros(len(sig), dtype=np.int16)
            (O.lib().orc_conv_f64_to_s16 if sig.dtype == np.float64 else O.lib().orc_conv_f32_to_s16)(O._ptr(np.ascontiguousarray(sig)), len(sig), O._ptr(out))
            return out
        raise ValueError(dt)

    def resample(sig, out_rate, out_dt, frames):
        nonlocal rate
        if out_rate == rate:
            return as_fmt(sig, out_dt), frames
        bi, bo = sig.dtype.itemsize, np.dtype(out_dt).itemsize
        work = np.float32 if bi <= 4 else np.float64
        assert not (bi <= 2 and bo <= 2)
        src = as_fmt(sig, work)
        y = O.swr_resample(src, rate, out_rate, flush=True)
        nf, done = [], 0
        for f in frames:
            cnt = min(O.swr_out_count(f["start"] + f["nb"], rate, out_rate), len(y))
            if cnt > done:
                g = dict(f); g.update(start=done, nb=cnt - done); nf.append(g); done = cnt
        if len(y) > done:
            nf.append(dict(start=done, nb=len(y) - done, ready=1 << 62, astats_pos=-1, hop=-1, tick=-1))
        rate = out_rate
        return as_fmt(y, out_dt), nf

    for idx, (name, o, pos) in enumerate(nodes):
        last = idx == len(nodes) - 1
        if name == "aformat":
            out_rate = int(_get(o, "sample_rates", "r", default=rate))
            sf = _get(o, "sample_fmts", "f", default="")
            dt = {"s16": np.int16, "s32": np.int32, "flt": np.float32, "dbl": np.float64, "": cur.dtype}[sf]
            if last and not want_pcm and out_rate != rate:
                tot = O.swr_out_count(len(cur), rate, out_rate, flush=True)
                nf, done = [], 0
                for f in frames:
                    cnt = min(O.swr_out_count(f["start"] + f["nb"], rate, out_rate), tot)
                    if cnt > done:
                        g = dict(f); g.update(start=done, nb=cnt - done); nf.append(g); done = cnt
                if tot > done:
                    nf.append(dict(start=done, nb=tot - done, ready=1 << 62, astats_pos=-1, hop=-1, tick=-1))
                frames, rate, cur = nf, out_rate, np.zeros(tot, dtype=np.int16)
            else:
                cur, frames = resample(cur, out_rate, dt, frames)
        elif name == "aresample":
            cur, frames = resample(cur, int(pos[0]) if pos else int(o["sample_rate"]), cur.dtype, frames)
        elif name == "asetnsamples":
            nn = int(_get(o, "n", "nb_out_samples", default=1024))
            frames = _reframe(frames, len(cur), nn)
            if frames and frames[-1]["nb"] < nn:
                tot = frames[-1]["start"] + nn
                frames[-1]["nb"] = nn
                cur = np.concatenate([cur, np.zeros(tot - len(cur), dtype=cur.dtype)])
        elif name == "atrim":
            st_us, du_us = round(float(o.get("start", 0)) * 1e6), round(float(o.get("duration", 0)) * 1e6)
            s0 = (st_us * rate + 500000) // 1000000
            ln_ = (du_us * rate + 500000) // 1000000 if du_us > 0 else 1 << 60
            a, b = min(max(s0, 0), len(cur)), min(len(cur), s0 + ln_)
            nf = []
            for f in frames:
                lo, hi = max(f["start"], a), min(f["start"] + f["nb"], b)
                if hi > lo:
                    g = dict(f); g.update(start=lo - a, nb=hi - lo); nf.append(g)
            frames, cur = nf, cur[a:max(a, b)]
        elif name == "asetpts":
            pass
        elif name in ("highpass", "lowpass"):
            cur = O.biquad(cur, rate, name, float(_get(o, "f", "frequency", default=3000)), float(_get(o, "w", "width", default=0.707)),
                           normalize=_get(o, "n", "normalize", default="0") in ("1", "true"),
                           tdii=_get(o, "a", "transform", default="di") == "tdii", mix=float(_get(o, "m", "mix", default=1.0)))
        elif name == "anlmdn":
            cur = as_fmt(cur, np.float32)
            p = float(_get(o, "p", default=0.002))
            cur = O.anlmdn(cur, rate, float(_get(o, "s", default=0.00001)), p, float(_get(o, "r", default=0.006)), float(_get(o, "m", default=11)))
            frames = _reframe(frames, len(cur), 2 * int((round(p * 1e6) * rate + 500000) // 1000000) + 1)
        elif name == "afftdn":
            cur = as_fmt(cur, np.float32)
            nt = {"w": 0, "white": 0, "v": 1, "s": 2, "custom": 3, "c": 3}[_get(o, "nt", default="w")]
            bn = [float(v) for v in o["bn"].split("|") if v.strip()] if "bn" in o else None
            cur = O.afftdn(cur, rate, float(_get(o, "nr", default=12)), float(_get(o, "nf", default=-50)), nt, bn,
                           _get(o, "tn", default="0") in ("1", "true"), float(_get(o, "ad", default=0.5)))
            frames = _reframe(frames, len(cur), rate // 80)
        elif name == "agate":
            cur = O.agate(as_fmt(cur, np.float64), rate, float(o.get("threshold", 0.125)), float(o.get("ratio", 2)), float(o.get("attack", 20)),
                          float(o.get("release", 250)), float(o.get("range", 0.06125)), float(o.get("knee", 2.828427125)),
                          float(o.get("makeup", 1)), o.get("detection", "rms") == "rms")
        elif name == "acompressor":
            cur = O.acompressor(as_fmt(cur, np.float64), rate, float(o.get("threshold", 0.125)), float(o.get("ratio", 2)), float(o.get("attack", 20)),
                                float(o.get("release", 250)), float(o.get("makeup", 1)), float(o.get("knee", 2.82843)),
                                float(o.get("mix", 1)), o.get("detection", "rms") == "rms")
        elif name == "deesser":
            cur = O.deesser(as_fmt(cur, np.float64), rate, float(o.get("i", 0)), float(o.get("m", 0.5)), float(o.get("f", 0.5)))
        elif name == "volume":
            v = pos[0] if pos else o["volume"]
            g = 10 ** (float(v[:-2]) / 20.0) if v.endswith("dB") else float(v)
            cur = O.volume_f32(as_fmt(cur, np.float32), g)
        elif name == "alimiter":
            cur = O.alimiter(as_fmt(cur, np.float64), rate, float(o.get("limit", 1)), float(o.get("attack", 5)), float(o.get("release", 50)),
                             float(o.get("level_in", 1)), float(o.get("level_out", 1)), o.get("level", "1") in ("1", "true"),
                             o.get("asc", "0") in ("1", "true"), float(o.get("asc_level", 0.5)))
        elif name == "adeclick":
            w, ov = float(_get(o, "w", default=55)), float(_get(o, "o", default=75))
            cur, _ = O.adeclick(as_fmt(cur, np.float64), rate, w, ov, float(_get(o, "a", default=2)), float(_get(o, "t", default=2)),
                                float(_get(o, "b", default=2)), _get(o, "m", default="a") in ("s", "save"))
            ws = int(rate * w / 1000.)
            frames = _reframe(frames, len(cur), max(int(ws * (1. - ov / 100.)), 1))
        elif name == "loudnorm":
            I, TP, LRA = float(_get(o, "I", "i", default=-24)), float(_get(o, "TP", "tp", default=-2)), float(_get(o, "LRA", "lra", default=7))
            mI, mTP = float(_get(o, "measured_I", "measured_i", default=0)), float(_get(o, "measured_TP", "measured_tp", default=99))
            mLRA, mTh = float(_get(o, "measured_LRA", "measured_lra", default=0)), float(o.get("measured_thresh", -70))
            dual = o.get("dual_mono", "false") == "true"
            offs = float(o.get("offset", 0))
            lin = O.loudnorm_is_linear(I, TP, LRA, mI, mTP, mLRA, mTh, o.get("linear", "true") == "true")
            if not lin:
                # query_formats(): outside linear mode both links of the filter run at 192 kHz / dbl
                cur, frames = resample(cur, 192000, np.float64, frames)
            cur = as_fmt(cur, np.float64)
            if not lin and last and not want_pcm and len(cur) >= 576000:
                # measure-only call (Pass 3): the four input_* values are those of the stream followed by its last 2.9 s
                # again (filter_frame meters the flush frame a second time); the discarded output is not produced
                mi = O.loudnorm_meter(np.concatenate([cur, cur[len(cur) - 556800:]]), rate, dual)
                ln = dict(normalization_type=1, input_i=mi["I"], input_tp=20 * math.log10(mi["sample_peak"]) if mi["sample_peak"] > 0 else -math.inf,
                          input_lra=mi["LRA"], input_thresh=mi["thresh"])
            else:
                n_in = len(cur)
                cur, ln = O.loudnorm(cur, rate, I, TP, LRA, mI, mTP, mLRA, mTh, offs, o.get("linear", "true") == "true", dual)
                if not lin:
                    # output frames: 100 ms after the 3 s first frame, one per consumed 100 ms frame, then the 2.9 s flush frame
                    if n_in < 576000:
                        pass                      # single-gain fall-back: frames pass through
                    else:
                        nf, pos, src_pos = [], 0, 576000
                        def ready_at(p):
                            f = next((f for f in frames if f["start"] <= p - 1 < f["start"] + f["nb"]), frames[-1])
                            return f["ready"]
                        nf.append(dict(start=0, nb=19200, ready=ready_at(576000), astats_pos=-1, hop=-1, tick=-1)); pos = 19200
                        while src_pos < n_in:
                            nb = min(19200, n_in - src_pos); src_pos += nb
                            nf.append(dict(start=pos, nb=nb, ready=ready_at(src_pos) if nb == 19200 else 1 << 62, astats_pos=-1, hop=-1, tick=-1)); pos += nb
                        nf.append(dict(start=pos, nb=556800, ready=1 << 62, astats_pos=-1, hop=-1, tick=-1))
                        frames = nf
        elif name == "astats":
            astats_sig, astats_rate = cur, rate
            for f in frames:
                f["astats_pos"] = f["start"] + f["nb"]
            astats_overall = o.get("measure_perchannel", "all") in ("0", "none")
        elif name == "aspectralstats":
            cur = as_fmt(cur, np.float32)
            spec_rows = O.aspectralstats(cur, rate, int(o.get("win_size", 2048)))
            frames = _reframe(frames, len(cur), int(o.get("win_size", 2048)) // 2)
            for j, f in enumerate(frames):
                f["hop"] = j
        elif name == "ebur128":
            cur = as_fmt(cur, np.float64)
            tp = "true" in o.get("peak", "none")
            r128 = O.ebur128(cur, rate, dualmono=o.get("dualmono", "false") == "true", true_peak=tp)
            r128["tp"] = tp
            T = rate // 10
            frames = _reframe(frames, len(cur), T)
            for k, f in enumerate(frames):
                if f["nb"] == T:
                    f["tick"] = k
        else:
            raise ValueError("unsupported filter " + name)

    # sink-frame records
    recs = []
    last_tick = max([f["tick"] for f in frames] + [-1])
    last_as = max([i for i, f in enumerate(frames) if f["astats_pos"] >= 0] + [-1])
    for i, f in enumerate(frames):
        rec = dict(first_sample=f["start"], nb_samples=f["nb"], ready=f["ready"], M=NAN, S=NAN, I=NAN, LRA=NAN, true_peak=NAN,
                   sample_peak=NAN, spectral=[NAN] * 13, astats=None)
        if r128 is not None and f["tick"] >= 0 and f["tick"] < r128["n_ticks"]:
            k = f["tick"]
            rec.update(M=wire("%.3f", r128["M"][k]), S=wire("%.3f", r128["S"][k]), sample_peak=wire("%.3f", r128["sample_peak_cum"][k]))
            if r128["tp"]:
                rec.update(true_peak=wire("%.3f", r128["true_peak_cum"][k]))
            if k == last_tick:
                rec.update(I=wire("%.3f", r128["I"]), LRA=wire("%.3f", r128["LRA"]))
        if spec_rows is not None and 0 <= f["hop"] < len(spec_rows):
            rec["spectral"] = [wire("%g", float(v)) for v in spec_rows[f["hop"]]]
        if astats_sig is not None and i == last_as:
            a = O.astats(astats_sig[: f["astats_pos"]], astats_rate)
            rec["astats"] = {k: wire("%f", a[k]) for k in AS_NAMES if k != "Number_of_samples"}
            rec["astats"]["Number_of_samples"] = a["nb_samples"]
            rec["overall_only"] = astats_overall
        recs.append(rec)
    return dict(pcm=cur, rate=rate, meta=recs, loudnorm=ln)


def _close(a, b, atol, rtol=0.0):
    if isinstance(a, float) and math.isnan(a):
        return isinstance(b, float) and math.isnan(b)
    if math.isinf(a) or math.isinf(b):
        return a == b
    return abs(a - b) <= atol + rtol * abs(b)


ASTATS_TOL = {  # (atol, rtol) on the "%f"-printed values
    "Noise_floor_count": (1e9, 0.0),     # ties on equal window maxima are float-representation dependent
    "Dynamic_range": (0.02, 0.0),        # 20log10(2*peak / smallest non-zero |sample|): the smallest f64 sample of a
                                         # processed stream is itself rounding noise
    "Min_difference": (1e-9, 1e-6), "Flat_factor": (1e-3, 0.0),
    "Entropy": (2e-6, 0), "Zero_crossings": (0, 0), "Bit_depth": (0, 0), "Number_of_samples": (0, 0),
}


def assert_meta_close(got, exp, spectral_rtol=2e-3, spectral_atol=1e-9, astats_atol=2e-6, roundoff_only_below_lufs=None):
    """got: list of gpudsp.FrameMeta; exp: list of dicts from analysis_meta().
    roundoff_only_below_lufs: behind f32 stages, a sink frame whose momentary loudness is below this level (the
    start-up latency of afftdn / anlmdn: ~ -118 LUFS) holds nothing but those stages' float round-off, whose
    spectral SHAPE depends on the FFT's butterfly order; only its levels are compared."""
    assert len(got) == len(exp), (len(got), len(exp))
    # signed statistics (skewness, slope, decrease) cancel towards 0: scale their absolute tolerance
    # by the column's magnitude over the stream
    col_scale = [max([abs(e["spectral"][k]) for e in exp if math.isfinite(e["spectral"][k])] + [0.0]) for k in range(13)]
    # the rolloff (column 12) is the frequency of the BIN where the cumulative energy crosses 85 %: quantised to rate / win_size,
    # so round-off that moves the crossing by a hair moves the value by a whole bin.  Behind f32 stages whose noise floor depends
    # on where a lane starts, one bin either way is accepted in at most 1 % of the frames.
    rolls = sorted({e["spectral"][12] for e in exp if math.isfinite(e["spectral"][12])})
    roll_gaps = [b - a for a, b in zip(rolls, rolls[1:]) if b - a > 1e-9]
    roll_step = min(roll_gaps) if roll_gaps else 0.0
    roll_flips = 0
    for i, (g, e) in enumerate(zip(got, exp)):
        assert g.first_sample == e["first_sample"] and g.nb_samples == e["nb_samples"], (i, g.first_sample, e)
        for name, gv, ev in (("M", g.r128_M, e["M"]), ("S", g.r128_S, e["S"]), ("I", g.r128_I, e["I"]),
                             ("LRA", g.r128_LRA, e["LRA"])):
            assert _close(gv, ev, 0.0011), (i, name, gv, ev)
        for name, gv, ev in (("true_peak", g.r128_true_peak, e["true_peak"]),
                             ("sample_peak", g.r128_sample_peak, e["sample_peak"])):
            assert _close(gv, ev, 0.0011), (i, name, gv, ev)
        roundoff_only = (roundoff_only_below_lufs is not None and isinstance(e["M"], float) and math.isfinite(e["M"])
                         and e["M"] < roundoff_only_below_lufs)
        for k in range(13):
            if roundoff_only:
                assert math.isnan(g.spectral[k]) == math.isnan(e["spectral"][k]), (i, "spectral", k)
                continue
            ok = _close(g.spectral[k], e["spectral"][k], spectral_atol + 2e-4 * col_scale[k] + (2e-4 if k in (4, 5) else 0.0), spectral_rtol)
            if not ok and k == 12 and roll_step > 0 and abs(g.spectral[k] - e["spectral"][k]) <= roll_step * 1.01:       # ("%g" keeps six digits: the printed steps vary in the last one)
                roll_flips += 1
                ok = roll_flips <= max(1, len(exp) // 100)
            assert ok, (i, "spectral", k, g.spectral[k], e["spectral"][k], "rolloff step / flips", roll_step, roll_flips)
        if e["astats"] is None:
            assert all(math.isnan(g.astats[k]) for k in range(len(AS_NAMES))), (i, "unexpected astats")
        else:
            if not e.get("overall_only"):
                for k, name in enumerate(AS_NAMES):
                    atol, rtol = ASTATS_TOL.get(name, (astats_atol, 1e-9 if astats_atol < 1e-5 else 1e-4))
                    if astats_atol >= 1e-5 and name in ("Zero_crossings", "Zero_crossings_rate", "Entropy"):
                        atol, rtol = 1e-6, 1e-3       # samples hovering around zero flip sign with float round-off
                    if astats_atol >= 1e-5 and name == "Dynamic_range":
                        # behind afftdn the smallest non-zero |sample| of the stream is the inverse FFT's round-off in
                        # the start-up latency (1e-14 .. 1e-18 depending on butterfly order): only its presence is checked
                        assert math.isfinite(g.astats[k]) == math.isfinite(e["astats"][name]), (i, name)
                        continue
                    assert _close(g.astats[k], e["astats"][name], atol, rtol), (i, name, g.astats[k], e["astats"][name])
            assert _close(g.astats_overall_RMS_level, e["astats"]["RMS_level"], astats_atol), (i, "overall rms")
            assert _close(g.astats_overall_Peak_level, e["astats"]["Peak_level"], astats_atol), (i, "overall peak")
