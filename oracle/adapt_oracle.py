"""ORACLE (test infrastructure only -- never imported by the product): plain-Python restatement of the reference's host logic
between Pass 1 and Pass 2, used for seeded differential tests against libjtdsp's jt_adapt.cu.

Follows, function by function:
  analyser_noise_seed.go:78-241      roomToneScore, computeSilenceMedians, estimateNoiseFloorAndThreshold
  analyser_vad.go:38-783             histogram, Otsu split, clamp, hysteresis, gap tolerance, speech runs, low-cluster region,
                                     noise profile, gate statistics, floored fraction, detectVoiceActivity
  analyser_candidates_shared.go      refineToSubregion, getIntervalsInRange, accumulateIntervalMetrics, window scorers,
                                     levelVariance, measureSpeechCandidateFromIntervals
  analyser_candidates_speech.go      rolloff / flux / voicing scores, findBestSpeechRegion, scoreSpeechCandidateGrounded
  adaptive*.go, filters.go:607-989   AdaptConfig and the Pass-2 spec string
Pinned the same way as the C++ (the reference's unit-test tables run through both in tests/); durations are integer
nanoseconds, intervals are dicts {ts, rms, peak, M, S, tp, sp, spectral: {name: value}}.
"""
import bisect
import math

NS, MS = 1_000_000_000, 1_000_000
HOP = 250 * MS
INF, NAN = math.inf, math.nan
SP = ["mean", "variance", "centroid", "spread", "skewness", "kurtosis", "entropy", "flatness", "crest", "flux", "slope", "decrease", "rolloff"]


def gomax(a, b):
    return NAN if (a != a or b != b) else (a if a > b else b)


def gomin(a, b):
    return NAN if (a != a or b != b) else (a if a < b else b)


def gosorted(xs):
    return sorted(xs, key=lambda v: (0, 0.0) if v != v else (1, v))         # slices.Sort: NaN first


def seconds(d):
    return float(d // NS) + float(d % NS) / 1e9 if d >= 0 else -(float((-d) // NS) + float((-d) % NS) / 1e9)


def floored(l):
    return math.isinf(l) or l != l or l <= -115.0


def level(iv, axis=0):
    return iv["rms"] if axis == 1 else iv["M"]


def intervals_for(d, hop):
    return 0 if hop <= 0 else (d + hop // 2) // hop


def pct(sorted_vals, p):
    if not sorted_vals:
        return 0.0
    p = gomax(0.0, gomin(100.0, p))
    return sorted_vals[int(p / 100 * float(len(sorted_vals) - 1))]


def in_range(ivs, start, end):
    ts = [iv["ts"] for iv in ivs]
    lo = bisect.bisect_left(ts, start)
    hi = lo
    while hi < len(ivs) and ts[hi] < end:
        hi += 1
    return lo, hi


# ---- seed -----------------------------------------------------------------------------------------------------------
def estimate_noise_floor(ivs):
    if len(ivs) < 10:
        return 0.0, 0.0, False
    lv = gosorted([iv["M"] for iv in ivs])
    fx = gosorted([iv["spectral"]["flux"] for iv in ivs])
    l50, f50 = lv[len(lv) // 2], fx[len(fx) // 2]
    scored = []
    for i, iv in enumerate(ivs):
        amp = 1.0
        if iv["M"] > l50:
            amp = 1.0 - (iv["M"] - l50) / 6.0
            if amp < 0:
                amp = 0.0
        fl = 1.0
        if f50 > 0 and iv["spectral"]["flux"] > f50:
            ratio = iv["spectral"]["flux"] / f50
            if ratio > 1:
                fl = 1.0 / ratio
        scored.append((i, iv["M"], 0.6 * amp + 0.4 * fl))

    def key(t):
        i, l, s = t
        return ((0, 0.0) if s != s else (1, -s), (0, 0.0) if l != l else (1, l), i)      # score desc, level asc (NaN first), index
    # cmp.Compare(b.score, a.score): descending with NaN scores LAST (NaN compares below everything, so b NaN -> a first)
    scored.sort(key=lambda t: ((1, 0.0) if t[2] != t[2] else (0, -t[2]), (0, 0.0) if t[1] != t[1] else (1, t[1]), t[0]))
    cnt = min(max(len(scored) // 5, 8), len(scored))
    mx, seen = -120.0, False
    for i in range(cnt):
        l = scored[i][1]
        if floored(l):
            continue
        if not seen or l > mx:
            mx, seen = l, True
    if not seen:
        return 0.0, 0.0, False
    return mx, mx + 1.0, True


# ---- histogram / split ----------------------------------------------------------------------------------------------
def build_hist(ivs, axis, bw):
    lv = [level(iv, axis) for iv in ivs if not floored(level(iv, axis))]
    if bw <= 0 or not lv:
        return dict(bins=[], bw=0.0, lo=0.0, hi=0.0, count=0)
    lo, hi = min(lv), max(lv)
    nb = int((hi - lo) / bw) + 1
    bins = [0] * nb
    for l in lv:
        bins[min(int((l - lo) / bw), nb - 1)] += 1
    return dict(bins=bins, bw=bw, lo=lo, hi=hi, count=len(lv))


def centre(h, i):
    return h["lo"] + (float(i) + 0.5) * h["bw"]


def otsu(h):
    nb = len(h["bins"])
    if h["count"] == 0 or nb < 2:
        return (h["lo"] + h["hi"]) / 2
    total = float(h["count"])
    sum_all = 0.0
    for i, c in enumerate(h["bins"]):
        sum_all += centre(h, i) * float(c)
    wb = sb = best = 0.0
    best_i = -1
    for i in range(nb - 1):
        wb += float(h["bins"][i])
        sb += centre(h, i) * float(h["bins"][i])
        wf = total - wb
        if wb == 0 or wf == 0:
            continue
        d = sb / wb - (sum_all - sb) / wf
        var = wb * wf * d * d
        if var > best:
            best, best_i = var, i
    if best_i < 0:
        return (h["lo"] + h["hi"]) / 2
    return h["lo"] + float(best_i + 1) * h["bw"]


def hysteresis_margin(h, split):
    w = c = 0.0
    for i, n in enumerate(h["bins"]):
        ce = centre(h, i)
        if ce >= split:
            w += ce * float(n)
            c += float(n)
    upper = split if c == 0 else w / c
    dist = upper - split
    return 1.0 if dist <= 0 else dist * 0.25


def clamp_split(split, floor, p75):
    lower = floor + 2.0
    if p75 < lower:
        return lower
    return gomax(lower, gomin(p75, split))


def veto_ok(iv):
    s = iv["spectral"]
    return 200.0 <= s["centroid"] <= 6000.0 and s["entropy"] < 0.70


def is_speech(iv, split, axis=0):
    return level(iv, axis) >= split and veto_ok(iv)


def gap_tolerance(flags, hop):
    fl, ce = intervals_for(2 * NS, hop), intervals_for(10 * NS, hop)
    idx = [i for i, f in enumerate(flags) if f]
    if not idx:
        return fl
    gaps, g = [], 0
    for i in range(idx[0], idx[-1] + 1):
        if flags[i]:
            if g > 0:
                gaps.append(float(g))
            g = 0
        else:
            g += 1
    if not gaps:
        return fl
    p75 = int(math.floor(pct(sorted(gaps), 75) + 0.5))
    return max(fl, min(ce, p75))


def speech_runs(ivs, split, margin, tol, axis, hop):
    min_iv = intervals_for(10 * NS, hop)
    if len(ivs) < min_iv or min_iv <= 0:
        return []
    high, low = split + margin, split - margin
    runs = []
    st = dict(in_run=False, start=0, count=0, last=0, pending=0)

    def flush(end_idx):
        if st["in_run"] and st["count"] >= min_iv:
            e = ivs[end_idx]["ts"] + hop
            runs.append((st["start"], e))
        st.update(in_run=False, count=0, pending=0)
    for i, iv in enumerate(ivs):
        l, ok = level(iv, axis), veto_ok(iv)
        if not st["in_run"]:
            if l >= high and ok:
                st.update(in_run=True, start=iv["ts"], count=1, last=i, pending=0)
            continue
        if l >= split and ok:
            st["count"] += 1
            st["last"] = i
            st["pending"] = 0
            continue
        if l >= split and not ok:
            flush(st["last"])
            continue
        if l < low:
            st["pending"] += 1
            if st["pending"] > tol:
                flush(st["last"])
    flush(st["last"])
    return runs


# ---- regions --------------------------------------------------------------------------------------------------------
def accumulate(ivs):
    a = dict(rms=0.0, peak=-120.0, tp=-120.0, sp=-120.0, spec={k: 0.0 for k in SP}, M=0.0, S=0.0)
    for iv in ivs:
        a["rms"] += iv["rms"]
        if iv["peak"] > a["peak"]:
            a["peak"] = iv["peak"]
        for k in SP:
            a["spec"][k] += iv["spectral"][k]
        a["M"] += iv["M"]
        a["S"] += iv["S"]
        if iv["tp"] > a["tp"]:
            a["tp"] = iv["tp"]
        if iv["sp"] > a["sp"]:
            a["sp"] = iv["sp"]
    return a


def region_sample(ivs):
    a, n = accumulate(ivs), float(len(ivs))
    rms = a["rms"] / n
    return dict(rms=rms, peak=a["peak"], crest=a["peak"] - rms, spectral={k: a["spec"][k] / n for k in SP}, M=a["M"] / n, S=a["S"] / n,
                tp=a["tp"], sp=a["sp"])


def score_interval_window(ivs):
    return sum_seq(iv["rms"] for iv in ivs) / float(len(ivs)) if ivs else 0.0


def sum_seq(it):
    s = 0.0
    for v in it:
        s += v
    return s


def rolloff_score(r):
    if 4000.0 <= r <= 8000.0:
        return 1.0
    if 2500.0 <= r < 4000.0:
        return 0.5 + 0.5 * (r - 2500.0) / (4000.0 - 2500.0)
    if 8000.0 < r <= 10000.0:
        return 0.5 + 0.5 * (10000.0 - r) / (10000.0 - 8000.0)
    return 0.0


def flux_score(f):
    if f <= 0.004:
        return 1.0
    if f <= 0.010:
        return 1.0 - (f - 0.004) / (0.010 - 0.004) * 0.3
    if f <= 0.020:
        return 0.7 - (f - 0.010) / (0.020 - 0.010) * 0.3
    if f <= 0.030:
        return 0.4 - (f - 0.020) / (0.030 - 0.020) * 0.2
    return 0.2


def score_speech_window(ivs):
    if not ivs:
        return 0.0
    n = float(len(ivs))
    g = lambda k: sum_seq(iv["spectral"][k] for iv in ivs) / n
    aku, afl, ace, aro, afx = g("kurtosis"), g("flatness"), g("centroid"), g("rolloff"), g("flux")
    arm = sum_seq(iv["rms"] for iv in ivs) / n
    var = sum_seq((iv["spectral"]["kurtosis"] - aku) ** 2 for iv in ivs) / n
    voiced = sum(1 for iv in ivs if iv["spectral"]["kurtosis"] > 4.5)
    voicing = gomax(0.0, gomin((float(voiced) / n) / 0.6, 1.0))
    s_ku = gomax(0.0, gomin(aku / 7.5, 1.0))
    s_fl = gomax(0.0, gomin(1.0 - afl, 1.0))
    s_ce = 0.0
    if 200.0 <= ace <= 6000.0:
        s_ce = 1.0 - (abs(ace - 3100.0) / 2900.0) * 0.5
    s_co = gomax(0.0, gomin(1.0 - (var / 100.0), 1.0))
    s_rm = gomax(0.0, gomin((arm - (-30.0)) / 18.0, 1.0)) if arm > -30.0 else 0.0
    return (s_ku * 0.15 + s_fl * 0.10 + s_ce * 0.10 + s_co * 0.10 + s_rm * 0.10 + voicing * 0.15 + rolloff_score(aro) * 0.15 +
            flux_score(afx) * 0.15)


def level_variance(ivs, axis=0):
    if not ivs:
        return 0.0
    n = float(len(ivs))
    mean = sum_seq(level(iv, axis) for iv in ivs) / n
    return sum_seq((level(iv, axis) - mean) ** 2 for iv in ivs) / n


def refine(ivs, region, window, minimum, score, higher_wins):
    start, end = region
    if end - start <= window:
        return region, False
    lo, hi = in_range(ivs, start, end)
    if hi == lo:
        return region, False
    w, mn = window // HOP, minimum // HOP
    if hi - lo < mn:
        return region, False
    w = min(w, hi - lo)
    best, best_score = 0, score(ivs[lo:lo + w])
    for s in range(1, hi - lo - w + 1):
        sc = score(ivs[lo + s:lo + s + w])
        if (sc > best_score) if higher_wins else (sc < best_score):
            best, best_score = s, sc
    st = ivs[lo + best]["ts"]
    return (st, st + w * HOP), True


def grounded_score(rms, duration, floor_db, var):
    snr = rms - floor_db
    if snr <= 0:
        s = 0.0
    elif snr < 20.0:
        s = 0.5 * (snr / 20.0)
    elif snr >= 40.0:
        s = 1.0
    else:
        s = 0.5 + 0.5 * (snr - 20.0) / (40.0 - 20.0)
    d = 1.0 if duration >= 30 * NS else gomax(0.0, gomin(seconds(duration) / seconds(30 * NS), 1.0))
    tie = gomax(0.0, gomin(1.0 - (var / 25.0), 1.0)) * 0.02
    return s * 0.6 + d * 0.4 + tie


def measure_candidate(ivs, region):
    lo, hi = in_range(ivs, *region)
    if hi == lo:
        return None
    sub = ivs[lo:hi]
    c = region_sample(sub)
    c.update(region=region, voicing=float(sum(1 for iv in sub if iv["spectral"]["kurtosis"] > 4.5)) / float(len(sub)),
             refined=False, orig=None, score=0.0)
    return c


def find_best(ivs, regions, floor_db):
    cands, best, best_score, fb, fb_score = [], None, 0.0, None, 0.0
    for reg in regions:
        c = measure_candidate(ivs, reg)
        if c is None:
            continue
        lo, hi = in_range(ivs, *reg)
        c["score"] = grounded_score(c["rms"], reg[1] - reg[0], floor_db, level_variance(ivs[lo:hi], 0))
        cands.append(c)
        if fb is None or c["score"] > fb_score:
            fb, fb_score = reg, c["score"]
        if c["score"] >= 0.3 and (best is None or c["score"] > best_score):
            best, best_score = reg, c["score"]
    if best is None:
        best = fb
    if best is not None and best[1] - best[0] > 60 * NS:
        orig = best
        refined, _ = refine(ivs, best, 60 * NS, 30 * NS, score_speech_window, True)
        if refined != orig:
            rc = measure_candidate(ivs, refined)
            if rc is not None:
                lo, hi = in_range(ivs, *refined)
                rc["score"] = grounded_score(rc["rms"], refined[1] - refined[0], floor_db, level_variance(ivs[lo:hi], 0))
                rc.update(refined=True, orig=orig)
                for i, c in enumerate(cands):
                    if c["region"][0] == orig[0]:
                        cands[i] = rc
                        break
                best = refined
    return best, cands


def low_cluster_region(ivs, split, axis, hop):
    best, run_start, in_run = None, 0, False

    def close(end_idx):
        nonlocal best, in_run
        e = ivs[end_idx]["ts"] + hop
        if best is None or e - run_start > best[1] - best[0]:
            best = (run_start, e)
        in_run = False
    for i, iv in enumerate(ivs):
        if level(iv, axis) < split:
            if not in_run:
                run_start, in_run = iv["ts"], True
            continue
        if in_run:
            close(i - 1)
    if in_run:
        close(len(ivs) - 1)
    if best is None:
        return None
    refined, _ = refine(ivs, best, 10 * NS, 8 * NS, score_interval_window, False)
    return refined


def gate_statistics(ivs, split, axis, speech_region):
    noise = [level(iv, axis) for iv in ivs if not floored(level(iv, axis)) and level(iv, axis) < split]
    voiced = []
    if speech_region is not None:
        lo, hi = in_range(ivs, *speech_region)
        voiced = [level(iv, axis) for iv in ivs[lo:hi] if is_speech(iv, split, axis)]
    v, n = pct(gosorted(voiced), 10.0), pct(gosorted(noise), 95.0)
    return v, n, v - n


def floored_fraction(ivs, axis=0):
    if not ivs:
        return 0.0
    return float(sum(1 for iv in ivs if level(iv, axis) != level(iv, axis) or level(iv, axis) <= -115.0)) / float(len(ivs))


def detect(ivs, seed):
    """detectVoiceActivity (analyser_vad.go:728-783)"""
    axis, hop = 0, HOP
    h = build_hist(ivs, axis, 1.0)
    levels = gosorted([level(iv, axis) for iv in ivs if not floored(level(iv, axis))])
    split = clamp_split(otsu(h), seed, pct(levels, 75))
    floor = gomax(pct(levels, 10.0), seed + 2.0)
    flags = [is_speech(iv, split, axis) for iv in ivs]
    margin = hysteresis_margin(h, split)
    tol = gap_tolerance(flags, hop)
    runs = speech_runs(ivs, split, margin, tol, axis, hop)
    out = dict(split=split, floor=floor, margin=margin, tol=tol, runs=runs, noise_region=None, noise_profile=None, room_tone=None)
    nreg = low_cluster_region(ivs, split, axis, hop)
    if nreg is not None:
        lo, hi = in_range(ivs, nreg[0], nreg[0] + (nreg[1] - nreg[0]))
        if hi > lo:
            s = region_sample(ivs[lo:hi])
            out["noise_region"] = nreg
            out["noise_profile"] = dict(start=nreg[0], duration=nreg[1] - nreg[0], floor=floor, peak=s["peak"], crest=s["crest"],
                                        entropy=s["spectral"]["entropy"], spectral=s["spectral"])
            out["room_tone"] = s
    best, cands = find_best(ivs, runs, floor if out["noise_profile"] is not None else -INF)
    out["cands"] = cands
    out["speech"] = next((c for c in cands if best is not None and c["region"][0] == best[0]), None)
    out["voiced_low"], out["noise_high"], out["separation"] = gate_statistics(ivs, split, axis, out["speech"]["region"] if out["speech"] else None)
    out["floored_fraction"] = floored_fraction(ivs, axis)
    out["voice_activated"] = out["floored_fraction"] >= 0.20
    return out


def detect_full(ivs):
    seed, thr, ok = estimate_noise_floor(ivs)
    if not ok:
        seed, thr = -115.0, max(-70.0, min(-35.0, -115.0 + 6.0))
    out = detect(ivs, seed)
    out.update(prescan=seed, detect_level=thr)
    return out


# ---- AdaptConfig + spec (adaptive*.go, filters.go) ---------------------------------------------------------------------
def go_g(v):
    """fmt %g: shortest round-trip digits"""
    if v == 0:
        return "0"
    r = repr(float(v))
    if "e" in r or "E" in r:
        mant, _, ex = r.lower().partition("e")
        ex = int(ex)
    else:
        mant, ex = r, None
    if ex is None:
        digits = mant.lstrip("-").replace(".", "").lstrip("0") or "0"
        x = int(math.floor(math.log10(abs(v))))
        if -4 <= x < 6:
            if mant.endswith(".0"):
                mant = mant[:-2]
            return mant
        d = digits.rstrip("0") or "0"
        return ("-" if v < 0 else "") + d[0] + ("." + d[1:] if len(d) > 1 else "") + "e%s%02d" % ("-" if x < 0 else "+", abs(x))
    if mant.endswith(".0"):
        mant = mant[:-2]
    if -4 <= ex < 6:
        return repr(float(v))
    return mant + "e%s%02d" % ("-" if ex < 0 else "+", abs(ex))


def db_lin(db):
    return math.pow(10.0, db / 20.0)


def band_noise(bands):
    fin = [b for b in bands if math.isfinite(b)]
    if not bands or not fin:
        return ""
    mean = sum_seq(fin) / float(len(fin))
    return "|".join("0.0" if not math.isfinite(b) else "%.1f" % gomax(-24.0, gomin(24.0, b - mean)) for b in bands)


def adapt_spec(meas, va, speech_bands=None, noise_bands=None):
    """AdaptConfig on DefaultFilterConfig + BuildFilterSpec.  meas: input_i, input_lra, rms_level, peak_level (Go zero values when
    unmeasured); va: detect_full() output; speech_bands (body, sib) or None; noise_bands: 15 values or None."""
    # tuneNoiseReduction
    afftdn = "afftdn=nr=12:nt=w:tn=1"
    if va["voice_activated"]:
        afftdn = ""
    elif va["floor"] != 0:
        fl = gomax(-80.0, gomin(-20.0, va["floor"]))
        nt = "nt=w"
        finite = [b for b in (noise_bands or []) if math.isfinite(b)]
        measured = noise_bands is not None and len(finite) >= 10
        if va["noise_profile"] is not None and measured and not (va["separation"] < 12.0) and va["noise_profile"]["spectral"]["flatness"] >= 0.45:
            bn = band_noise(noise_bands)
            if bn:
                nt = "nt=custom:bn=" + bn
        afftdn = "afftdn=nr=12:%s:tn=0" % nt + (":nf=" + go_g(fl) if fl < 0 else "")
    # tuneSpeechGate
    ratio = 1.5 if meas["input_lra"] > 15.0 else 2.0
    narrow = False
    if va["speech"] is not None:
        t = gomax(-80.0, gomin(va["voiced_low"] - 6.0, -25.0))
        narrow = va["separation"] < 12.0
    else:
        crest, peak = (va["noise_profile"]["crest"], va["noise_profile"]["peak"]) if va["noise_profile"] is not None else (15.0, 0.0)
        gap = max(-16.0 - meas["input_i"], 0.0)
        if crest > 20.0 and peak != 0 and gap < 25.0:
            t = peak + 3.0
        else:
            t = gomax(va["floor"] + 12.0 / (1.0 - 1.0 / ratio), -40.0)
        t = gomax(-80.0, gomin(t, -25.0))
    thr = db_lin(t)
    if not math.isfinite(thr) or thr <= 0:
        thr = 0.01
    rng = db_lin(-(8.0 if narrow else 14.0))
    # tuneDeesser
    inten = 0.0
    if va["speech"] is not None and speech_bands is not None:
        ex = speech_bands[1] - speech_bands[0]
        if ex < -6.0:
            inten = 0.0
        elif ex < -3.0:
            inten = (ex + 6.0) / 3.0 * 0.6
        elif ex < 0.0:
            inten = 0.6 + (ex + 3.0) / 3.0 * (0.85 - 0.6)
        else:
            inten = 0.85
    # tuneLevellingCompressor
    if va["speech"] is not None:
        rms = va["speech"]["rms"]
        if meas["rms_level"] < 0 and not (math.isinf(meas["rms_level"]) and meas["rms_level"] < 0):
            rms = gomax(rms, meas["rms_level"])
        ct = gomax(-45.0, gomin(rms + 9.0, -6.0))
    elif not math.isfinite(meas["peak_level"]):
        ct = -18.0
    else:
        ct = gomax(-45.0, gomin(meas["peak_level"] - 20.0, -6.0))
    if not math.isfinite(ct):
        ct = -18.0
    parts = ["aformat=channel_layouts=mono",
             "highpass=f=80:poles=2:width_type=q:width=0.707:normalize=1:a=tdii",
             "lowpass=f=20500:poles=2:width_type=q:width=0.707:normalize=1:a=tdii",
             "anlmdn=s=0.00001:p=0.0060:r=0.0020:m=3" + ("," + afftdn if afftdn else ""),
             "agate=threshold=%.6f:ratio=%.1f:attack=5.00:release=200:range=%.4f:knee=3.0:detection=rms:makeup=1.0" % (thr, ratio, rng),
             "acompressor=threshold=%.6f:ratio=3.0:attack=10:release=200:makeup=1.00:knee=4.0:detection=rms:mix=1.00" % db_lin(ct)]
    if inten > 0:
        parts.append("deesser=i=%.2f:m=0.50:f=0.80" % inten)
    parts += ["astats=metadata=1:measure_perchannel=all,aspectralstats=win_size=2048:win_func=hann:measure=all,"
              "ebur128=metadata=1:peak=sample+true:dualmono=true:target=-16",
              "aformat=sample_rates=44100:channel_layouts=mono:sample_fmts=s16,asetnsamples=n=4096"]
    return ",".join(parts)
