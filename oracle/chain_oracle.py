"""ORACLE (test infrastructure only -- never imported by the product): the reference's ProcessAudio (processor.go:78-216) composed
entirely from oracle parts, so that tests/, smoke() and bench.py's CPU legs can run the whole path without touching libjtdsp:

    Pass 1   collectAnalysisFrames            oracle_graph.pass1_analyse           analyser.go:538-650
    detector detectVoiceActivity + election   adapt_oracle.detect_full             analyser_vad.go:728-783, analyser_candidates_*.go
    bands    measureSpeechBandRMS x 17        orc_biquad + orc_astats              analyser_bands.go:33-167, analyser_noise_bands.go:15-119
    adapt    AdaptConfig + BuildFilterSpec    adapt_oracle.adapt_spec              adaptive*.go, filters.go:968-989
    Pass 2   processWithFilters               oracle_graph.run_spec                processor.go:255-373
    regions  MeasureOutputRegions             oracle_graph.region_sample           analyser_output.go:95-297
    Pass 3   measureWithLoudnorm + planners   plan_pass3 / run_spec                normalise.go:226-346, 373-425, 539-561
    Pass 4   applyLoudnormAndMeasure          plan_pass4 / run_spec                normalise.go:583-632, 924-957, 1198-1334
"""
import math

import numpy as np

import adapt_oracle as AO
import jt_oracle as O
import oracle_graph as OG

PASS1_SPEC = ("aformat=channel_layouts=mono,astats=metadata=1:measure_perchannel=all,"
              "aspectralstats=win_size=2048:win_func=hann:measure=all,"
              "ebur128=metadata=1:peak=sample+true:dualmono=true:target=-16")                       # filters.go:42-45, 623-627
DEFAULT_PASS2_SPEC = ("aformat=channel_layouts=mono,"
                      "highpass=f=80:poles=2:width_type=q:width=0.707:normalize=1:a=tdii,"
                      "lowpass=f=20500:poles=2:width_type=q:width=0.707:normalize=1:a=tdii,"
                      "anlmdn=s=0.00001:p=0.0060:r=0.0020:m=3,afftdn=nr=12:nt=w:tn=1,"
                      "agate=threshold=0.010000:ratio=2.0:attack=5.00:release=200:range=0.1995:knee=3.0:detection=rms:makeup=1.0,"
                      "acompressor=threshold=0.125893:ratio=3.0:attack=10:release=200:makeup=1.00:knee=4.0:detection=rms:mix=1.00,"
                      "astats=metadata=1:measure_perchannel=all,aspectralstats=win_size=2048:win_func=hann:measure=all,"
                      "ebur128=metadata=1:peak=sample+true:dualmono=true:target=-16,"
                      "aformat=sample_rates=44100:channel_layouts=mono:sample_fmts=s16,asetnsamples=n=4096")   # DefaultFilterConfig, filters.go:353-355

# normalise.go:30-60
MIN_LIMITER_CEILING_DB, BRICKWALL_HEADROOM_DB, CUSHION_DB, LINEAR_SAFETY = -24.0, 0.9, 0.2, 0.1
TP_MAX, TP_MIN = 0.0, -9.0

AFFTDN_CENTRES = [80, 125, 195, 290, 440, 660, 1000, 1500, 2250, 3350, 5000, 7500, 11200, 16000, 24000]   # analyser_noise_bands.go:15-17


def band_plan():
    """speechBandPlan (analyser_bands.go:98-103) then afftdnBandEdgesHz (analyser_noise_bands.go:34-52): 17 (lo, hi) pairs"""
    c = AFFTDN_CENTRES
    lo, hi = [1000.0, 6000.0], [3000.0, 9000.0]
    for i in range(15):
        lo.append(c[0] / math.sqrt(c[1] / c[0]) if i == 0 else math.sqrt(c[i - 1] * c[i]))
        hi.append(c[14] * math.sqrt(c[14] / c[13]) if i == 14 else math.sqrt(c[i] * c[i + 1]))
    return lo, hi


def band_rms(mono, rate, start_ns, dur_ns, lo, hi):
    """measureSpeechBandRMS (analyser_bands.go:33-104): atrim=start=%f:duration=%f, highpass, lowpass, astats Overall RMS"""
    st, du = float("%f" % (start_ns / 1e9)), float("%f" % (dur_ns / 1e9))
    s0 = (round(st * 1e6) * rate + 500000) // 1000000
    n = (round(du * 1e6) * rate + 500000) // 1000000
    reg = mono[s0:s0 + n]
    return [float("%f" % O.astats(O.biquad(O.biquad(reg, rate, "highpass", l), rate, "lowpass", h), rate)["RMS_level"]) for l, h in zip(lo, hi)]


def db_lin(db):
    return math.pow(10.0, db / 20.0)


def pre_limiter_prefix(pre_gain, ceiling, needed):
    """buildPreLimiterPrefix (normalise.go:446-480)"""
    if not needed:
        return ""
    s = "volume=%.1fdB," % pre_gain if pre_gain > 0 else ""
    return s + "alimiter=limit=%.6f:attack=5:release=100:level_in=1:level_out=1:level=0:latency=1:asc=1:asc_level=0.8" % db_lin(ceiling)


def plan_pass3(out_i, out_tp, t_i=-16.0, t_tp=-1.0, t_lra=20.0):
    """calculateLimiterCeiling, calculatePreGain, planLimiterForLoudnorm (normalise.go:373-425, 539-561) and the Pass-3 spec
    (normalise.go:257-264) -> (spec, plan)"""
    gain = t_i - out_i
    projected = out_tp + gain
    ceiling, needed, clamped = 0.0, False, False
    if projected > t_tp:
        ceiling, needed = t_tp - gain, True
        if ceiling < MIN_LIMITER_CEILING_DB:
            ceiling, clamped = MIN_LIMITER_CEILING_DB, True
    pre_gain, re_ceil = 0.0, 0.0
    ideal = t_tp - gain
    if ideal < MIN_LIMITER_CEILING_DB:
        pre_gain = MIN_LIMITER_CEILING_DB - ideal
        re_ceil = t_tp - (t_i - (out_i + pre_gain))
    if clamped:
        ceiling = re_ceil
    prefix = pre_limiter_prefix(pre_gain, ceiling, needed)
    ln = "loudnorm=I=%.1f:TP=%.1f:LRA=%.1f:dual_mono=true:print_format=json" % (t_i, t_tp, t_lra)
    return (prefix + "," + ln if prefix else ln), dict(ceiling=ceiling, pre_gain=pre_gain, gain=gain, needed=needed, clamped=clamped)


def plan_pass4(plan, p3, t_i=-16.0, t_tp=-1.0, t_lra=20.0, source_rate=44100):
    """loudnormInternalTargetTP, calculateLinearModeTarget, loudnormTPTargets, buildLoudnormFilterSpec
    (normalise.go:583-632, 1198-1203, 1231-1334); p3: the Pass-3 loudnorm JSON as numbers -> (spec, effective target, offset)"""
    m_i, m_tp = float("%.2f" % p3["input_i"]), float("%.2f" % p3["input_tp"])
    m_lra, m_th = float("%.2f" % p3["input_lra"]), float("%.2f" % p3["input_thresh"])
    internal_tp = m_tp + (t_i - m_i) + LINEAR_SAFETY + CUSHION_DB
    max_linear = internal_tp - m_tp + m_i - LINEAR_SAFETY
    eff = t_i if t_i <= max_linear else max_linear
    offset = eff - m_i
    emitted_tp = max(TP_MIN, min(internal_tp, TP_MAX))
    brickwall = t_tp - BRICKWALL_HEADROOM_DB
    spec = pre_limiter_prefix(plan["pre_gain"], plan["ceiling"], plan["needed"])
    if spec:
        spec += ","
    spec += ("loudnorm=I=%.2f:TP=%.2f:LRA=%.1f:measured_I=%.2f:measured_TP=%.2f:measured_LRA=%.2f:measured_thresh=%.2f:offset=%.2f:"
             "dual_mono=true:linear=true:print_format=json" % (eff, emitted_tp, t_lra, m_i, m_tp, m_lra, m_th, offset))
    if source_rate > 0:
        spec += ",aresample=%d" % source_rate
    spec += ",adeclick=t=1.7:w=55:o=50:m=s"
    spec += ",alimiter=limit=%.6f:attack=1:release=50:level_in=1:level_out=1:level=0:latency=1:asc=1:asc_level=0.8" % db_lin(brickwall)
    spec += (",astats=metadata=1:measure_perchannel=all,aspectralstats=win_size=2048:win_func=hann:measure=all,"
             "ebur128=metadata=1:peak=sample+true:dualmono=true,"
             "aformat=sample_rates=44100:channel_layouts=mono:sample_fmts=s16,asetnsamples=n=4096")
    return spec, eff, offset


def to_adapt_intervals(oiv):
    """oracle_graph.pass1_analyse intervals -> adapt_oracle's dicts"""
    return [dict(ts=d["ts_ns"], rms=d["rms"], peak=d["pk"], M=d["M"], S=d["S"], tp=d["tp"], sp=d["sp"],
                 spectral={name: d["spectral"][k] for k, name in enumerate(AO.SP)}) for d in oiv]


def analyse_adaptive(x, rate, channels=1):
    """AnalyseAudio + AdaptConfig (analyser.go:325-372, processor.go:37-69) -> dict(meas, intervals, va, speech_bands, noise_bands, spec)"""
    meas, oiv = OG.pass1_analyse(x, rate, channels)
    ivs = to_adapt_intervals(oiv)
    va = AO.detect_full(ivs)
    mono = OG.downmix(x, channels)
    lo, hi = band_plan()
    speech_bands = noise_bands = None
    if va["speech"] is not None:
        r = va["speech"]["region"]
        if r[1] - r[0] > 0:
            speech_bands = band_rms(mono, rate, r[0], r[1] - r[0], lo[:2], hi[:2])
    if va["noise_profile"] is not None and va["noise_profile"]["duration"] > 0:
        p = va["noise_profile"]
        noise_bands = band_rms(mono, rate, p["start"], p["duration"], lo[2:], hi[2:])
    a = meas["astats"] or {}
    m = dict(input_i=meas["input_i"], input_lra=meas["input_lra"], rms_level=a.get("RMS_level", 0.0), peak_level=a.get("Peak_level", 0.0))
    spec = AO.adapt_spec(m, va, speech_bands, noise_bands)
    return dict(meas=meas, intervals=oiv, va=va, speech_bands=speech_bands, noise_bands=noise_bands, spec=spec)


def last_loudness(meta):
    last = [m for m in meta if not math.isnan(m["I"])][-1]
    return last["I"], (-120.0 if last["true_peak"] <= 0 else 20 * math.log10(last["true_peak"])), last["LRA"]


def output_regions(pcm, va):
    """MeasureOutputRegions (analyser_output.go:261-297) on a pass's s16 44.1 kHz output"""
    out = dict(room_tone=None, speech=None)
    if va["noise_profile"] is not None:
        try:
            out["room_tone"] = OG.region_sample(pcm, 44100, va["noise_profile"]["start"], va["noise_profile"]["duration"])[0]
        except Exception:      # noqa: BLE001  (a failing region is a warning in the reference, its sample stays absent)
            pass
    if va["speech"] is not None:
        r = va["speech"]["region"]
        try:
            out["speech"] = OG.region_sample(pcm, 44100, r[0], r[1] - r[0])[0]
        except Exception:      # noqa: BLE001
            pass
    return out


def process_audio(x, rate, channels=1, adaptive=True, regions=True, pass2_spec=None):
    """ProcessAudio (processor.go:78-216) -> dict(pcm, spec2, spec3, spec4, p3, p4, filtered=(I, TP, LRA), final=(I, TP, LRA), ...)"""
    res = {}
    if adaptive and pass2_spec is None:
        an = analyse_adaptive(x, rate, channels)
        res["analysis"] = an
        pass2_spec = an["spec"]
    else:
        OG.run_spec(PASS1_SPEC, x, rate, channels, want_pcm=False)
        pass2_spec = pass2_spec or DEFAULT_PASS2_SPEC
    p2 = OG.run_spec(pass2_spec, x, rate, channels)
    res["spec2"], res["pass2"] = pass2_spec, p2
    res["filtered"] = last_loudness(p2["meta"])
    if adaptive and regions and "analysis" in res:
        res["filtered_regions"] = output_regions(p2["pcm"], res["analysis"]["va"])
    spec3, plan = plan_pass3(res["filtered"][0], res["filtered"][1])
    p3 = OG.run_spec(spec3, p2["pcm"], 44100, want_pcm=False)
    res["spec3"], res["plan"], res["p3"] = spec3, plan, p3["loudnorm"]
    spec4, eff, off = plan_pass4(plan, p3["loudnorm"])
    p4 = OG.run_spec(spec4, p2["pcm"], 44100)
    res["spec4"], res["p4"], res["pass4"], res["pcm"] = spec4, p4["loudnorm"], p4, p4["pcm"]
    res["final"] = last_loudness(p4["meta"])
    if adaptive and regions and "analysis" in res:
        res["final_regions"] = output_regions(p4["pcm"], res["analysis"]["va"])
    return res
