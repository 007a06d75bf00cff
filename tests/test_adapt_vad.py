"""Voice-activity detector / region election (csrc/jt_adapt.cu) against the reference's own unit-test tables.

Every test here replays a case of /root/reference/internal/processor/analyser_vad_test.go (cited per test) through the
C ABI: same synthetic interval streams, same expected values.  Host-only -- runs without a GPU.
"""
import math

import pytest

from jivetalking_b200 import adapt as A

HOP = A.HOP_NS
S = A.NS_S
AX = A.AXIS_MOMENTARY
LEVEL_FLOOR = -115.0            # vadLevelFloorDB analyser_vad.go:57
NOISE_MARGIN = 2.0              # speechMinimumNoiseMarginDB analyser_candidates_speech.go:21
VA_FRACTION = 0.20              # vadVoiceActivatedFraction analyser_vad.go:673


def vad_interval(idx, lufs):    # analyser_vad_test.go:11-22
    return A.interval(idx * HOP, rms=lufs, momentary=lufs, centroid=2000.0, entropy=0.40)


def vad_speech(idx):
    return vad_interval(idx, -15)


def vad_quiet(idx):
    return vad_interval(idx, -60)


def vad_loud_non_speech(idx):   # analyser_vad_test.go:553-558
    s = vad_interval(idx, -15)
    s.spectral[A.SP_NAMES.index("centroid")] = 9000
    return s


def vad_speech_rich_at(idx, rms):   # analyser_vad_test.go:769-778
    s = vad_interval(idx, rms)
    s.rms_level, s.peak_level = rms, rms + 12
    for k, v in (("kurtosis", 6.0), ("rolloff", 6000.0), ("flux", 0.004), ("flatness", 0.2)):
        s.spectral[A.SP_NAMES.index(k)] = v
    return s


def vad_speech_rich(idx):
    return vad_speech_rich_at(idx, -16.0)


def seed_interval(level, flux):     # analyser_vad_test.go:339-345
    return A.interval(0, rms=level, momentary=level, flux=flux)


@pytest.mark.parametrize("d,hop,want", [(10 * S, 250 * A.NS_MS, 40), (2 * S, 250 * A.NS_MS, 8), (2 * S, 100 * A.NS_MS, 20),
                                        (10 * S, 100 * A.NS_MS, 100), (10 * S, 0, 0)])
def test_intervals_for_duration(d, hop, want):       # analyser_vad_test.go:24-44
    assert A.intervals_for_duration(d, hop) == want


def test_build_level_histogram():                     # analyser_vad_test.go:46-105
    iv, idx = [], 0
    for i in range(30):
        iv.append(vad_interval(idx, -50 + float(i % 3))); idx += 1
    for i in range(30):
        iv.append(vad_interval(idx, -20 + float(i % 3))); idx += 1
    iv.append(vad_interval(idx, -130))
    h = A.build_level_histogram(iv, AX, 2.0)
    assert h.count == 60
    assert sum(h.bins) == h.count
    low = sum(c for i, c in enumerate(h.bins) if h.bin_centre(i) < -40)
    high = sum(c for i, c in enumerate(h.bins) if h.bin_centre(i) > -30)
    valley = h.count - low - high
    assert low and high and valley == 0
    rms_iv = [A.interval(0, rms=-60, momentary=-10), A.interval(0, rms=-58, momentary=-10)]
    hr = A.build_level_histogram(rms_iv, A.AXIS_RMS, 2.0)
    assert hr.max_level <= -50


def test_otsu_split():                                # analyser_vad_test.go:107-163
    iv = [vad_interval(i, -50 + float(i % 2)) for i in range(40)] + [vad_interval(40 + i, -18 + float(i % 2)) for i in range(40)]
    split = A.otsu_split(A.build_level_histogram(iv, AX, 1.0))
    assert -49 < split < -18
    # single mode stays within clamp bounds
    iv = [vad_interval(i, -18 + float(i % 2)) for i in range(80)]
    h = A.build_level_histogram(iv, AX, 1.0)
    p75 = A.percentile_of_sorted(A.vad_levels(iv, AX), 75)
    split = A.clamp_split(A.otsu_split(h), -60.0, p75)
    assert -60.0 + NOISE_MARGIN - 0.001 <= split <= p75 + 0.001
    # degenerate low split pinned to the lower bound
    iv = [vad_interval(i, -50 + float(i % 2)) for i in range(80)]
    h = A.build_level_histogram(iv, AX, 1.0)
    p75 = A.percentile_of_sorted(A.vad_levels(iv, AX), 75)
    assert abs(A.clamp_split(A.otsu_split(h), -48.0, p75) - (-48.0 + NOISE_MARGIN)) < 0.001


def test_percentile_floor():                          # analyser_vad_test.go:165-188
    levels = [-60 + float(i) for i in range(100)]
    assert A.percentile_floor(levels, -200.0) == A.percentile_of_sorted(levels, 10.0)
    assert A.percentile_floor([-90, -89, -88, -87, -86], -50.0) == -50.0 + NOISE_MARGIN


def test_floored_fraction():                          # analyser_vad_test.go:190-333
    iv = [vad_interval(i, -15) for i in range(40)] + [vad_interval(40 + i, -130) for i in range(40)] + \
         [vad_interval(80 + i, -math.inf) for i in range(20)]
    got = A.floored_fraction(iv)
    assert abs(got - 0.6) < 0.001 and got >= VA_FRACTION
    iv = [vad_interval(i, -55) for i in range(70)] + [vad_interval(70 + i, -15) for i in range(30)]
    assert A.floored_fraction(iv) == 0
    iv = [vad_interval(i, -15) for i in range(95)] + [vad_interval(95 + i, -60) for i in range(5)]
    assert A.floored_fraction(iv) == 0
    assert A.floored_fraction([vad_interval(i, -130) for i in range(30)]) == 1.0
    assert abs(A.floored_fraction([vad_interval(0, math.nan), vad_interval(1, -15)]) - 0.5) < 0.001
    iv = [vad_interval(i, math.nan) for i in range(25)] + [vad_interval(25 + i, -120) for i in range(25)] + \
         [vad_interval(50 + i, -15) for i in range(50)]
    assert abs(A.floored_fraction(iv) - 0.5) < 0.001
    assert A.floored_fraction([vad_interval(i, math.nan) for i in range(20)]) == 1.0
    assert A.floored_fraction([]) == 0


def test_floored_fraction_boundary():                 # analyser_vad_test.go:478-516
    def build(floored, total):
        return [vad_interval(i, -130) for i in range(floored)] + [vad_interval(floored + i, -15) for i in range(total - floored)]
    got = A.floored_fraction(build(20, 100))
    assert abs(got - 0.20) < 0.001 and got >= VA_FRACTION
    got = A.floored_fraction(build(19, 100))
    assert abs(got - 0.19) < 0.001 and not got >= VA_FRACTION


def test_estimate_noise_floor_tied_score_order_independent():     # analyser_vad_test.go:359-395
    iv = [seed_interval(-80 + float(i), 0.01) for i in range(25)] + [seed_interval(-30 + float(i), 0.50) for i in range(25)]
    fa, ta, oka = A.estimate_noise_floor_and_threshold(iv)
    fb, tb, okb = A.estimate_noise_floor_and_threshold(iv[::-1])
    assert oka and okb and fa == fb and ta == tb


def test_estimate_noise_floor_truncation_picks_lowest_rms():      # analyser_vad_test.go:397-431
    iv = [seed_interval(-56 - float(i), 0.01) for i in range(25)] + [seed_interval(-30 + float(i), 0.50) for i in range(25)]
    floor, thr, ok = A.estimate_noise_floor_and_threshold(iv)
    assert ok
    count = max(50 // 5, 8)
    assert abs(floor - (-80.0 + count - 1)) < 0.001
    assert thr == floor + 1.0                                      # silenceThresholdHeadroomDB


def test_estimate_noise_floor_excludes_floored():                  # analyser_vad_test.go:433-461
    iv = [seed_interval(-130, 0.01) for _ in range(3)] + [seed_interval(-70 + float(i), 0.01) for i in range(40)] + \
         [seed_interval(-10 + float(i), 0.50) for i in range(10)]
    floor, _, ok = A.estimate_noise_floor_and_threshold(iv)
    assert ok and floor > LEVEL_FLOOR


def test_estimate_noise_floor_all_floored_not_ok():               # analyser_vad_test.go:463-476
    _, _, ok = A.estimate_noise_floor_and_threshold([seed_interval(-130, 0.01) for _ in range(15)])
    assert not ok
    _, _, ok = A.estimate_noise_floor_and_threshold([seed_interval(-50, 0.01) for _ in range(9)])   # < silenceThresholdMinIntervals
    assert not ok


@pytest.mark.parametrize("level,centroid,entropy,want", [(-20, 2000, 0.4, True), (-20, 8000, 0.4, False), (-20, 2000, 0.9, False),
                                                         (-40, 2000, 0.4, False)])
def test_is_speech_interval(level, centroid, entropy, want):      # analyser_vad_test.go:518-543
    assert A.is_speech_interval(A.interval(0, momentary=level, centroid=centroid, entropy=entropy), -30.0) is want


def _runs(iv):
    return A.build_speech_runs(iv, -30.0, 3.0, A.intervals_for_duration(2 * S, HOP))


def _seq(*parts):
    iv, idx = [], 0
    for n, mk in parts:
        for _ in range(n):
            iv.append(mk(idx)); idx += 1
    return iv


def test_build_speech_runs():                                      # analyser_vad_test.go:560-690
    tol = A.intervals_for_duration(2 * S, HOP)
    min_n = A.intervals_for_duration(10 * S, HOP)
    assert (tol, min_n) == (8, 40)
    assert len(_runs(_seq((50, vad_speech), (tol - 1, vad_quiet), (50, vad_speech)))) == 1
    assert len(_runs(_seq((50, vad_speech), (tol + 5, vad_quiet), (50, vad_speech)))) == 2
    assert len(_runs(_seq((50, vad_speech), (3, lambda i: vad_interval(i, -31)), (50, vad_speech)))) == 1
    assert len(_runs(_seq((50, vad_speech), (1, vad_loud_non_speech), (50, vad_speech)))) == 2
    assert len(_runs(_seq((50, vad_speech), (1, vad_quiet), (50, vad_speech)))) == 1
    assert len(_runs(_seq((min_n - 1, vad_speech), (6, vad_quiet)))) == 0
    # run bounds: start of the first speech interval .. last speech interval + hop (analyser_vad.go:494-504)
    runs = _runs(_seq((5, vad_quiet), (50, vad_speech), (20, vad_quiet)))
    assert len(runs) == 1 and runs[0].start_ns == 5 * HOP and runs[0].end_ns == 55 * HOP and runs[0].duration_ns == 50 * HOP


def test_gap_tolerance_intervals():                                # analyser_vad_test.go:692-731
    flags = []
    for n, v in ((5, True), (4, False), (5, True), (6, False), (5, True), (12, False), (5, True), (30, False), (5, True), (20, False)):
        flags += [v] * n
    floor, ceil = A.intervals_for_duration(2 * S, HOP), A.intervals_for_duration(10 * S, HOP)
    want = max(floor, min(ceil, int(round(A.percentile_of_sorted([4, 6, 12, 30], 75)))))
    assert A.gap_tolerance_intervals(flags) == want == 12
    assert A.gap_tolerance_intervals([True, True, True, False, False]) == floor
    assert A.gap_tolerance_intervals([False] * 10) == floor


def _bimodal(lo, hi):
    return A.build_level_histogram([vad_interval(i, lo) for i in range(40)] + [vad_interval(40 + i, hi) for i in range(40)], AX, 1.0)


def test_hysteresis_margin():                                      # analyser_vad_test.go:733-747
    near, far = A.hysteresis_margin(_bimodal(-40, -30), -30.0), A.hysteresis_margin(_bimodal(-50, -10), -30.0)
    assert near > 0 and far > near
    # fraction of the split-to-upper-mode distance (analyser_vad.go:394-400): upper mode centre -9.5, split -30 -> 0.25 * 20.5
    assert abs(far - 0.25 * 20.5) < 1e-12


def test_elect_speech_profile():                                   # analyser_vad_test.go:785-832
    iv = _seq((140, lambda i: vad_speech_rich_at(i, -16.0)), (20, lambda i: vad_interval(i, -75)), (200, lambda i: vad_speech_rich_at(i, -34.0)))
    runs = A.build_speech_runs(iv, -45, 3, A.intervals_for_duration(2 * S, HOP))
    assert len(runs) == 2
    best, cands = A.find_best_speech_region(runs, iv, -60.0)
    assert best is not None and cands
    assert best.start_ns == 0                                      # the wide-SNR run A, not the longer run B
    elected = [c for c in cands if c.region.start_ns == best.start_ns][0]
    assert elected.sample.rms_level != 0 and elected.sample.crest_factor != 0


def test_pick_low_cluster_region():                                # analyser_vad_test.go:834-876
    iv = _seq((10, lambda i: vad_interval(i, -60)), (20, vad_speech_rich), (50, lambda i: vad_interval(i, -60)))
    region = A.pick_low_cluster_region(iv, -30)
    assert region is not None and region.start_ns >= 30 * HOP
    assert region.duration_ns == 10 * S                            # refined to the 10 s golden window
    prof = A.extract_noise_profile(region, iv)
    assert prof is not None and prof.spectral[A.SP_NAMES.index("centroid")] != 0
    assert prof.warning == 0                                       # 10 s is inside [8 s, 18 s]
    assert A.pick_low_cluster_region([vad_speech(i) for i in range(20)], -30) is None


def test_extract_noise_profile_spectral_fields():                  # analyser_vad_test.go:886-948
    iv = [A.interval(0, rms=-60, peak=-50, mean=1.0, variance=2.0, centroid=1400, spread=300, skewness=0.5, kurtosis=2.0,
                     entropy=0.4, flatness=0.3, crest=6.0, flux=0.02, slope=-0.4, decrease=0.10, rolloff=6000),
          A.interval(HOP, rms=-58, peak=-48, mean=3.0, variance=4.0, centroid=1600, spread=500, skewness=1.5, kurtosis=4.0,
                     entropy=0.6, flatness=0.5, crest=10.0, flux=0.06, slope=-0.2, decrease=0.14, rolloff=8000)]
    prof = A.extract_noise_profile(A.Region(0, 2 * HOP, 2 * HOP), iv)
    assert prof is not None
    assert abs(prof.entropy - 0.5) < 0.001
    want = dict(mean=2.0, variance=3.0, centroid=1500, spread=400, skewness=1.0, kurtosis=3.0, entropy=0.5, flatness=0.4, crest=8.0,
                flux=0.04, slope=-0.3, decrease=0.12, rolloff=7000)
    for k, v in want.items():
        assert abs(prof.spectral[A.SP_NAMES.index(k)] - v) < 0.001, k
    assert prof.measured_noise_floor == -59.0 and prof.peak_level == -48.0 and prof.crest_factor == 11.0
    assert prof.warning == 1                                       # 0.5 s < idealDurationMin


def test_derive_gate_statistics():                                 # analyser_vad_test.go:950-1158
    split = -30.0
    iv = [vad_interval(i, -60 + float(i)) for i in range(20)] + [vad_interval(20 + i, -25 + float(i)) for i in range(21)]
    v, n, s = A.derive_gate_statistics(iv, split, AX, A.Region.of(20 * HOP, 41 * HOP))
    assert (v, n, s) == (-23.0, -42.0, 19.0)
    iv = [vad_interval(i, -20 + float(i)) for i in range(11)] + [vad_loud_non_speech(11 + i) for i in range(5)]
    v, _, _ = A.derive_gate_statistics(iv, split, AX, A.Region.of(0, 16 * HOP))
    assert v == -19.0
    iv = [vad_interval(i, -25) for i in range(10)] + [vad_interval(10 + i, -15) for i in range(11)]
    v, _, _ = A.derive_gate_statistics(iv, split, AX, A.Region.of(10 * HOP, 21 * HOP))
    assert v == -15.0
    iv = [vad_interval(i, -60 + float(i)) for i in range(20)]
    assert A.derive_gate_statistics(iv, split, AX, None) == (0.0, -42.0, 42.0)
    iv = [vad_interval(i, -20 + float(i)) for i in range(11)]
    v, n, _ = A.derive_gate_statistics(iv, split, AX, A.Region.of(0, 11 * HOP))
    assert n == 0 and v == -19.0
    iv = [vad_interval(0, -55), vad_interval(1, -12)]
    assert A.derive_gate_statistics(iv, split, AX, A.Region.of(HOP, 2 * HOP)) == (-12.0, -55.0, 43.0)
    iv = [vad_interval(i, -50 + float(i)) for i in range(11)]
    v, n, _ = A.derive_gate_statistics(iv, -45.0, AX, A.Region.of(0, 11 * HOP))
    assert (v, n) == (-45.0, -47.0)
    iv = [vad_interval(i, -130) for i in range(10)] + [vad_interval(10 + i, -60 + float(i)) for i in range(20)]
    assert A.derive_gate_statistics(iv, split, AX, None)[1] == -42.0


def test_detect_voice_activity():                                  # analyser_vad_test.go:1160-1222
    iv = _seq((60, lambda i: vad_interval(i, -55)), (80, vad_speech_rich))
    va, runs, cands = A.vad_detect(iv, -70.0)
    assert va.has_speech_profile and va.has_noise_profile and va.has_room_tone_sample
    assert va.floor_source == 3                                    # "vad_percentile"
    assert -120 < va.floor < -16
    assert va.voiced_low_percentile != 0 and va.noise_high_percentile != 0 and va.gate_separation_db > 0
    h = A.build_level_histogram(iv, AX, 1.0)
    levels = A.vad_levels(iv, AX)
    split = A.clamp_split(A.otsu_split(h), -70, A.percentile_of_sorted(levels, 75))
    assert va.split == split
    want = A.derive_gate_statistics(iv, split, AX, va.speech_profile.region)
    assert (va.voiced_low_percentile, va.noise_high_percentile, va.gate_separation_db) == want
    assert va.noise_profile.measured_noise_floor == va.floor == A.percentile_floor(levels, -70)
    assert not va.voice_activated and va.floored_fraction == 0
    assert len(runs) == 1 and runs[0].start_ns == 60 * HOP and len(cands) == 1 and va.speech_profile_index == 0


def test_detect_voice_activity_no_profile():                       # analyser_vad_test.go:1224-1242
    va, runs, _ = A.vad_detect([vad_interval(i, -55) for i in range(60)], -70.0)
    assert not va.has_speech_profile and va.voiced_low_percentile == 0 and not runs


@pytest.mark.parametrize("level,want", [(-40.0, False), (LEVEL_FLOOR, True), (LEVEL_FLOOR - 1, True), (math.inf, True), (-math.inf, True),
                                        (math.nan, True)])
def test_is_floored_level(level, want):                            # analyser_vad_test.go:1244-1265
    # a floored level never enters the histogram (buildLevelHistogram skips it)
    h = A.build_level_histogram([vad_interval(0, level)], AX, 1.0)
    assert (h.count == 0) is want


def test_full_entry_seeds_from_intervals():
    """jt_detect_voice_activity = buildInputMeasurements' seed (analyser.go:374-395) + detectVoiceActivity +
    assignInputMeasurementSuggestions (analyser.go:515-531)."""
    iv = _seq((60, lambda i: vad_interval(i, -55)), (80, vad_speech_rich))
    seed, thr, ok = A.estimate_noise_floor_and_threshold(iv)
    assert ok
    m = A.new_measurements(input_i=-20.0, input_lra=6.0, Dynamic_range=60.0, RMS_level=-22.0, Peak_level=-4.0, Noise_floor=-70.0)
    va, _, _ = A.detect_voice_activity(m, iv)
    ref, _, _ = A.vad_detect(iv, seed)
    assert va.floor_prescan == seed and va.room_tone_detect_level == thr
    assert (va.split, va.floor, va.margin, va.gap_tolerance) == (ref.split, ref.floor, ref.margin, ref.gap_tolerance)
    assert va.floor_astats == -70.0
    assert va.reduction_headroom == max(0.0, min(60.0, -22.0 - va.floor))
    # fully gated capture: no measurable room tone -> the -115 sentinel and the -70 detect level (analyser.go:376-391)
    gated = [vad_interval(i, -130) for i in range(30)]
    va, _, _ = A.detect_voice_activity(m, gated)
    assert va.floor_prescan == LEVEL_FLOOR and va.room_tone_detect_level == -70.0 and va.voice_activated
