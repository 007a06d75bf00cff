"""Speech-candidate measurement, scoring, election and golden-window refinement (csrc/jt_adapt.cu) against the reference's
unit-test tables: /root/reference/internal/processor/analyser_candidates_speech_test.go and the candidate sections of
analyser_test.go (cited per test).  Host-only."""
import math

import pytest

from jivetalking_b200 import adapt as A

HOP, S, MS = A.HOP_NS, A.NS_S, A.NS_MS
MIN_SNR, ADEQ_MIN = 20.0, 30 * S
MIN_VIABLE = 0.3


def grounded_candidate(rms, duration_ns):                  # analyser_candidates_speech_test.go:11-18
    c = A.SpeechCandidate()
    c.region = A.Region(0, 0, duration_ns)
    c.sample.rms_level = rms
    return c


def speech_run_intervals(start_ns, count, level):          # analyser_candidates_speech_test.go:117-128
    return [A.interval(start_ns + i * HOP, rms=level, momentary=level, peak=level + 12.0) for i in range(count)]


def make_test_intervals(rms_vals):                         # analyser_test.go makeTestIntervals
    return [A.interval(i * HOP, rms=r) for i, r in enumerate(rms_vals)]


def make_speech_test_intervals(count, rms):                # analyser_test.go:387-406
    return [A.interval(i * HOP, rms=rms, centroid=1500.0, entropy=0.5) for i in range(count)]


def scorable(start_ns, count, kurtosis, flatness, centroid, rms):    # analyser_test.go:587-601
    return [A.interval(start_ns + i * HOP, rms=rms, kurtosis=kurtosis, flatness=flatness, centroid=centroid, rolloff=6000.0, flux=0.003)
            for i in range(count)]


def region(start_s, end_s):
    return A.Region.of(int(start_s * S), int(end_s * S))


def test_grounded_snr_monotonicity():                      # analyser_candidates_speech_test.go:50-73
    floor, dur = -60.0, 45 * S
    sc = lambda rms: A.score_speech_candidate_grounded(grounded_candidate(rms, dur), floor, 0.0)
    assert sc(floor + 45.0) > sc(floor + 25.0)
    assert sc(floor + MIN_SNR - 10.0) < sc(floor + MIN_SNR + 5.0)
    # the four arms of groundedSNRScore (analyser_candidates_speech.go:402-413), weights 0.6 / 0.4 / 0.02
    assert sc(floor) == pytest.approx(0.4 + 0.02)
    assert sc(floor + 10.0) == pytest.approx(0.6 * 0.25 + 0.4 + 0.02)
    assert sc(floor + 30.0) == pytest.approx(0.6 * 0.75 + 0.4 + 0.02)
    assert sc(floor + 50.0) == pytest.approx(1.02)


def test_grounded_duration_saturation():                   # analyser_candidates_speech_test.go:75-96
    sc = lambda d: A.score_speech_candidate_grounded(grounded_candidate(-20.0, d), -60.0, 0.0)
    assert sc(ADEQ_MIN) == sc(ADEQ_MIN * 3)
    assert sc(ADEQ_MIN // 2) < sc(ADEQ_MIN)
    assert sc(ADEQ_MIN // 2) == pytest.approx(0.6 + 0.4 * 0.5 + 0.02)


def test_grounded_consistency_tie_break():                 # analyser_candidates_speech_test.go:98-112
    c = grounded_candidate(-20.0, 45 * S)
    assert A.score_speech_candidate_grounded(c, -60.0, 1.0) > A.score_speech_candidate_grounded(c, -60.0, 9.0)
    assert A.score_speech_candidate_grounded(c, -60.0, 25.0) == A.score_speech_candidate_grounded(c, -60.0, 90.0)


def test_find_best_voice_activated_case():                 # analyser_candidates_speech_test.go:132-158
    n = ADEQ_MIN // HOP + 4
    short = speech_run_intervals(0, n, -18.0)
    short_end = n * HOP
    long_start = short_end + 5 * S
    long_ = speech_run_intervals(long_start, n * 3, -38.0)
    long_end = long_start + 3 * n * HOP
    best, _ = A.find_best_speech_region([A.Region.of(0, short_end), A.Region.of(long_start, long_end)], short + long_, -60.0)
    assert best is not None and best.start_ns == 0


def test_find_best_always_elects():                        # analyser_candidates_speech_test.go:162-184
    run = speech_run_intervals(0, 12, -33.0)
    best, cands = A.find_best_speech_region([A.Region.of(0, 12 * HOP)], run, -35.0)
    assert best is not None and best.start_ns == 0 and len(cands) == 1
    assert cands[0].score < MIN_VIABLE


def test_find_best_all_below_snr_minimum_elects_highest():  # analyser_candidates_speech_test.go:193-219
    lo = speech_run_intervals(0, 74, -49.35)
    lo_end = 74 * HOP
    hi_start = lo_end + 5 * S
    hi = speech_run_intervals(hi_start, 81, -48.46)
    best, _ = A.find_best_speech_region([A.Region.of(0, lo_end), A.Region.of(hi_start, hi_start + 81 * HOP)], lo + hi, -60.0)
    assert best is not None and best.start_ns == hi_start


def test_level_variance():                                 # analyser_candidates_speech_test.go:221-239
    flat = [A.interval(i * HOP, rms=-20.0) for i in range(20)]
    spread = [A.interval(i * HOP, rms=-20.0 + (4.0 if i % 2 == 0 else -4.0)) for i in range(20)]
    assert A.level_variance(flat, A.AXIS_RMS) < 1e-9
    assert A.level_variance(spread, A.AXIS_RMS) == pytest.approx(16.0)
    assert A.level_variance([], A.AXIS_RMS) == 0


@pytest.mark.parametrize("start,end,count,first,last", [(0, 20 * S, 80, 0, 19750 * MS), (5 * S, 15 * S, 40, 5 * S, 14750 * MS),
                                                        (25 * S, 30 * S, 0, None, None), (0, 2 * S, 8, 0, 1750 * MS)])
def test_get_intervals_in_range(start, end, count, first, last):    # analyser_test.go:264-333
    iv = make_test_intervals([0.0] * 80)
    lo, n = A.get_intervals_in_range(iv, start, end)
    assert n == count
    if count:
        assert round(iv[lo].timestamp_s * 1e9) == first and round(iv[lo + n - 1].timestamp_s * 1e9) == last


@pytest.mark.parametrize("vals,want", [([-70, -70, -70, -70], -70.0), ([-60, -70, -80, -70], -70.0), ([-65.5], -65.5), ([], 0.0)])
def test_score_interval_window(vals, want):                # analyser_test.go:335-380
    assert abs(A.score_interval_window(make_test_intervals(vals)) - want) < 0.001


def test_measure_speech_candidate():                       # analyser_test.go:408-468
    iv = [A.interval(i * HOP, rms=-20.0, peak=-8.0, centroid=1500.0, flatness=0.3, kurtosis=5.0, entropy=0.5) for i in range(40)]
    iv[20].peak_level = -5.0
    m = A.measure_speech_candidate(region(0, 10), iv)
    assert m.sample.rms_level == -20.0 and m.sample.peak_level == -5.0 and m.sample.crest_factor == 15.0
    assert m.sample.spectral[A.SP_NAMES.index("centroid")] == 1500.0
    assert m.voicing_density == 1.0                        # kurtosis 5.0 > 4.5 everywhere
    assert A.measure_speech_candidate(region(100, 110), make_speech_test_intervals(40, -20.0)) is None


def test_find_best_speech_region():                        # analyser_test.go:470-520
    iv = make_speech_test_intervals(400, -18.0)
    best, _ = A.find_best_speech_region([region(0, 35), region(40, 90), region(95, 100)], iv)
    assert best is not None and best.start_ns == 0
    best, cands = A.find_best_speech_region([], make_speech_test_intervals(200, -18.0))
    assert best is None and not cands
    _, cands = A.find_best_speech_region([region(0, 35), region(40, 80)], iv)
    assert len(cands) == 2


def test_find_best_all_below_min_acceptable_falls_back():  # analyser_test.go:522-575
    def short_run(start_s, rms):
        return [A.interval(int(start_s * S) + i * HOP, rms=rms, momentary=rms, peak=rms + 10.0) for i in range(40)]
    iv = short_run(0, -33.0) + short_run(15, -27.0)
    best, cands = A.find_best_speech_region([region(0, 10), region(15, 25)], iv, -35.0)
    assert best is not None and best.start_ns == 15 * S and len(cands) == 2
    assert all(c.score < MIN_VIABLE for c in cands) and cands[1].score > cands[0].score


def _pause_heavy():
    return [A.interval(i * HOP, rms=-35.0, kurtosis=15.0 if i % 2 == 0 else 1.0, flatness=0.8, centroid=7000.0, rolloff=12000.0, flux=0.05)
            for i in range(40)]


@pytest.mark.parametrize("iv,lo,hi", [
    (scorable(0, 40, 6.0, 0.1, 2000.0, -15.0), 0.80, 1.0),
    (_pause_heavy(), 0.0, 0.40),
    ([], 0.0, 0.0),
    (scorable(0, 40, 2.0, 0.8, 7000.0, -32.0), 0.25, 0.50),
    (scorable(0, 40, 6.0, 0.1, 4400.0, -15.0), 0.75, 0.95),
    (scorable(0, 40, 6.0, 0.1, 2000.0, -28.0), 0.75, 0.90)])
def test_score_speech_interval_window(iv, lo, hi):         # analyser_test.go:603-723
    s = A.score_speech_interval_window(iv)
    assert lo <= s <= hi and 0.0 <= s <= 1.0


def test_score_speech_interval_window_hand_computed():
    """analyser_candidates_shared.go:194-292 term by term for one window (kurtosis 6, flatness 0.1, centroid 2000,
    RMS -15, rolloff 6000, flux 0.003)."""
    want = (6.0 / 7.5) * 0.15 + 0.9 * 0.10 + (1.0 - (abs(2000.0 - 3100.0) / 2900.0) * 0.5) * 0.10 + 1.0 * 0.10 + \
           ((-15.0 + 30.0) / 18.0) * 0.10 + 1.0 * 0.15 + 1.0 * 0.15 + 1.0 * 0.15
    assert A.score_speech_interval_window(scorable(0, 40, 6.0, 0.1, 2000.0, -15.0)) == pytest.approx(want, abs=1e-15)


@pytest.mark.parametrize("cand,iv,start,dur", [
    (region(10, 50), scorable(10 * S, 160, 6.0, 0.1, 2000.0, -15.0), 10 * S, 40 * S),
    (region(0, 120), scorable(0, 480, 6.0, 0.1, 2000.0, -15.0), 0, 60 * S),
    (region(0, 120), scorable(0, 240, 3.0, 0.5, 2000.0, -25.0) + scorable(60 * S, 240, 8.0, 0.08, 2000.0, -12.0), 60 * S, 60 * S),
    (region(0, 90), scorable(0, 100, 6.0, 0.1, 2000.0, -15.0), 0, 90 * S),
    (region(200, 320), scorable(0, 480, 6.0, 0.1, 2000.0, -15.0), 200 * S, 120 * S)])
def test_refine_to_golden_speech_subregion(cand, iv, start, dur):   # analyser_test.go:725-847
    r = A.refine_to_golden_speech_subregion(cand, iv)
    assert (r.start_ns, r.duration_ns, r.end_ns) == (start, dur, start + dur)


def test_find_best_with_refinement():                      # analyser_test.go:849-964
    iv = scorable(0, 240, 4.0, 0.3, 2000.0, -20.0) + scorable(60 * S, 240, 7.0, 0.1, 2000.0, -14.0)
    best, cands = A.find_best_speech_region([region(0, 120)], iv)
    assert best is not None and cands
    refined = [c for c in cands if c.was_refined]
    assert refined and refined[0].original_start_ns == 0 and refined[0].original_duration_ns == 120 * S
    assert refined[0].region.duration_ns <= 60 * S
    best, cands = A.find_best_speech_region([region(0, 45)], scorable(0, 180, 6.0, 0.1, 2000.0, -15.0))
    assert best is not None and not any(c.was_refined for c in cands) and best.duration_ns == 45 * S
    iv = scorable(0, 120, 2.0, 0.6, 3500.0, -28.0) + scorable(30 * S, 240, 8.0, 0.05, 2000.0, -12.0) + scorable(90 * S, 120, 2.0, 0.6, 3500.0, -28.0)
    best, _ = A.find_best_speech_region([region(0, 120)], iv)
    assert best is not None and 30 * S <= best.start_ns <= 60 * S and best.duration_ns == 60 * S
    assert best.start_ns == 30 * S


def test_find_best_snr_margin():                           # analyser_test.go:966-1026
    iv = scorable(0, 140, 6.0, 0.1, 1500.0, -20.0)
    score = lambda floor: A.find_best_speech_region([region(0, 35)], iv, floor)[1][0].score
    assert score(-55.0) > score(-30.0)
    assert score(-math.inf) >= score(-40.0)
    assert score(-math.inf) == pytest.approx(1.02)
