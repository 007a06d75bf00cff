"""Pins the oracle's libswresample restatement (oracle/orc_swr.c) against outputs of the REAL
FFmpeg libswresample 6.1.100 (tests/golden/swr_golden.npz, made by scripts/make_golden_swr.py),
and live against the library when this image has it."""
import os
import numpy as np
import pytest
import jt_oracle as O

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "swr_golden.npz"))
RATES = [(48000, 192000), (48000, 44100), (44100, 192000), (96000, 44100)]


@pytest.mark.parametrize("name", ["tone", "noise"])
@pytest.mark.parametrize("rates", RATES)
@pytest.mark.parametrize("flush", [0, 1])
def test_resample_f64_matches_real_swr(name, rates, flush):
    x = G[f"in_{name}"]
    ref = G[f"dbl_{name}_{rates[0]}_{rates[1]}_{flush}"]
    y = O.swr_resample(x, rates[0], rates[1], flush=bool(flush))
    assert len(y) == len(ref)                       # bit-exact output COUNT (index arithmetic)
    assert np.max(np.abs(y - ref)) < 2e-15          # taps summed in a different order than the SIMD asm


INEXACT = [(22050, 192000), (11025, 192000), (47999, 44100)]


@pytest.mark.parametrize("rates", INEXACT)
@pytest.mark.parametrize("flush", [0, 1])
def test_inexact_ratio_linear_interpolated_path_matches_real_swr(rates, flush):
    """reduced phase count > 1024 (a 22.05 kHz source into ebur128's 192 kHz true-peak oversampler): 1024 phases, fractional
    index advance, linear interpolation between neighbouring phases (resample_template.c resample_linear)"""
    x = G["in_noise"][:2000]
    ref = G[f"dbl_noise_{rates[0]}_{rates[1]}_{flush}"]
    y = O.swr_resample(x, rates[0], rates[1], flush=bool(flush))
    assert len(y) == len(ref)
    assert np.max(np.abs(y - ref)) < 2e-15
    if rates == (22050, 192000) and flush:
        yf = O.swr_resample(x.astype(np.float32), 22050, 192000, flush=True)
        assert len(yf) == len(G["flt_22050_192000"]) and np.max(np.abs(yf - G["flt_22050_192000"])) < 5e-7


def test_resample_f32_internal_path():
    x = G["in_noise"].astype(np.float32)
    ref = G["flt_44100_192000"]
    y = O.swr_resample(x, 44100, 192000, flush=True)
    assert len(y) == len(ref) and np.max(np.abs(y - ref)) < 5e-7
    # s16 input, dbl output, rate change -> swr picks FLTP internally
    s16 = G["in_s16"]
    xf = np.zeros(len(s16), dtype=np.float32)
    O.lib().orc_conv_s16_to_f32(O._ptr(s16), len(s16), O._ptr(xf))
    y = O.swr_resample(xf, 44100, 192000, flush=True).astype(np.float64)
    ref = G["s16_to_dbl_44100_192000"]
    assert len(y) == len(ref) and np.max(np.abs(y - ref)) < 5e-7


def test_format_conversions_and_downmix():
    x = G["in_noise"] * 4.0
    out = np.zeros(len(x), dtype=np.int16)
    O.lib().orc_conv_f64_to_s16(O._ptr(x), len(x), O._ptr(out))
    assert np.array_equal(out, G["dbl_to_s16"])
    xf = x.astype(np.float32)
    O.lib().orc_conv_f32_to_s16(O._ptr(xf), len(xf), O._ptr(out))
    assert np.array_equal(out, G["flt_to_s16"])
    st = G["in_stereo_f32"]
    m = np.zeros(len(st) // 2, dtype=np.float32)
    O.lib().orc_downmix_stereo_f32(O._ptr(st), len(m), O._ptr(m))
    assert np.max(np.abs(m - G["stereo_f32_to_mono"])) < 1.2e-7
    st16 = G["in_stereo_s16"]
    m16 = np.zeros(len(st16) // 2, dtype=np.int16)
    O.lib().orc_downmix_stereo_s16(O._ptr(st16), len(m16), O._ptr(m16))
    assert np.max(np.abs(m16.astype(int) - G["stereo_s16_to_mono"].astype(int))) <= 1


def test_out_counts_edge_cases():
    for n in (0, 1, 31, 32, 33, 34, 100):
        for (a, b) in RATES:
            c0 = O.swr_out_count(n, a, b, flush=False)
            c1 = O.swr_out_count(n, a, b, flush=True)
            assert 0 <= c0 <= c1


def test_live_against_library_when_present():
    import ref_swr
    if not ref_swr.available():
        pytest.skip("libswresample not in this image")
    rng = np.random.default_rng(3)
    for n in (150, 4097, 30011):      # streams shorter than the filter (n <= filter_length) take a swr corner path not restated
        x = rng.standard_normal(n) * 0.1
        for (a, b) in RATES + [(44100, 48000)] + INEXACT + [(44100, 47999), (37800, 192000)]:
            ref = ref_swr.convert(x, "dbl", a, "dbl", b, frame=1000, flush=True)
            y = O.swr_resample(x, a, b, flush=True)
            assert len(y) == len(ref) and (len(y) == 0 or np.max(np.abs(y - ref)) < 2e-15)


def test_s32_conversions_and_downmix_match_real_swr():
    """32-bit integer samples (what 24-bit FLAC / WAV decode to; analyser_metrics.go:285,329 handles them): the oracle's
    numpy restatements of audioconvert.c's s32 rows and of the flt-internal normalised rematrix, against the real library."""
    import oracle_graph as OG
    s32 = G["in_s32"]
    assert np.array_equal(OG.to_f32(s32), G["s32_to_flt"])
    assert np.array_equal(OG.to_f64(s32), G["s32_to_dbl"])
    assert np.array_equal((s32 >> 16).astype(np.int16), G["s32_to_s16"])
    x = G["in_noise"] * 4.0
    assert np.array_equal(OG.f32_to_s32(x.astype(np.float32)), G["flt_to_s32"])
    assert np.array_equal(OG.f64_to_s32(x), G["dbl_to_s32"])
    assert np.array_equal(OG.downmix(G["in_stereo_s32"], 2), G["stereo_s32_to_mono"])
    # a 32-bit integer input with a rate change runs swr's FLTP-internal resampler
    y = O.swr_resample(OG.to_f32(s32), 48000, 192000, flush=True).astype(np.float64)
    ref = G["s32_to_dbl_48000_192000"]
    assert len(y) == len(ref) and np.max(np.abs(y - ref)) < 5e-7


@pytest.mark.parametrize("rates", [(192000, 44100), (192000, 48000)])
def test_downsample_from_192k_matches_real_swr(rates):
    """the aresample barrier behind a dynamic-mode loudnorm (normalise.go:1294-1304)"""
    ref = G[f"dbl_noise_{rates[0]}_{rates[1]}_1"]
    y = O.swr_resample(G["in_noise"], rates[0], rates[1], flush=True)
    assert len(y) == len(ref) and np.max(np.abs(y - ref)) < 2e-15


@pytest.mark.parametrize("rate", [22050, 48000, 44100])
def test_ebur128_true_peak_is_the_real_resamplers_peak(rate):
    """ebur128 peak=true feeds swr (-> 192 kHz, dbl) 100 ms at a time and keeps the running maximum of what comes out: the oracle's
    per-tick cumulative true peak against the REAL libswresample driven the same way -- including a 22.05 kHz source, whose ratio
    to 192 kHz takes swr's linear-interpolated path."""
    import ref_swr
    if not ref_swr.available():
        pytest.skip("libswresample not in this image")
    rng = np.random.default_rng(rate)
    tick = rate // 10
    x = (rng.standard_normal(tick * 23 + 517) * 0.2) * np.hanning(tick * 23 + 517)
    r = O.ebur128(x, rate, true_peak=True)
    counts = []
    y = ref_swr.convert(x[:r["n_ticks"] * tick], "dbl", rate, "dbl", 192000, frame=tick, flush=False, per_call_counts=counts)
    assert len(counts) == r["n_ticks"]
    ends = np.cumsum(counts)
    want = np.array([np.max(np.abs(y[:e])) if e else 0.0 for e in ends])
    assert np.max(np.abs(r["true_peak_cum"] - want)) < 1e-14
