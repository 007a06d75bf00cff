"""Deterministic synthetic PCM for tests and bench.py.

`reference_test_audio` is a vectorised port of the reference's own generator
(internal/processor/testutil_test.go:27-137: sine + LCG white noise + optional noise-only
gap, LCG state*1664525+1013904223 seeded 12345, int16 truncation toward zero).
`speech_like` is the C2/C3/C4/C5 recipe of SURVEY.md 8d: a harmonic carrier with a syllabic
envelope, stepped phrase levels and noise-only pauses, so loudness range is > 0 (FFmpeg's
loudnorm only stays in linear mode when measured_LRA != 0).
"""
import numpy as np

_A, _C = np.uint32(1664525), np.uint32(1013904223)


def lcg_uniform(n, seed=12345):
    """n draws of the reference's nextRandom(): uniform in [-1, 1]."""
    if n <= 0:
        return np.zeros(0)
    out = np.empty(n, dtype=np.float64)
    state = np.uint32(seed)
    block = 1 << 22
    with np.errstate(over="ignore"):
        k = min(block, n)
        a = np.cumprod(np.full(k, _A, dtype=np.uint32), dtype=np.uint32)          # A^(j+1)
        geo = np.concatenate(([np.uint32(1)], a[:-1]))                              # A^j
        c = (np.cumsum(geo, dtype=np.uint32) * _C).astype(np.uint32)               # C * sum_{i<=j} A^i
        pos = 0
        while pos < n:
            m = min(k, n - pos)
            s = (a[:m] * state + c[:m]).astype(np.uint32)
            out[pos:pos + m] = s.astype(np.float64) / float(0xFFFFFFFF) * 2.0 - 1.0
            state = s[m - 1]
            pos += m
    return out


def reference_test_audio(duration_s=5.0, rate=44100, tone_hz=440.0, tone_db=-23.0, noise_db=-60.0,
                         gap_start_s=0.0, gap_dur_s=0.0, seed=12345):
    """int16 mono, exactly the samples generateTestAudio() writes into its WAV."""
    n = int(duration_s * rate)
    tone_amp = 10.0 ** (tone_db / 20.0) if (tone_hz > 0 and tone_db < 0) else 0.0
    noise_amp = 10.0 ** (noise_db / 20.0) if noise_db < 0 else 0.0
    i = np.arange(n)
    s = np.zeros(n)
    if tone_amp > 0:
        s += tone_amp * np.sin(2.0 * np.pi * tone_hz * (i / float(rate)))
    rnd = lcg_uniform(n, seed) if noise_amp > 0 else None
    if noise_amp > 0:
        s += noise_amp * rnd
    if gap_dur_s > 0:
        g0, g1 = int(gap_start_s * rate), int((gap_start_s + gap_dur_s) * rate)
        g1 = min(g1, n)
        s[g0:g1] = noise_amp * rnd[g0:g1] if noise_amp > 0 else 0.0
    s = np.clip(s, -1.0, 1.0)
    return np.trunc(s * 32767.0).astype(np.int16)


def speech_like(duration_s, rate=48000, seed=12345, dtype=np.float32):
    """Mono speech-like programme in [-1, 1]: 180 Hz harmonic carrier (spectral centroid in the
    voice band), 3-5 Hz syllabic envelope, phrase levels stepping among {-30,-24,-18,-14} dBFS
    every 5-20 s, 1-3 s noise-only pauses about every 15 s and one >= 10 s pause; -58 dBFS white
    noise floor throughout (reference LCG)."""
    n = int(duration_s * rate)
    rng = np.random.default_rng(seed)
    t = np.arange(n, dtype=np.float64) / rate
    f0 = 180.0
    car = np.zeros(n)
    for h, a in ((1, 1.0), (2, 0.6), (3, 0.45), (4, 0.3), (5, 0.22), (7, 0.15), (10, 0.1), (14, 0.06), (20, 0.04)):
        car += a * np.sin(2 * np.pi * f0 * h * t + 0.37 * h)
    car /= np.max(np.abs(car)) + 1e-12
    syl = 0.55 + 0.45 * np.sin(2 * np.pi * (3.0 + 2.0 * rng.random()) * t) * np.sin(2 * np.pi * 0.31 * t + 1.0)
    # piecewise-constant phrase level and speech/pause mask
    level = np.zeros(n)
    mask = np.ones(n)
    pos, long_pause_done = 0.0, False
    next_pause = 12.0 + 6.0 * rng.random()
    while pos < duration_s:
        seg = 5.0 + 15.0 * rng.random()
        a, b = int(pos * rate), min(int((pos + seg) * rate), n)
        level[a:b] = 10.0 ** (rng.choice([-30.0, -24.0, -18.0, -14.0]) / 20.0)
        pos += seg
    pos = next_pause
    k = int(0.02 * rate)            # 20 ms raised-cosine edges so pauses do not click
    ramp = 0.5 * (1.0 + np.cos(np.pi * (np.arange(k) + 0.5) / max(k, 1)))      # 1 -> 0
    while pos < duration_s:
        if not long_pause_done and pos > min(30.0, duration_s * 0.4):
            dur, long_pause_done = 10.5, True
        else:
            dur = 1.0 + 2.0 * rng.random()
        a, b = int(pos * rate), min(int((pos + dur) * rate), n)
        mask[a:b] = 0.0
        if a - k >= 0:
            mask[a - k:a] = np.minimum(mask[a - k:a], ramp)
        if b + k <= n:
            mask[b:b + k] = np.minimum(mask[b:b + k], ramp[::-1])
        pos += dur + 11.0 + 8.0 * rng.random()
    x = car * syl * level * mask + 10.0 ** (-58.0 / 20.0) * lcg_uniform(n, seed)
    return np.clip(x, -1.0, 1.0).astype(dtype)


def stereo_from_mono(x, delay=7, gain=0.9):
    """C4: R = gain * L delayed by `delay` samples; returns interleaved stereo."""
    r = np.zeros_like(x)
    r[delay:] = x[:-delay] * gain
    out = np.empty(2 * len(x), dtype=x.dtype)
    out[0::2] = x
    out[1::2] = r
    return out


def podcast_like(duration_s, rate=48000, seed=2024, dtype=np.float32, sibilance_db=None):
    """speech_like()'s carrier with a conversational duty cycle: 12-25 s phrases separated by 4-9 s room-tone pauses
    (one of 14 s), so more than a fifth of the 250 ms intervals is room tone -- what the reference's noise-floor seed
    (top 20 % most room-tone-like intervals, analyser_noise_seed.go:150-223) needs to find the floor and its
    voice-activity detector needs to elect a speech profile and a room-tone region.  sibilance_db adds a 6-9 kHz
    noise band to the phrases at that level relative to the carrier (drives the de-esser branch)."""
    n = int(duration_s * rate)
    rng = np.random.default_rng(seed)
    t = np.arange(n, dtype=np.float64) / rate
    car = np.zeros(n)
    for h, a in ((1, 1.0), (2, 0.6), (3, 0.45), (4, 0.3), (5, 0.22), (7, 0.15), (10, 0.1), (14, 0.06), (20, 0.04)):
        car += a * np.sin(2 * np.pi * 180.0 * h * t + 0.37 * h)
    car /= np.max(np.abs(car)) + 1e-12
    if sibilance_db is not None:
        w = lcg_uniform(n, seed + 77)
        spec = np.fft.rfft(w)
        f = np.fft.rfftfreq(n, 1.0 / rate)
        spec[(f < 6000.0) | (f > 9000.0)] = 0.0
        sib = np.fft.irfft(spec, n)
        sib *= 10.0 ** (sibilance_db / 20.0) * np.sqrt(np.mean(car ** 2)) / (np.sqrt(np.mean(sib ** 2)) + 1e-30)
        car = car + sib
    syl = 0.6 + 0.4 * np.sin(2 * np.pi * (3.0 + 2.0 * rng.random()) * t) * np.sin(2 * np.pi * 0.31 * t + 1.0)
    gain = np.zeros(n)
    k = int(0.02 * rate)
    ramp = 0.5 * (1.0 - np.cos(np.pi * (np.arange(k) + 0.5) / max(k, 1)))      # 0 -> 1
    pos, first_pause = 3.0 + 2.0 * rng.random(), True
    while pos < duration_s:
        seg = 12.0 + 13.0 * rng.random()
        a, b = int(pos * rate), min(int((pos + seg) * rate), n)
        lvl = 10.0 ** (rng.choice([-26.0, -22.0, -18.0, -15.0]) / 20.0)
        g = np.full(b - a, lvl)
        m = min(k, (b - a) // 2)
        if m > 0:
            g[:m] *= ramp[:m]
            g[-m:] *= ramp[:m][::-1]
        gain[a:b] = g
        pos += seg + (14.0 if first_pause else 4.0 + 5.0 * rng.random())
        first_pause = False
    x = car * syl * gain + 10.0 ** (-58.0 / 20.0) * lcg_uniform(n, seed)
    return np.clip(x, -1.0, 1.0).astype(dtype)
