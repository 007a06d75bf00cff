"""GPU parity, one filter at a time through the C ABI (jt_run_graph with a one-filter spec)
against the CPU oracle, with the reference's production parameters (filters.go, normalise.go)."""
import numpy as np
import pytest
import jt_oracle as O
import oracle_graph as OG
from jivetalking_b200 import gpudsp, synth

pytestmark = pytest.mark.gpu


def rms(a):
    return float(np.sqrt(np.mean(np.square(a.astype(np.float64))))) if len(a) else 0.0


def run(ctx, spec, x, rate):
    got = ctx.run_graph(spec, x, rate, want_meta=False)
    exp = OG.run_spec(spec, x, rate)
    assert got["rate"] == exp["rate"] and len(got["pcm"]) == len(exp["pcm"])
    assert got["pcm"].dtype == exp["pcm"].dtype
    return got["pcm"], exp["pcm"]


@pytest.fixture(scope="module")
def speech48():
    return synth.speech_like(25.0, 48000, seed=12345)


@pytest.fixture(scope="module")
def speech44_f64():
    return synth.speech_like(25.0, 44100, seed=777).astype(np.float64) * 3.0


@pytest.mark.parametrize("spec", [
    "highpass=f=80:poles=2:width_type=q:width=0.707:normalize=1:a=tdii",
    "lowpass=f=20500:poles=2:width_type=q:width=0.707:normalize=1:a=tdii",
    "highpass=f=1000.000000:p=2,lowpass=f=3000.000000:p=2",
])
def test_biquads_f32(ctx, speech48, spec):
    g, e = run(ctx, spec, speech48, 48000)
    # identical unfused float arithmetic.  A float32 biquad with poles this close to 1 carries its own
    # round-off noise (~1/(1-r)^2 ulps), and two runs that start from different states never re-merge
    # bit-for-bit: the lanes agree with the sequential run to within that noise, not below it.
    assert np.max(np.abs(g - e)) <= 5e-5 * max(1.0, np.max(np.abs(e)))
    assert rms(g - e) < 1e-5


def test_biquads_s16_and_f64(ctx, speech48):
    s16 = np.clip(np.round(speech48 * 32768 * 2), -32768, 32767).astype(np.int16)
    g, e = run(ctx, "highpass=f=80:poles=2:width_type=q:width=0.707:normalize=1:a=tdii", s16, 48000)
    assert np.max(np.abs(g.astype(int) - e.astype(int))) <= 1
    g, e = run(ctx, "lowpass=f=3000:p=2", speech48.astype(np.float64), 48000)
    assert np.max(np.abs(g - e)) < 1e-12


def test_anlmdn(ctx, speech48):
    g, e = run(ctx, "anlmdn=s=0.00001:p=0.0060:r=0.0020:m=3", speech48, 48000)
    assert np.max(np.abs(g - e)) < 2e-6 and rms(g - e) < 1e-7
    # the denoiser must actually act in the noise-only pauses (weights inside the cut-off)
    x = (synth.lcg_uniform(48000 * 3, 5) * 10 ** (-75 / 20)).astype(np.float32)
    g, e = run(ctx, "anlmdn=s=0.00001:p=0.0060:r=0.0020:m=3", x, 48000)
    assert rms(e[48000:]) < 0.9 * rms(x[48000:])
    assert np.max(np.abs(g - e)) < 1e-8


def test_anlmdn_44k_and_ragged(ctx):
    x = synth.speech_like(3.3, 44100, seed=9)[:-17]
    g, e = run(ctx, "anlmdn=s=0.00001:p=0.0060:r=0.0058:m=11", x, 44100)       # the pre-retune golden-test parameters
    assert np.max(np.abs(g - e)) < 2e-6


def _nlm_mixed(rate, seed):
    """bursts that decay into room tone of -58 .. -90 dBFS, digital silence, a tone fading out: hops on both sides of the
    screening decision (k_nlm_screen) and groups that mix them"""
    rng = np.random.default_rng(seed)
    parts = []
    for k, floor_db in enumerate((-58.0, -70.0, -90.0, None, -66.0)):
        n = int(rate * (0.35 + 0.11 * k))
        t = np.arange(n) / rate
        burst = synth.speech_like(n / rate + 0.1, rate, seed=seed + k)[:n] * np.exp(-t * (25.0 + 10 * k))
        tone = 0.2 * np.sin(2 * np.pi * (140 + 60 * k) * t) * np.exp(-t * 40.0)
        noise = 0.0 if floor_db is None else rng.standard_normal(n) * 10 ** (floor_db / 20)
        parts.append(burst + tone + noise)
        if floor_db is None:
            parts.append(np.zeros(int(rate * 0.05)))
    return np.concatenate(parts).astype(np.float32)


@pytest.mark.parametrize("rate,spec", [(48000, "anlmdn=s=0.00001:p=0.0060:r=0.0020:m=3"), (96000, "anlmdn=s=0.00001:p=0.0060:r=0.0020:m=3"),
                                       (44100, "anlmdn=s=0.00001:p=0.0060:r=0.0058:m=11"), (48000, "anlmdn=s=0.0001:p=0.0004:r=0.0010:m=5"),
                                       (192000, "anlmdn=s=0.00001:p=0.0030:r=0.0020:m=3")])
def test_anlmdn_screen_is_exact(ctx, monkeypatch, rate, spec):
    """the screening pass only skips hops whose output is the delayed input: bit-identical to walking every hop"""
    x = _nlm_mixed(rate, 31)[:-13]
    monkeypatch.setenv("JT_ANLMDN_NO_SCREEN", "1")
    full = ctx.run_graph(spec, x, rate, want_meta=False)["pcm"]
    monkeypatch.delenv("JT_ANLMDN_NO_SCREEN")
    got = ctx.run_graph(spec, x, rate, want_meta=False)["pcm"]
    assert np.array_equal(got.view(np.uint32), full.view(np.uint32))
    exp = OG.run_spec(spec, x, rate)["pcm"]
    assert np.max(np.abs(got - exp)) < 2e-6
    # both kinds of hop are present: samples the denoiser changed and samples it passed through
    K = int(round(float(spec.split("p=")[1].split(":")[0]) * rate)); S = int(round(float(spec.split("r=")[1].split(":")[0]) * rate))
    d = K + S
    same = exp[d:] == x[:len(x) - d]
    assert 0.02 < np.mean(same) < 0.98


@pytest.mark.parametrize("spec", ["afftdn=nr=12:nt=w:tn=1", "afftdn=nr=12:nt=w:tn=0:nf=-58",
                                  "afftdn=nr=12:nt=custom:bn=2.5|1.0|0.5|0.0|-0.5|-1.0|-1.5|-2.0|-2.0|-1.0|0.0|1.0|2.0|3.0|4.0:tn=0:nf=-62"])
def test_afftdn(ctx, speech48, spec):
    g, e = run(ctx, spec, speech48, 48000)
    assert rms(g - e) < 2e-6 and np.max(np.abs(g - e)) < 5e-5


def test_afftdn_44k(ctx):
    x = synth.speech_like(6.0, 44100, seed=4)
    g, e = run(ctx, "afftdn=nr=12:nt=w:tn=1", x, 44100)
    assert rms(g - e) < 2e-6


def test_gate_compressor_deesser(ctx, speech48):
    x = speech48.astype(np.float64)
    g, e = run(ctx, "agate=threshold=0.019953:ratio=2.0:attack=5.00:release=200:range=0.1995:knee=3.0:detection=rms:makeup=1.0", x, 48000)
    assert np.max(np.abs(g - e)) < 1e-12
    g, e = run(ctx, "acompressor=threshold=0.031623:ratio=3.0:attack=10:release=200:makeup=1.00:knee=4.0:detection=rms:mix=1.00", x, 48000)
    assert np.max(np.abs(g - e)) < 1e-12
    g, e = run(ctx, "deesser=i=0.45:m=0.50:f=0.80", x, 48000)
    assert np.max(np.abs(g - e)) < 1e-10


def test_limiters(ctx, speech44_f64):
    x = speech44_f64
    for spec in ("alimiter=limit=0.319890:attack=5:release=100:level_in=1:level_out=1:level=0:latency=1:asc=1:asc_level=0.8",
                 "alimiter=limit=0.803526:attack=1:release=50:level_in=1:level_out=1:level=0:latency=1:asc=1:asc_level=0.8",
                 "volume=4.2dB,alimiter=limit=0.063096:attack=5:release=100:level_in=1:level_out=1:level=0:latency=1:asc=1:asc_level=0.8"):
        g, e = run(ctx, spec, x, 44100)
        assert np.max(np.abs(e)) <= float(spec.split("limit=")[1].split(":")[0]) + 1e-12
        assert np.max(np.abs(g - e)) < 1e-9, spec


def test_adeclick(ctx, speech44_f64):
    x = O.alimiter(speech44_f64, 44100, 0.32, 5, 100, level=False, asc=True, asc_level=0.8)
    x[30000] += 0.4          # a real click
    g, e = run(ctx, "adeclick=t=1.7:w=55:o=50:m=s", x, 44100)
    assert np.array_equal(g, e)                      # same operations in the same order: bit-exact
    assert abs(e[30000] - x[30000]) > 0.2


def test_loudnorm_linear_and_dynamic_measure(ctx, speech44_f64):
    s16 = np.clip(np.round(speech44_f64 * 0.2 * 32768), -32768, 32767).astype(np.int16)
    spec = "loudnorm=I=-16.0:TP=-1.0:LRA=20.0:dual_mono=true:print_format=json"
    got = ctx.run_graph(spec, s16, 44100, want_pcm=False, want_meta=False)
    exp = OG.run_spec(spec, s16, 44100, want_pcm=False)
    ln = got["loudnorm"]
    assert ln.valid and ln.normalization_type == 1
    for k in ("input_i", "input_tp", "input_lra", "input_thresh"):
        assert abs(getattr(ln, k) - exp["loudnorm"][k]) < 5e-3, k
    spec = ("loudnorm=I=-16.00:TP=-0.70:LRA=20.0:measured_I=%.2f:measured_TP=%.2f:measured_LRA=%.2f:measured_thresh=%.2f:offset=%.2f:"
            "dual_mono=true:linear=true:print_format=json" % (ln.input_i, ln.input_tp - 3.0, max(ln.input_lra, 0.5), ln.input_thresh, -16.0 - ln.input_i))
    got = ctx.run_graph(spec, s16, 44100, want_meta=False)
    exp = OG.run_spec(spec, s16, 44100)
    assert got["loudnorm"].normalization_type == exp["loudnorm"]["normalization_type"]
    if got["loudnorm"].normalization_type == 0:
        assert np.max(np.abs(got["pcm"] - exp["pcm"])) < 1e-12
        for k in ("input_i", "output_i", "output_tp", "output_lra", "target_offset"):
            assert abs(getattr(got["loudnorm"], k) - exp["loudnorm"][k]) < 5e-3, k


def test_band_rms_17(ctx, speech48):
    centres = [80, 125, 195, 290, 440, 660, 1000, 1500, 2250, 3350, 5000, 7500, 11200, 16000, 24000]
    lo, hi = [1000.0, 6000.0], [3000.0, 9000.0]
    for i in range(15):
        lo.append(centres[0] / (centres[1] / centres[0]) ** 0.5 if i == 0 else (centres[i - 1] * centres[i]) ** 0.5)
        hi.append(centres[14] * (centres[14] / centres[13]) ** 0.5 if i == 14 else (centres[i] * centres[i + 1]) ** 0.5)
    start, dur = 3.25, 9.5
    got, found = ctx.band_rms(speech48, 48000, start, dur, lo, hi)
    s0, n = round(start * 48000), round(dur * 48000)
    reg = speech48[s0:s0 + n]
    for b in range(17):
        y = O.biquad(O.biquad(reg, 48000, "highpass", lo[b]), 48000, "lowpass", hi[b])
        exp = O.astats(y, 48000)["RMS_level"]
        # (the 24 kHz band's 29.4 kHz low-pass lies past Nyquist: af_biquads.c config_filter() puts it in bypass, so the
        #  band measures the 19.6 kHz high-passed region -- both sides implement that, and the band is compared like the rest)
        if np.isfinite(exp):
            assert found[b] == 1 and abs(got[b] - exp) < 2e-3, (b, got[b], exp)
    assert np.isfinite(got[16])


def test_unsupported_and_bad_specs_fail_loudly(ctx):
    x = np.zeros(4800, dtype=np.float32)
    with pytest.raises(gpudsp.JtError) as e:
        ctx.run_graph("aecho=0.8:0.9:1000:0.3", x, 48000)
    assert e.value.code == -5
    with pytest.raises(gpudsp.JtError) as e:
        ctx.run_graph("highpass=f=abc", x, 48000)
    assert e.value.code == -4
    with pytest.raises(gpudsp.JtError):
        ctx.run_graph("highpass=f=80", np.zeros(9600, dtype=np.float32), 48000, channels=2)
