"""BASELINE.json configs[1] at FULL size (60 min, 48 kHz mono f32) through size-independent properties: the oracle
finishes a few seconds of audio, not an hour, so the full-size run is tied to the oracle-checked small runs by
 * causality: every stage is causal up to a bounded look-ahead, so the first minutes of the 60-min Pass-2 output
   must equal the Pass-2 output of the 10-min prefix (the chunk-parallel kernels cut the hour into thousands of
   lanes / hops / windows: any warm-up or boundary error shows up here);
 * linearity of the meters: I(g*x) = I(x) + 20 log10 g, same for the peaks; LRA unchanged;
 * the chain's own contract: s16 / 44.1 kHz output of the expected length, loudness on target, true peak under the ceiling."""
import math
import numpy as np
import pytest

import bench
from jivetalking_b200 import gpudsp

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def hour():
    return bench.make_input(12345, 60)


def test_meters_are_linear_at_full_size(ctx, hour):
    m1, iv1 = ctx.analyse(hour, 48000)
    m2, iv2 = ctx.analyse((hour * np.float32(0.5)).astype(np.float32), 48000)
    db = 20 * math.log10(0.5)
    assert abs((m2.input_i - m1.input_i) - db) < 0.0021            # two "%.3f" roundings
    # peaks travel as LINEAR values printed "%.3f" (analyser_metrics.go:665-670): at 0.1 full scale one count is 0.09 dB
    assert abs((m2.input_tp - m1.input_tp) - db) < 0.1 and abs((m2.input_sp - m1.input_sp) - db) < 0.1
    assert abs(m2.input_lra - m1.input_lra) < 0.011
    assert len(iv1) == len(iv2) and 14000 < len(iv1) < 14100 and m1.sink_frames == 36000    # ~256 ms intervals, 100 ms sink frames
    # (the first intervals hold sink frames without an M key, which count as 0 in the interval mean: analyser_metrics.go:849-872)
    d = [b.momentary_lufs - a.momentary_lufs for a, b in list(zip(iv1, iv2))[4:] if a.momentary_lufs > -60]
    assert abs(np.median(d) - db) < 0.002 and max(abs(v - db) for v in d) < 0.01


def test_full_hour_prefix_causality_and_contract(ctx, hour):
    out, res = ctx.process_audio(hour, 48000)
    # contract (processor.go:379-384, filters.go:523-532): mono s16 44.1 kHz in 4096-sample frames, -16 LUFS, <= -1 dBTP
    assert out.dtype == np.int16 and len(out) % 4096 == 0
    assert abs(len(out) - len(hour) * 44100 / 48000) <= 4096
    assert res.pass4.normalization_type == 0 and abs(res.final.input_i - (-16.0)) <= 0.1
    assert res.final.input_tp <= -1.0 + 0.1
    assert abs(res.pass3.input_i - res.filtered.input_i) < 0.05           # loudnorm's meter vs ebur128 on the same signal
    # causality: Pass 2 of the 10 min prefix == first 10 min of Pass 2 of the hour (minus the stages' look-ahead)
    spec = gpudsp.default_pass2_spec()
    n10 = 10 * 60 * 48000
    p_full = ctx.run_graph(spec, hour, 48000, want_meta=False)["pcm"]
    p_pre = ctx.run_graph(spec, hour[:n10], 48000, want_meta=False)["pcm"]
    keep = len(p_pre) - 44100                                            # the last second feels the end of the prefix
    d = p_full[:keep].astype(np.int32) - p_pre[:keep].astype(np.int32)
    assert np.max(np.abs(d)) <= 1 and np.mean(d != 0) < 1e-4, (int(np.max(np.abs(d))), float(np.mean(d != 0)))
