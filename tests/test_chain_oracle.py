"""The oracle's own composition of ProcessAudio (oracle/chain_oracle.py -- what bench.py's reference arm and the full-path parity
tests run) against the product's host-side planners through the C ABI: band plan, Pass-3 / Pass-4 spec strings (the reference's
golden strings are replayed on the product side in tests/test_abi.py), and a short end-to-end run on CPU.  Host-only."""
import math
import random

import numpy as np

import chain_oracle as CO
from jivetalking_b200 import adapt as A
from jivetalking_b200 import gpudsp, synth


def test_specs_and_band_plan_agree_with_the_product():
    assert CO.PASS1_SPEC == gpudsp.pass1_spec() and CO.DEFAULT_PASS2_SPEC == gpudsp.default_pass2_spec()
    lo, hi = A.band_plan()
    olo, ohi = CO.band_plan()
    assert list(lo) == olo and list(hi) == ohi


def test_pass3_pass4_planners_agree_with_the_product():
    rnd = random.Random(7)
    for _ in range(300):
        out_i, out_tp = rnd.uniform(-50, -8), rnd.uniform(-30, 0.5)
        spec3, plan = gpudsp.build_pass3_spec(out_i, out_tp)
        ospec3, oplan = CO.plan_pass3(out_i, out_tp)
        assert spec3 == ospec3
        assert (plan.limiter_needed != 0, plan.limiter_clamped != 0) == (oplan["needed"], oplan["clamped"])
        assert plan.limiter_ceiling_db == oplan["ceiling"] and plan.limiter_pregain_db == oplan["pre_gain"]
        st = gpudsp.LoudnormStats()
        p3 = dict(input_i=rnd.uniform(-40, -10), input_tp=rnd.uniform(-25, 0), input_lra=rnd.choice([0.0, rnd.uniform(0, 25)]),
                  input_thresh=rnd.uniform(-50, -20))
        for k, v in p3.items():
            setattr(st, k, v)
        spec4, eff, off = gpudsp.build_pass4_spec(plan, st)
        ospec4, oeff, ooff = CO.plan_pass4(oplan, p3)
        assert spec4 == ospec4 and eff == oeff and off == ooff


def test_process_audio_oracle_runs_the_whole_path():
    x = synth.podcast_like(40.0, 48000, seed=5)
    r = CO.process_audio(x, 48000)
    an = r["analysis"]
    assert an["va"]["speech"] is not None and r["spec2"].startswith("aformat=channel_layouts=mono,highpass=f=80")
    assert r["pcm"].dtype == np.int16 and len(r["pcm"]) % 4096 == 0 and len(r["pcm"]) >= 40 * 44100
    assert abs(r["final"][0] - (-16.0)) <= 0.6 and r["final"][1] <= -0.9
    assert math.isfinite(r["p3"]["input_i"]) and r["p4"]["normalization_type"] in (0, 1)
    if an["va"]["noise_profile"] is not None:
        assert r["final_regions"]["room_tone"] is not None
