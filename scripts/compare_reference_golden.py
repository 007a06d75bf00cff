#!/usr/bin/env python
"""Compare libjtdsp with golden vectors dumped from the REFERENCE itself (integration/golden_dump_test.go, run on a box with
Go + libffmpeg.a):  python scripts/compare_reference_golden.py /path/to/jt_golden   (needs a B200 for libjtdsp).

For every <case>.json: the same WAV (located with jt_wav_parse) and the same spec string go through jt_run_graph; sink-frame
count, nb_samples, every metadata key the reference saw (lavfi.r128.*, lavfi.astats.*, lavfi.aspectralstats.*) and the sink
PCM are compared with the tolerances of tests/ (r128 through the %.3f wire, astats 2e-6, spectral 2e-3 relative, s16 PCM
1e-4 RMS of full scale / float PCM 1e-5).  Prints one line per case and a summary; exit status 1 when any case is outside."""
import json
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from jivetalking_b200 import gpudsp

R128 = {"lavfi.r128.M": "r128_M", "lavfi.r128.S": "r128_S", "lavfi.r128.I": "r128_I", "lavfi.r128.LRA": "r128_LRA",
        "lavfi.r128.LRA.low": "r128_LRA_low", "lavfi.r128.LRA.high": "r128_LRA_high",
        "lavfi.r128.true_peak": "r128_true_peak", "lavfi.r128.sample_peak": "r128_sample_peak"}
NP_OF_AVFMT = {1: np.int16, 3: np.float32, 4: np.float64}


def close(a, b, atol, rtol=0.0):
    if math.isnan(a) or math.isnan(b):
        return math.isnan(a) and math.isnan(b)
    if math.isinf(a) or math.isinf(b):
        return a == b
    return abs(a - b) <= atol + rtol * abs(b)


def compare_case(ctx, d, path):
    case = json.load(open(path))
    pcm, rate, ch = gpudsp.wav_parse(open(os.path.join(d, case["wav"]), "rb").read())
    got = ctx.run_graph(case["spec"], pcm, rate, channels=ch, want_pcm=bool(case.get("pcm")))
    bad = []
    if len(got["meta"]) != len(case["frames"]):
        bad.append(f"sink frames {len(got['meta'])} vs {len(case['frames'])}")
    for i, (g, e) in enumerate(zip(got["meta"], case["frames"])):
        if g.nb_samples != e["nb_samples"]:
            bad.append(f"frame {i}: nb_samples {g.nb_samples} vs {e['nb_samples']}")
            break
        for key, val in e["meta"].items():
            v = float(val)
            if key in R128:
                if not close(getattr(g, R128[key]), v, 0.0011):
                    bad.append(f"frame {i}: {key} {getattr(g, R128[key])} vs {v}")
            elif key.startswith("lavfi.astats.1."):
                name = key[len("lavfi.astats.1."):]
                if name in gpudsp.AS_NAMES and not close(g.astats[gpudsp.AS_NAMES.index(name)], v, 2e-6, 1e-9):
                    bad.append(f"frame {i}: {key} {g.astats[gpudsp.AS_NAMES.index(name)]} vs {v}")
            elif key == "lavfi.astats.Overall.RMS_level" and not close(g.astats_overall_RMS_level, v, 2e-6):
                bad.append(f"frame {i}: {key} {g.astats_overall_RMS_level} vs {v}")
            elif key == "lavfi.astats.Overall.Peak_level" and not close(g.astats_overall_Peak_level, v, 2e-6):
                bad.append(f"frame {i}: {key} {g.astats_overall_Peak_level} vs {v}")
            elif key.startswith("lavfi.aspectralstats.1."):
                name = key[len("lavfi.aspectralstats.1."):]
                if name in gpudsp.SP_NAMES and not close(g.spectral[gpudsp.SP_NAMES.index(name)], v, 1e-9, 2e-3):
                    bad.append(f"frame {i}: {key} {g.spectral[gpudsp.SP_NAMES.index(name)]} vs {v}")
        if len(bad) > 8:
            break
    if case.get("pcm"):
        fmt = case["frames"][0]["format"] if case["frames"] else 1
        ref = np.fromfile(os.path.join(d, case["pcm"]), dtype=NP_OF_AVFMT[fmt])
        out = got["pcm"]
        if len(ref) != len(out) or ref.dtype != out.dtype:
            bad.append(f"pcm {out.dtype}[{len(out)}] vs {ref.dtype}[{len(ref)}]")
        else:
            scale = 32768.0 if ref.dtype == np.int16 else 1.0
            diff = (out.astype(np.float64) - ref.astype(np.float64)) / scale
            rms = float(np.sqrt(np.mean(diff * diff))) if len(diff) else 0.0
            if rms > (1e-4 if ref.dtype == np.int16 else 1e-5):
                bad.append(f"pcm rms diff {rms:.3e}, max {float(np.max(np.abs(diff))):.3e}")
    return case["name"], bad


def main():
    d = sys.argv[1]
    failures = 0
    with gpudsp.Context(0) as ctx:
        for f in sorted(os.listdir(d)):
            if not f.endswith(".json"):
                continue
            try:
                name, bad = compare_case(ctx, d, os.path.join(d, f))
            except gpudsp.JtError as e:
                name, bad = f, [f"libjtdsp: {e}"]
            print(("ok   " if not bad else "DIFF ") + name + ("" if not bad else ": " + "; ".join(bad[:4])))
            failures += bool(bad)
    print(f"{failures} case(s) outside tolerance")
    return 1 if failures else 0


if __name__ == "__main__":
    sys.exit(main())
