"""Kernel-group timings (CUDA events on the library's stream) of one filter spec over a synthetic stream.
usage: python scripts/time_filter.py '<spec>' [minutes] [rate] [dtype f32|f64|s16]"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jivetalking_b200 import gpudsp, synth

spec = sys.argv[1]
minutes = float(sys.argv[2]) if len(sys.argv) > 2 else 10
rate = int(sys.argv[3]) if len(sys.argv) > 3 else 48000
dt = sys.argv[4] if len(sys.argv) > 4 else "f32"
x = np.concatenate([synth.speech_like(min(600.0, minutes * 60 - m * 600), rate, seed=12345 + m) for m in range(int(np.ceil(minutes / 10)))])
if dt == "f64":
    x = x.astype(np.float64)
elif dt == "s16":
    x = np.clip(np.round(x * 32768), -32768, 32767).astype(np.int16)
with gpudsp.Context(0) as ctx:
    ctx.run_graph(spec, x, rate, want_meta=False)
    ctx.reset_counters(); ctx.enable_timing(True)
    for _ in range(3):
        ctx.run_graph(spec, x, rate, want_meta=False)
    print({k: round(ms / 3, 3) for k, ms, n in ctx.kernel_timings()})
