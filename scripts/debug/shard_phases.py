import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
from jivetalking_b200 import adapt as A, gpudsp, synth
minutes = int(sys.argv[1]) if len(sys.argv) > 1 else 20
x = np.concatenate([synth.podcast_like(600.0, 48000, seed=5 + m) for m in range(minutes // 10)])
ctx = gpudsp.Context(0)
for rep in range(2):
    t0 = time.perf_counter(); pcm, res, an = A.process_audio_adaptive(ctx, x, 48000); t1 = time.perf_counter() - t0
print("adaptive unchunked: %.1f ms" % (t1 * 1e3))
for rep in range(2):
    ctx.reset_counters(); ctx.enable_timing(True)
    t0 = time.perf_counter(); own, first, res2, an2, tm = A.process_audio_sharded(ctx, x, 48000, 1, len(x), 1, 0); t2 = time.perf_counter() - t0
    tk = ctx.kernel_timings(); ctx.enable_timing(False)
print("sharded world=1: %.1f ms" % (t2 * 1e3))
print({k: round(v * 1e3, 2) for k, v in tm.as_dict().items() if isinstance(v, float)})
print(sorted([(round(ms, 2), k) for k, ms, n in tk], reverse=True)[:25])
