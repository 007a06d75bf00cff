import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
from jivetalking_b200 import adapt as A, gpudsp, shard, synth
x = synth.stereo_from_mono(synth.podcast_like(100.0, 96000, seed=72))
ctxs = [gpudsp.Context(0) for _ in range(2)]
pcm, infos = shard.process_stream_sharded_call(ctxs, x, 96000, channels=2)
ref, res1, an1 = A.process_audio_adaptive(ctxs[0], x, 96000, 2)
r0, a0, t0 = infos[0]
print("spec equal", a0.pass2_spec == an1.pass2_spec)
print(a0.pass2_spec.decode()); print(an1.pass2_spec.decode())
for k in ("input_i", "input_tp", "input_lra", "input_thresh"):
    print(k, getattr(r0.pass3, k), getattr(res1.pass3, k))
print("filtered", r0.filtered.input_i, res1.filtered.input_i, r0.filtered.input_tp, res1.filtered.input_tp)
print("plan", r0.limiter_needed, res1.limiter_needed, r0.limiter_ceiling_db, res1.limiter_ceiling_db, r0.gain_db, res1.gain_db, r0.effective_target_i, res1.effective_target_i)
d = pcm.astype(np.int64) - ref.astype(np.int64)
big = np.abs(d) > 2
print("frac big", big.mean(), "rms lsb", np.sqrt((d * d).mean()))
# where are they
sec = 44100
for s in range(0, len(d), 5 * sec):
    seg = d[s:s + 5 * sec]; rs = ref[s:s + 5 * sec].astype(np.float64)
    ratio = np.sum(seg * rs) / max(np.sum(rs * rs), 1)
    print(s // sec, "s: frac", (np.abs(seg) > 2).mean().round(4), "gain ratio", ratio)
