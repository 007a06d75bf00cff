"""which filter of the adapted Pass-2 spec makes a mid-stream chunk differ from the whole-stream run (96 kHz stereo)"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
from jivetalking_b200 import adapt as A, gpudsp, synth
x = synth.stereo_from_mono(synth.podcast_like(100.0, 96000, seed=72))
rate, ch = 96000, 2
total = x.size // ch
ctx = gpudsp.Context(0)
an, iv = A.analyse_adaptive(ctx, x, rate, ch)
spec = an.pass2_spec.decode()
nodes = spec.split(",")
p = A.sharded_plan(total, rate, 2, 1)
win = x[p.local_first * ch:(p.local_first + p.n_local) * ch]
print("plan", p.unit, p.own_first, p.owned, p.local_first, p.n_local)
for upto in range(2, 8):
    sub = ",".join(nodes[:upto])
    try:
        whole = ctx.run_graph(sub, x, rate, channels=ch, want_meta=False)
        part = ctx.graph_chunk(sub, win, rate, ch, p.local_first, p.own_first, p.owned, total, want_blob=False)
    except gpudsp.JtError as e:
        print(upto, nodes[upto - 1][:20], "ERR", e); continue
    w = whole["pcm"][part["out_first"]: part["out_first"] + part["n_out"]].astype(np.float64)
    g = part["pcm"].astype(np.float64)
    d = np.abs(w - g)
    scale = np.abs(w).max() + 1e-30
    k = int(d.argmax())
    print(upto, nodes[upto - 1][:24], "max abs diff", d.max(), "rel", d.max() / scale, "at", k, "of", len(d),
          "mean", d.mean(), "frac>1e-6rel", float((d > 1e-6 * scale).mean()))
