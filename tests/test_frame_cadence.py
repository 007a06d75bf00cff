"""Sink-frame cadence of the graph executor (csrc/jt_graph.cu) against a literal, level-by-level restatement of libavfilter's
re-framing (ff_inlink_consume_samples with min = max = F: a new frame takes its properties from the queued frame holding its
first sample and can be built once its last sample has arrived).  The executor collapses consecutive uniform re-framings in one
pass (frames_collapse) instead of materialising every intermediate list; this test holds that shortcut to the plain rule on the
chains the reference's specs produce (filters.go:42-68, normalise.go:1231-1334) and on random ones.  Host-only (dry plan)."""
import ctypes as C
import random

import numpy as np
import pytest

from jivetalking_b200 import gpudsp


def plan(spec, n, rate, frame=4096):
    L = gpudsp.lib()
    L.jt_debug_graph_frames.restype = C.c_int64
    L.jt_debug_graph_frames.argtypes = [C.c_char_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int64), C.c_int64]
    nf = L.jt_debug_graph_frames(spec.encode(), n, rate, 1, gpudsp.FMT_FLT, frame, 0, None, 0)
    assert nf >= 0, nf
    out = np.zeros((max(nf, 1), 6), dtype=np.int64)
    got = L.jt_debug_graph_frames(spec.encode(), n, rate, 1, gpudsp.FMT_FLT, frame, 0, out.ctypes.data_as(C.POINTER(C.c_int64)), nf)
    assert got == nf
    return [tuple(int(v) for v in row) for row in out[:nf]]


def source(n, F):
    return [dict(start=s, nb=min(F, n - s), ready=s + min(F, n - s), astats_pos=-1, hop=-1, tick=-1) for s in range(0, n, F)]


def reframe(old, n, F):
    new, a, b = [], 0, 0
    for s in range(0, n, F):
        f = dict(start=s, nb=min(F, n - s), ready=0, astats_pos=-1, hop=-1, tick=-1)
        while a + 1 < len(old) and old[a + 1]["start"] <= s:
            a += 1
        last = s + f["nb"] - 1
        b = max(b, a)
        while b + 1 < len(old) and old[b + 1]["start"] <= last:
            b += 1
        if old:
            f.update(astats_pos=old[a]["astats_pos"], hop=old[a]["hop"], tick=old[a]["tick"], ready=old[b]["ready"])
        new.append(f)
    return new


def literal(nodes, n, rate, frame=4096):
    fr = source(n, frame)
    for name, arg in nodes:
        if name == "anlmdn":
            fr = reframe(fr, n, 2 * int(round(arg * rate)) + 1)
        elif name == "afftdn":
            fr = reframe(fr, n, rate // 80)
        elif name == "adeclick":
            ws = int(rate * arg[0] / 1000.0)
            fr = reframe(fr, n, max(int(ws * (1.0 - arg[1] / 100.0)), 1))
        elif name == "astats":
            for f in fr:
                f["astats_pos"] = f["start"] + f["nb"]
        elif name == "aspectralstats":
            fr = reframe(fr, n, 1024)
            for j, f in enumerate(fr):
                f["hop"] = j
        elif name == "ebur128":
            tick = rate // 10
            fr = reframe(fr, n, tick)
            for k, f in enumerate(fr):
                if f["nb"] == tick:
                    f["tick"] = k
    return [(f["start"], f["nb"], f["ready"], f["astats_pos"], f["hop"], f["tick"]) for f in fr]


TEXT = {"anlmdn": lambda a: f"anlmdn=s=0.00001:p={a:.4f}:r=0.0020:m=3", "afftdn": lambda a: "afftdn=nr=12:nt=w:tn=0:nf=-50",
        "adeclick": lambda a: f"adeclick=t=1.7:w={a[0]}:o={a[1]}:m=s", "astats": lambda a: "astats=metadata=1:measure_perchannel=all",
        "aspectralstats": lambda a: "aspectralstats=win_size=2048:win_func=hann:measure=all",
        "ebur128": lambda a: "ebur128=metadata=1:peak=sample+true:dualmono=true:target=-16"}


def spec_of(nodes):
    return ",".join(["aformat=channel_layouts=mono"] + [TEXT[n](a) for n, a in nodes])


PASS1 = [("astats", None), ("aspectralstats", None), ("ebur128", None)]
PASS2 = [("anlmdn", 0.006), ("afftdn", None)] + PASS1
PASS4 = [("adeclick", (55, 50))] + PASS1


@pytest.mark.parametrize("nodes", [PASS1, PASS2, PASS4, [("anlmdn", 0.006)], [("ebur128", None), ("aspectralstats", None), ("astats", None)],
                                   [("aspectralstats", None), ("astats", None), ("anlmdn", 0.002), ("ebur128", None), ("astats", None)]])
@pytest.mark.parametrize("rate,n", [(48000, 0), (48000, 1), (48000, 4095), (48000, 4800), (48000, 100003), (44100, 61234), (96000, 250001)])
def test_reference_chains(nodes, rate, n):
    assert plan(spec_of(nodes), n, rate) == literal(nodes, n, rate)


@pytest.mark.parametrize("seed", range(40))
def test_random_chains(seed):
    rnd = random.Random(seed)
    rate = rnd.choice([44100, 48000, 96000, 32000])
    n = rnd.choice([rnd.randint(1, 3000), rnd.randint(3000, 200000)])
    frame = rnd.choice([4096, 1024, 1152, 4608])
    pool = [("anlmdn", rnd.choice([0.002, 0.006, 0.0004])), ("afftdn", None), ("adeclick", (rnd.choice([55, 20]), rnd.choice([50, 75]))),
            ("astats", None), ("aspectralstats", None), ("ebur128", None)]
    nodes = [rnd.choice(pool) for _ in range(rnd.randint(1, 7))]
    assert plan(spec_of(nodes), n, rate, frame) == literal(nodes, n, rate, frame)
