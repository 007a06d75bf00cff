"""jt_prefetch_input: the double-buffered upload of a worker that processes file after file.  A call that finds its input
prefetched must give exactly what the plain call gives; a prefetch nobody consumes, a different pointer or a different size must
not matter."""
import numpy as np
import pytest

from jivetalking_b200 import adapt as A
from jivetalking_b200 import gpudsp, synth

pytestmark = pytest.mark.gpu


def test_prefetched_input_gives_the_same_result(ctx):
    torch = pytest.importorskip("torch")
    xs = [synth.podcast_like(40.0, 48000, seed=s) for s in (31, 32, 33)]
    want = [A.process_audio_adaptive(ctx, x, 48000) for x in xs]
    pinned = [torch.from_numpy(x).pin_memory() for x in xs]
    out = torch.empty(int(len(xs[0]) * 44100 / 48000) + 3 * 4096, dtype=torch.int16).pin_memory()
    ctx.prefetch_input_ptr(pinned[0].data_ptr(), len(xs[0]), 1, gpudsp.FMT_FLT)
    for k in range(3):
        if k + 1 < 3:
            ctx.prefetch_input_ptr(pinned[k + 1].data_ptr(), len(xs[k + 1]), 1, gpudsp.FMT_FLT)
        res, an = A.process_audio_adaptive_ptr(ctx, pinned[k].data_ptr(), len(xs[k]), 48000, 1, gpudsp.FMT_FLT, out.data_ptr(), out.numel(), False)
        pcm_w, res_w, an_w = want[k]
        assert res.n_out == len(pcm_w) and np.array_equal(out[: res.n_out].numpy(), pcm_w)
        assert an.pass2_spec == an_w.pass2_spec and res.final.input_i == res_w.final.input_i


def test_unused_and_mismatched_prefetches_are_harmless(ctx):
    torch = pytest.importorskip("torch")
    a, b = synth.podcast_like(20.0, 48000, seed=41), synth.podcast_like(20.0, 48000, seed=42)
    pa = torch.from_numpy(a).pin_memory()
    ctx.prefetch_input_ptr(pa.data_ptr(), len(a), 1, gpudsp.FMT_FLT)          # never consumed: b is processed instead
    pcm_b, _, _ = A.process_audio_adaptive(ctx, b, 48000)
    ctx.prefetch_input_ptr(pa.data_ptr(), len(a) - 4800, 1, gpudsp.FMT_FLT)   # same pointer, other size: plain upload path
    pcm_a, _, _ = A.process_audio_adaptive(ctx, a, 48000)
    with gpudsp.Context(0) as fresh:
        assert np.array_equal(pcm_b, A.process_audio_adaptive(fresh, b, 48000)[0])
        assert np.array_equal(pcm_a, A.process_audio_adaptive(fresh, a, 48000)[0])
    with pytest.raises(gpudsp.JtError):
        ctx.prefetch_input_ptr(0, 100, 1, gpudsp.FMT_FLT)
