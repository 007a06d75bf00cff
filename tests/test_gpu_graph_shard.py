"""Passes 2, 3 and 4 of ONE stream cut into chunks (BASELINE.json configs[3]): jt_graph_chunk on every
chunk + jt_graph_merge must reproduce jt_run_graph on the whole stream -- audio within the round-off the f32 stages
carry anyway (a biquad / NLM sum started elsewhere differs in the last bits; bound: 1 LSB of the s16 output), the
measurements through the "%.3f" / "%.2f" wire.  afftdn's tracked noise floor (unbounded memory) crosses the cuts
through the exchange callback, everything else through context."""
import math
import os
import socket
import numpy as np
import pytest
import torch

from jivetalking_b200 import gpudsp, shard, synth

pytestmark = pytest.mark.gpu


def rms(a):
    return float(np.sqrt(np.mean(np.square(a.astype(np.float64))))) if len(a) else 0.0


def pcm_same_s16(g, e, frac_tol=2e-3):
    assert g.dtype == np.int16 and e.dtype == np.int16 and len(g) == len(e), (g.dtype, len(g), len(e))
    d = g.astype(np.int32) - e.astype(np.int32)
    assert rms(d) / 32768.0 < 2e-5, rms(d)                       # well inside the 1e-4 RMS contract
    assert np.mean(np.abs(d) > 1) < frac_tol, float(np.mean(np.abs(d) > 1))    # beyond 1 LSB: rare (adeclick flag flips)


def meta_same(got, exp, r128_tol=0.002):
    assert len(got) == len(exp)
    for a, b in zip(got, exp):
        assert a.first_sample == b.first_sample and a.nb_samples == b.nb_samples
        for k in ("r128_M", "r128_S", "r128_I", "r128_LRA", "r128_true_peak", "r128_sample_peak"):
            x, y = getattr(a, k), getattr(b, k)
            assert math.isnan(x) == math.isnan(y), (k, x, y)
            if not math.isnan(x) and y > -100.0:
                assert abs(x - y) <= r128_tol * max(1.0, abs(y) / 10), (k, a.first_sample, x, y)
    # spectral rows: behind the f32 stages the two runs differ by their round-off noise (~1e-5 of the signal), which IS the
    # signal of a statistic in a silent frame (flux, skewness near 0, "decrease" dividing by the DC bin after an 80 Hz
    # high-pass): per frame the bound is relative to the statistic's typical size over the stream, and the stream mean
    # (what the Go side accumulates, analyser_metrics.go:488-621) must agree to 2e-3
    G = np.array([[m.spectral[k] for k in range(gpudsp.SP_COUNT)] for m in got])
    E = np.array([[m.spectral[k] for k in range(gpudsp.SP_COUNT)] for m in exp])
    assert np.array_equal(np.isnan(G), np.isnan(E))
    rows = ~np.isnan(E[:, 0])
    if rows.any():
        G, E = G[rows], E[rows]
        scale = np.median(np.abs(E), axis=0)
        for k, name in enumerate(gpudsp.SP_NAMES):
            tol = 5e-3 * np.abs(E[:, k]) + 5e-2 * scale[k] + 1e-7
            bad = np.abs(G[:, k] - E[:, k]) > tol
            assert not bad.any(), (name, int(bad.sum()), G[bad, k][:3], E[bad, k][:3], scale[k])
            assert abs(G[:, k].mean() - E[:, k].mean()) <= 2e-3 * abs(E[:, k].mean()) + 1e-3 * scale[k], (name, G[:, k].mean(), E[:, k].mean())


def last_astats(meta):
    rows = [m for m in meta if not math.isnan(m.astats_overall_RMS_level)]
    return rows[-1] if rows else None


def astats_same(got, exp):
    a, b = last_astats(got), last_astats(exp)
    assert (a is None) == (b is None)
    if a is None:
        return
    assert a.first_sample == b.first_sample
    for name in ("RMS_level", "Peak_level", "Min_level", "Max_level", "Number_of_samples", "DC_offset", "Crest_factor"):
        k = gpudsp.AS_NAMES.index(name)
        x, y = a.astats[k], b.astats[k]
        assert math.isnan(x) == math.isnan(y) and (math.isnan(x) or abs(x - y) <= 2e-3 * max(1.0, abs(y))), (name, x, y)
    assert abs(a.astats_overall_RMS_level - b.astats_overall_RMS_level) < 2e-3


def stereo_of(mono):
    r = np.concatenate([np.zeros(7, np.float32), 0.9 * mono[:-7]])           # R = L delayed 7 samples x 0.9
    return np.stack([mono, r], axis=1).reshape(-1).astype(np.float32)


@pytest.fixture(scope="module")
def stream48():
    return synth.speech_like(100.0, 48000, seed=4242)


@pytest.fixture(scope="module")
def pass2_whole(ctx, stream48):
    return ctx.run_graph(gpudsp.default_pass2_spec(), stream48, 48000)


@pytest.mark.parametrize("n_chunks", [2, 5])
def test_pass2_chunked_equals_whole_stream(ctx, stream48, pass2_whole, n_chunks):
    spec = gpudsp.default_pass2_spec()                       # afftdn tn=1: the noise-floor carry crosses every cut
    assert gpudsp.graph_exchanges(spec) == 1
    out, mg = shard.run_graph_sharded(ctx, shard.LocalComm(n_chunks), spec, stream48, 48000, want_meta=True)
    pcm_same_s16(out, pass2_whole["pcm"])
    assert len(out) % 4096 == 0
    meta_same(mg["meta"], pass2_whole["meta"])
    astats_same(mg["meta"], pass2_whole["meta"])


def test_pass2_chunked_96k_stereo(ctx):
    rate = 96000
    mono = synth.speech_like(50.0, rate, seed=99)
    pcm = stereo_of(mono)
    spec = gpudsp.default_pass2_spec()
    whole = ctx.run_graph(spec, pcm, rate, channels=2)
    out, mg = shard.run_graph_sharded(ctx, shard.LocalComm(3), spec, pcm, rate, channels=2, want_meta=True)
    pcm_same_s16(out, whole["pcm"])
    meta_same(mg["meta"], whole["meta"])


def test_noise_floor_carry_is_what_makes_it_exact(ctx, stream48, pass2_whole):
    """Without the exchange a mid-stream chunk cannot know the tracked floor: the library refuses instead of guessing."""
    spec = gpudsp.default_pass2_spec()
    unit = gpudsp.graph_chunk_unit(spec, 48000)
    left, right = gpudsp.graph_chunk_context(spec, 48000)
    n = len(stream48)
    first = (n // 2) // unit * unit
    ctx.set_exchange(None, 1)
    with pytest.raises(gpudsp.JtError):
        ctx.graph_chunk(spec, stream48[first - left:], 48000, 1, first - left, first, n - first, n)
    with pytest.raises(gpudsp.JtError):                        # boundary off the unit grid
        ctx.graph_chunk(spec, stream48, 48000, 1, 0, 100, n - 100, n)
    with pytest.raises(gpudsp.JtError):                        # mid-stream chunk without left context
        ctx.graph_chunk(spec, stream48[first:], 48000, 1, first, first, n - first, n)


@pytest.mark.parametrize("limiter", [False, True])
def test_pass3_chunked_equals_whole_stream(ctx, pass2_whole, limiter):
    x = pass2_whole["pcm"]
    spec, _ = gpudsp.build_pass3_spec(-30.0, -3.0 if limiter else -20.0)
    assert ("alimiter" in spec) == limiter
    whole = ctx.run_graph(spec, x, 44100, want_pcm=False, want_meta=False)["loudnorm"]
    _, mg = shard.run_graph_sharded(ctx, shard.LocalComm(3), spec, x, 44100, want_pcm=False)
    got = mg["loudnorm"]
    assert got.valid == 1 and got.normalization_type == whole.normalization_type
    for k in ("input_i", "input_tp", "input_lra", "input_thresh"):
        assert abs(getattr(got, k) - getattr(whole, k)) < 2e-3, (k, getattr(got, k), getattr(whole, k))


def test_pass4_chunked_equals_whole_stream(ctx, pass2_whole):
    x = pass2_whole["pcm"]
    spec3, plan = gpudsp.build_pass3_spec(-30.0, -12.0)
    p3 = ctx.run_graph(spec3, x, 44100, want_pcm=False, want_meta=False)["loudnorm"]
    spec4, _, _ = gpudsp.build_pass4_spec(plan, p3)
    whole = ctx.run_graph(spec4, x, 44100)
    assert whole["loudnorm"].normalization_type == 0
    out, mg = shard.run_graph_sharded(ctx, shard.LocalComm(3), spec4, x, 44100, want_meta=True)
    pcm_same_s16(out, whole["pcm"], frac_tol=2e-4)            # f64 stages on identical input: flips are rarer still
    meta_same(mg["meta"], whole["meta"])
    astats_same(mg["meta"], whole["meta"])
    for k in ("input_i", "input_tp", "input_lra", "input_thresh", "output_i", "output_tp", "output_lra", "output_thresh", "target_offset"):
        assert abs(getattr(mg["loudnorm"], k) - getattr(whole["loudnorm"], k)) < 2e-3, k


def test_four_pass_chain_sharded_equals_single_gpu(ctx, stream48):
    pcm1, res1 = ctx.process_audio(stream48, 48000)
    pcm, r = shard.process_stream_sharded(ctx, shard.LocalComm(4), stream48, 48000)
    assert len(pcm) == len(pcm1) == res1.n_out
    d = (pcm.astype(np.int32) - pcm1.astype(np.int32)) / 32768.0
    assert rms(d) < 1e-4, rms(d)
    assert abs(r["final"].input_i - res1.final.input_i) <= 0.01
    assert abs(r["final"].input_tp - res1.final.input_tp) <= 0.01
    assert abs(r["final"].input_lra - res1.final.input_lra) <= 0.05
    assert abs(r["filtered"].input_i - res1.filtered.input_i) <= 0.005
    assert abs(r["pass3"].input_i - res1.pass3.input_i) <= 0.005
    assert abs(r["pass4"].output_i - res1.pass4.output_i) <= 0.005


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _nccl_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    x = synth.speech_like(100.0, 48000, seed=4242)
    with gpudsp.Context(rank) as c:
        pcm, r = shard.process_stream_sharded(c, shard.DistComm(dev), x, 48000)
        ok = True
        if rank == 0:
            pcm1, res1 = c.process_audio(x, 48000)
            d = (pcm.astype(np.int32) - pcm1.astype(np.int32)) / 32768.0
            ok = len(pcm) == len(pcm1) and rms(d) < 1e-4 and abs(r["final"].input_i - res1.final.input_i) <= 0.01
        # the adaptive spec derived on every rank from the chunked Pass 1 (shard.adapt_stream_sharded): identical on both
        # ranks and equal to the single-GPU analysis
        from jivetalking_b200 import adapt
        y = synth.podcast_like(80.0, 48000, sibilance_db=-14.0)
        _, _, va, _, _, spec = shard.adapt_stream_sharded(c, y, 48000, device=dev)
        why = ""
        if rank == 0:
            # the momentary / RMS values of the merged intervals are exact, the spectral rows agree to 2e-5 (test_gpu_stream_shard.py):
            # decisions, regions and the emitted spec must be identical, spectral region means only close
            an, _ = adapt.analyse_adaptive(c, y, 48000)
            w = an.voice_activity
            same_regions = ((va.speech_profile.region.start_ns, va.speech_profile.region.end_ns, va.noise_region.start_ns, va.noise_region.end_ns) ==
                            (w.speech_profile.region.start_ns, w.speech_profile.region.end_ns, w.noise_region.start_ns, w.noise_region.end_ns))
            exact = all(getattr(va, k) == getattr(w, k) for k in ("split", "floor", "floor_prescan", "margin", "gap_tolerance", "voiced_low_percentile",
                                                                   "noise_high_percentile", "voice_activated", "n_speech_regions"))
            fl = adapt.SP_NAMES.index("flatness")
            close = abs(va.noise_profile.spectral[fl] - w.noise_profile.spectral[fl]) <= 2e-5 * abs(w.noise_profile.spectral[fl])
            if not (spec == an.pass2_spec.decode() and same_regions and exact and close):
                ok = False
                why = f"spec_equal={spec == an.pass2_spec.decode()} regions={same_regions} exact={exact} close={close} | {spec} | {an.pass2_spec.decode()}"
    q.put((rank, ok, len(pcm), int(np.abs(pcm.astype(np.int64)).sum()), r["final"].input_i, r["final"].input_tp, spec, why if rank == 0 else ""))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_four_pass_chain_two_ranks_over_nccl():
    import torch.multiprocessing as mp
    world, port = 2, _free_port()
    mpc = mp.get_context("spawn")
    q = mpc.Queue()
    procs = [mpc.Process(target=_nccl_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=600) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert res[0][1] and res[1][1], res[0][7]
    assert res[0][2:7] == res[1][2:7]               # every rank holds the same output, measurements and adaptive spec
