"""One stream analysed in chunks (BASELINE.json configs[3]): chunk-wise jt_analyse_chunk + host merge
must reproduce the single-GPU jt_analyse of the whole stream -- the exchanged values are per-tick / per-hop /
per-frame, so windows, gating, LRA and intervals are evaluated on the merged lists and do not depend on the cut."""
import math
import os
import socket
import numpy as np
import pytest
import torch

from jivetalking_b200 import gpudsp, shard, synth

pytestmark = pytest.mark.gpu

EXACT_ASTATS = ["RMS_level", "Peak_level", "DC_offset", "Min_level", "Max_level", "Crest_factor", "Dynamic_range",
                "Entropy", "Bit_depth", "Number_of_samples", "Zero_crossings", "Zero_crossings_rate",
                "Max_difference", "Min_difference", "Noise_floor", "Noise_floor_count"]


def same(a, b, tol=0.0):
    if isinstance(a, float) and math.isnan(a):
        return isinstance(b, float) and math.isnan(b)
    return a == b or abs(a - b) <= tol * max(1.0, abs(b))


def check_equal(m, iv, m1, iv1):
    for k in ("input_i", "input_tp", "input_sp", "input_lra", "last_m", "last_s", "sink_frames", "spectral_frames", "duration_s"):
        assert getattr(m, k) == getattr(m1, k), (k, getattr(m, k), getattr(m1, k))
    for k, name in enumerate(gpudsp.SP_NAMES):
        # two hops share one complex FFT, and WHICH hops are paired depends on the cut: f32 round-off of the partner
        # hop leaks in at the 1e-7 level (then "%g" prints 6 digits)
        assert same(m.spectral_mean[k], m1.spectral_mean[k], 2e-5), (name, m.spectral_mean[k], m1.spectral_mean[k])
    for k, name in enumerate(gpudsp.AS_NAMES):
        if name in EXACT_ASTATS:
            assert same(m.astats[k], m1.astats[k], 1e-9), (name, m.astats[k], m1.astats[k])
        elif name in ("RMS_peak", "RMS_trough", "Mean_difference", "RMS_difference"):
            assert same(m.astats[k], m1.astats[k], 1e-6), (name, m.astats[k], m1.astats[k])       # sums in chunk order / 40-tau warm-up
        # Flat_factor: runs of the extreme value are split at chunk boundaries (documented in DESIGN.md)
    assert len(iv) == len(iv1)
    for a, b in zip(iv, iv1):
        for k in ("timestamp_s", "rms_level", "peak_level", "momentary_lufs", "short_term_lufs", "true_peak", "sample_peak",
                  "frame_count", "spectral_found"):
            assert getattr(a, k) == getattr(b, k), (k, getattr(a, k), getattr(b, k))
        for k in range(gpudsp.SP_COUNT):
            assert same(a.spectral[k], b.spectral[k], 2e-5), (k, a.spectral[k], b.spectral[k])


@pytest.fixture(scope="module")
def ctx():
    with gpudsp.Context(0) as c:
        yield c


@pytest.mark.parametrize("rate,channels,world", [(48000, 1, 3), (48000, 1, 8), (96000, 2, 2), (44100, 1, 4)])
def test_chunked_analysis_equals_whole_stream(ctx, rate, channels, world):
    unit = gpudsp.analyse_chunk_unit(rate)
    n = 9 * unit + 12345
    mono = synth.speech_like(n / rate + 0.01, rate, seed=77)[:n]
    if channels == 2:
        r = np.concatenate([np.zeros(7, np.float32), 0.9 * mono[:-7]])       # R = L delayed 7 samples x 0.9
        pcm = np.stack([mono, r], axis=1).reshape(-1).astype(np.float32)
    else:
        pcm = mono
    m1, iv1 = ctx.analyse(pcm, rate, channels)
    blobs = []
    for first, owned in shard.plan_stream_chunks(n, unit, world):
        if owned == 0:
            continue
        lo, hi = shard.local_range(first, owned, n, unit)
        blobs.append(ctx.analyse_chunk(pcm[lo * channels: hi * channels], rate, channels, lo, first, owned, n))
    m, iv = gpudsp.analyse_merge(blobs[::-1], n, rate)          # order of arrival does not matter
    check_equal(m, iv, m1, iv1)


def test_chunk_argument_errors(ctx):
    rate = 48000
    unit = gpudsp.analyse_chunk_unit(rate)
    x = synth.speech_like(8.0, rate, seed=1)
    n = len(x)
    with pytest.raises(gpudsp.JtError):
        ctx.analyse_chunk(x, rate, 1, 0, 100, n - 100, n)             # boundary not a multiple of the unit
    with pytest.raises(gpudsp.JtError):
        ctx.analyse_chunk(x[unit:], rate, 1, unit, unit, n - unit, n)     # mid-stream chunk without left context
    one = ctx.analyse_chunk(x, rate, 1, 0, 0, unit, n) if n > 2 * unit else None
    if one is not None:
        with pytest.raises(gpudsp.JtError):
            gpudsp.analyse_merge([one], n, rate)                      # a single chunk does not tile the stream


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
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    rate = 48000
    unit = gpudsp.analyse_chunk_unit(rate)
    n = 10 * unit + 999
    x = synth.speech_like(n / rate + 0.01, rate, seed=5)[:n]
    with gpudsp.Context(rank) as c:
        m, iv = shard.analyse_stream_sharded(c, x, rate, 1, device=torch.device("cuda", rank))
        if rank == 0:
            m1, iv1 = c.analyse(x, rate, 1)
            check_equal(m, iv, m1, iv1)
    q.put((rank, m.input_i, m.input_tp, m.input_lra, len(iv)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_two_ranks_over_nccl():
    import torch.multiprocessing as mp
    world, port = 2, _free_port()
    mpc = mp.get_context("spawn")
    q = mpc.Queue()
    procs = [mpc.Process(target=_nccl_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert res[0][1:] == res[1][1:]                 # every rank holds the same merged result
