"""N > 1 path on CPU: two gloo ranks deal files round-robin and reduce the timing the way bench.py does."""
import os
import socket
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from jivetalking_b200 import shard


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    files = shard.assign_files(5, rank, world)
    samples = 1000 * len(files)
    seconds = 1.0 + rank            # rank 1 is the slow one
    value, total, tmax = shard.job_throughput(samples, seconds)
    dist.barrier()
    out.put((rank, files, [shard.stream_seed(12345, f) for f in files], value, total, tmax))
    dist.destroy_process_group()


def test_two_rank_sharding_and_timing():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, f0, s0, v0, t0, m0), (r1, f1, s1, v1, t1, m1) = res
    assert f0 == [0, 2, 4] and f1 == [1, 3]                 # every file exactly once
    assert s0 == [12345, 12347, 12349] and s1 == [12346, 12348]
    assert t0 == t1 == 5000.0 and m0 == m1 == 2.0           # sum of samples, MAX of times
    assert v0 == v1 == 2500.0


def test_single_process_is_identity():
    assert shard.job_throughput(10.0, 2.0) == (5.0, 10.0, 2.0)
    assert shard.assign_files(3, 0, 1) == [0, 1, 2]


# ---- one stream over several ranks (configs[3]): chunk planning, the all-gather, merge validation -----------
def test_stream_chunk_plan_tiles_the_stream():
    from jivetalking_b200 import gpudsp
    for rate in (44100, 48000, 96000):
        unit = gpudsp.analyse_chunk_unit(rate)
        assert unit % (rate // 10) == 0 and unit % 4096 == 0 and unit % 1024 == 0
        for total in (1, unit - 1, unit, 7 * unit + 123, 64 * unit):
            for world in (1, 2, 3, 8):
                chunks = shard.plan_stream_chunks(total, unit, world)
                assert len(chunks) == world and chunks[0][0] == 0
                end = 0
                for first, owned in chunks:
                    assert first == end and (first % unit == 0 or owned == 0)
                    end = first + owned
                assert end == total
                sizes = [o for _, o in chunks if o and o % unit == 0]
                assert not sizes or max(sizes) - min(sizes) <= unit
                lo, hi = shard.local_range(*chunks[-1], total, unit)
                assert 0 <= lo <= chunks[-1][0] and hi <= total


def _gather_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    blob = bytes([rank + 1]) * (10 + 7 * rank)          # ragged sizes
    got = shard.allgather_blobs(blob)
    out.put((rank, got))
    dist.destroy_process_group()


def test_allgather_of_ragged_blobs_two_ranks():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gather_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in range(world):
        assert res[r] == [b"\x01" * 10, b"\x02" * 17]


def test_merge_rejects_blobs_that_do_not_tile():
    import pytest
    from jivetalking_b200 import gpudsp
    with pytest.raises(gpudsp.JtError):
        gpudsp.analyse_merge([b"\x00" * 256], 48000, 48000)     # no magic


# ---- Passes 2-4 of one stream over several ranks: geometry, the exchange, the PCM gather (host logic only) ---------
def test_graph_chunk_geometry():
    from jivetalking_b200 import gpudsp
    p2 = gpudsp.default_pass2_spec()
    assert gpudsp.graph_chunk_unit(p2, 48000) == 4800          # 100 ms ticks; afftdn hop 600 and the 160:147 period divide it
    assert gpudsp.graph_chunk_unit(p2, 96000) == 9600
    assert gpudsp.graph_exchanges(p2) == 1                     # afftdn tn=1: the tracked noise floor crosses the cuts
    assert gpudsp.graph_exchanges(p2.replace("tn=1", "tn=0")) == 0
    p3, plan = gpudsp.build_pass3_spec(-30.0, -3.0)
    assert gpudsp.graph_chunk_unit(p3, 44100) == 4410          # 44.1 -> 192 kHz: 100 ms is 19200 samples there
    st = gpudsp.LoudnormStats(input_i=-30.0, input_tp=-12.0, input_lra=5.0, input_thresh=-40.0, valid=1)
    p4, _, _ = gpudsp.build_pass4_spec(plan, st)
    assert gpudsp.graph_chunk_unit(p4, 44100) == 890820        # lcm(4410-sample tick, 1212-sample adeclick hop)
    for spec, rate in ((p2, 48000), (p2, 96000), (p3, 44100), (p4, 44100)):
        unit = gpudsp.graph_chunk_unit(spec, rate)
        left, right = gpudsp.graph_chunk_context(spec, rate)
        assert left % unit == 0 and right % unit == 0 and left >= 8 * rate and right >= rate
    assert gpudsp.graph_chunk_unit("atrim=start=1,astats", 48000) == 0     # not chunkable


def test_graph_merge_rejects_garbage():
    import pytest
    from jivetalking_b200 import gpudsp
    with pytest.raises(gpudsp.JtError):
        gpudsp.graph_merge(gpudsp.default_pass2_spec(), [b"\x00" * 512], 48000, 48000)


def test_local_comm_feeds_earlier_carries_only():
    comm = shard.LocalComm(3)
    comm.begin_pass()
    a = comm.exchange(b"A" * 32)
    b = comm.exchange(b"B" * 32)
    assert a == b"A" * 32 + bytes(64) and b == b"A" * 32 + b"B" * 32 + bytes(32)
    comm.begin_pass()
    assert comm.exchange(b"C" * 32)[:32] == b"C" * 32


def _dist_comm_worker(rank, world, port, out):
    import numpy as np
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    comm = shard.DistComm()
    got = comm.exchange(bytes([rank + 1]) * 32)
    pcm = comm.gather_pcm([np.arange(3 + 2 * rank, dtype=np.int16) + 100 * rank])
    blobs = comm.gather_blobs([b"x" * (5 + rank)])
    out.put((rank, got, pcm.tolist(), blobs))
    dist.destroy_process_group()


def test_dist_comm_exchange_and_gathers_two_ranks():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_dist_comm_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict((r, rest) for r, *rest in (q.get(timeout=120) for _ in range(world)))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in range(world):
        got, pcm, blobs = res[r]
        assert got == b"\x01" * 32 + b"\x02" * 32                      # rank order
        assert pcm == [0, 1, 2, 100, 101, 102, 103, 104]                # ragged chunks, stream order
        assert blobs == [b"x" * 5, b"x" * 6]


# ---- jt_process_audio_sharded's host side over gloo: the plan every rank derives and the all-gather callback contract ----------
def _plan_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from jivetalking_b200 import adapt
    total, rate = 37 * 60 * 96000 + 1234, 96000
    p = adapt.sharded_plan(total, rate, world, rank)
    comm = shard.DistComm(None)
    # the exchange primitive jt_set_exchange expects: fixed-size records, every rank's in rank order (an all-gather)
    rec = bytes([rank + 1]) * 24
    got = comm.exchange(rec)
    # variable-length payloads are built on it inside the library: lengths first, then padded payloads
    lens = comm.exchange(int(p.owned).to_bytes(8, "little"))
    # the in-place variant bench.py installs (jt_set_exchange with raw=True) gives the same bytes
    import ctypes as C
    sb = (C.c_ubyte * 24).from_buffer_copy(rec)
    rb = (C.c_ubyte * (24 * world))()
    comm.exchange_raw(C.addressof(sb), 24, C.addressof(rb))
    assert bytes(rb) == got
    dist.barrier()
    out.put((rank, (p.unit, p.own_first, p.owned, p.local_first, p.n_local), got, lens, total))
    dist.destroy_process_group()


def test_sharded_plan_and_exchange_over_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_plan_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, p0, g0, l0, total), (_, p1, g1, l1, _) = res
    assert g0 == g1 == bytes([1]) * 24 + bytes([2]) * 24 and l0 == l1
    assert [int.from_bytes(l0[i * 8:(i + 1) * 8], "little") for i in range(2)] == [p0[2], p1[2]]
    unit = p0[0]
    assert p0[1] == 0 and p1[1] == p0[2] and p0[2] + p1[2] == total and p0[2] % unit == 0          # the chunks tile the stream on the unit
    assert p1[3] % unit == 0 and p1[1] - p1[3] >= 8 * 96000 and p0[4] - p0[2] >= 96000             # 8 s of left context, 1 s of right
    assert p0[3] == 0 and p1[3] + p1[4] == total
