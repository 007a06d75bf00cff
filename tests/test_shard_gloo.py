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
