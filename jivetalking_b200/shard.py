"""Multi-GPU structure of the path: files per GPU (BASELINE.json configs[2]) and, for one long stream,
contiguous chunks per GPU with ONE all-gather of mergeable analysis values (configs[3]).

The reference's data parallelism is one goroutine per file behind a semaphore
(cmd/jivetalking/pool.go:122-153); here it is one process per GPU, files dealt round-robin, no
data-path collective.  torch.distributed (NCCL on GPUs, gloo in the CPU tests) carries only the
barrier and the max-over-ranks timing the bench contract asks for."""
import torch
import torch.distributed as dist


def assign_files(n_files, rank, world):
    """Indices of the files rank `rank` processes (round-robin, like a worker pool draining a queue)."""
    return list(range(rank, n_files, world))


def stream_seed(base_seed, file_index):
    """C3: 8 copies of the C2 recipe with seeds 12345..12352."""
    return base_seed + file_index


def job_throughput(samples_local, seconds_local, device=None):
    """Whole-job samples/s: sum of samples over ranks / max of times over ranks."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return samples_local / seconds_local, samples_local, seconds_local
    t = torch.tensor([float(seconds_local)], dtype=torch.float64, device=device)
    s = torch.tensor([float(samples_local)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(s, op=dist.ReduceOp.SUM)
    return float(s[0]) / float(t[0]), float(s[0]), float(t[0])


# ---- one long stream over several GPUs (configs[3]) ------------------------------------------------
def plan_stream_chunks(total_frames, unit, world):
    """Contiguous chunks [(first, owned)] tiling [0, total_frames): every boundary a multiple of `unit`
    (jt_analyse_chunk_unit: whole ticks / decoder frames / spectral hops), sizes balanced to within one
    unit, the last chunk takes the partial tail.  Ranks beyond the number of units get (total, 0)."""
    units = max(1, -(-total_frames // unit))
    per, extra = divmod(units, world)
    chunks, first = [], 0
    for r in range(world):
        n_units = per + (1 if r < extra else 0)
        owned = min(n_units * unit, max(total_frames - first, 0))
        chunks.append((first, owned))
        first += owned
    return chunks


def local_range(first, owned, total_frames, unit):
    """The frames a rank must hold to analyse its chunk: one unit of left context in mid-stream, one unit
    (or up to the stream's end) of right context."""
    lo = max(0, first - unit)
    hi = min(total_frames, first + owned + unit)
    return lo, hi


def allgather_blobs(blob, device=None):
    """Every rank's byte blob on every rank: one all-gather of padded uint8 tensors (NCCL on GPUs, gloo on
    CPU) plus one of the lengths.  Returns a list of bytes objects in rank order."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [bytes(blob)]
    world = dist.get_world_size()
    n = torch.tensor([len(blob)], dtype=torch.int64, device=device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n)
    cap = int(max(int(x[0]) for x in sizes))
    buf = torch.zeros(cap, dtype=torch.uint8, device=device)
    if len(blob):
        buf[: len(blob)] = torch.frombuffer(bytearray(blob), dtype=torch.uint8).to(buf.device)
    parts = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf)
    return [bytes(parts[r][: int(sizes[r][0])].cpu().numpy().tobytes()) for r in range(world)]


def analyse_stream_sharded(ctx, pcm, rate, channels=1, device=None):
    """Pass-1 analysis of ONE stream by all ranks of the process group.  `pcm` is the whole stream here (tests,
    bench); a production caller decodes only local_range().  Returns (Measurements, [Interval]) on every rank."""
    from . import gpudsp
    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    total = pcm.size // channels
    unit = gpudsp.analyse_chunk_unit(rate)
    first, owned = plan_stream_chunks(total, unit, world)[rank]
    blob = b""
    if owned > 0:
        lo, hi = local_range(first, owned, total, unit)
        blob = ctx.analyse_chunk(pcm.reshape(-1)[lo * channels: hi * channels], rate, channels, lo, first, owned, total)
    blobs = [b for b in allgather_blobs(blob, device=device) if len(b)]
    return gpudsp.analyse_merge(blobs, total, rate)
