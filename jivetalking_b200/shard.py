"""File-per-GPU sharding (BASELINE.json configs[2]) -- the only multi-GPU structure this path has.

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
