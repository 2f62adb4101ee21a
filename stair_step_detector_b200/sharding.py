"""Multi-GPU plumbing of the batched path: frames are independent (reference Pointcloud::process keeps no
cross-frame state, pointcloud.cpp:608-626), so a batch is split into contiguous frame ranges, one process per
GPU, and only per-frame Stairs records / timings are gathered. No collective touches the point data."""


def frame_range(rank, world, total_frames):
    """Contiguous, balanced split of [0, total_frames) over `world` ranks -> (first, count)."""
    base, rem = divmod(total_frames, world)
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


def global_frame_index(rank, frames_per_rank, local_index):
    """Weak-scaling layout used by bench.py: every rank owns frames_per_rank frames."""
    return rank * frames_per_rank + local_index


def reduce_timing(dist, values_max, values_sum, device=None):
    """MAX-reduce timings and SUM-reduce counters over ranks with torch.distributed (nccl or gloo)."""
    import torch
    tmax = torch.tensor(list(values_max), dtype=torch.float64, device=device)
    tsum = torch.tensor(list(values_sum), dtype=torch.int64, device=device)
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
    return [float(x) for x in tmax.cpu()], [int(x) for x in tsum.cpu()]


def gather_step_counts(dist, local_counts, device=None):
    """all_gather of the per-frame step counts (equal-sized shards) -> list indexed by global frame id."""
    import torch
    t = torch.tensor(list(local_counts), dtype=torch.int32, device=device)
    out = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return [int(x) for part in out for x in part.cpu()]
