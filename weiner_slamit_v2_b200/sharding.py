"""Multi-GPU plumbing for the ORB front end: frames (and frame pairs) are independent, so N GPUs
means N contiguous shards of the batch axis and NO collective on the data path.
The only cross-rank exchange is a final gather of counters / the max of the elapsed time, done
with torch.distributed (NCCL on GPUs, gloo in the CPU tests)."""
import torch
import torch.distributed as dist


def shard_range(total, rank, world):
    """Contiguous shard [lo, hi) of `total` items for `rank` of `world` (sizes differ by at most 1)."""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world of %d" % (rank, world))
    return total * rank // world, total * (rank + 1) // world


def gather_counters(counters, device="cpu"):
    """Sum a dict of integer counters (frames, keypoints, matches ...) over all ranks."""
    keys = sorted(counters)
    t = torch.tensor([float(counters[k]) for k in keys], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return {k: int(round(v)) for k, v in zip(keys, t.tolist())}


def max_over_ranks(values, device="cpu"):
    """Element-wise max of a list of floats over all ranks (device-timed durations)."""
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.tolist()
